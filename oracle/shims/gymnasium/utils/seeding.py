"""gymnasium.utils.seeding.np_random: PCG64 Generator seeded through SeedSequence (same as gymnasium)."""
import numpy as np


def np_random(seed=None):
    seed_seq = np.random.SeedSequence(seed)
    np_seed = seed_seq.entropy
    return np.random.Generator(np.random.PCG64(seed_seq)), np_seed

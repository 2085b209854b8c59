// hostsim.cpp — HOST build of the per-env bodies in spin_torque_rl_gym_b200/csrc/stt_env_core.cuh.
//
// TEST INFRASTRUCTURE ONLY (lives under tests/, never imported by the product package). It lets the CPU test-suite run the
// exact arithmetic the CUDA kernels execute (same templates, same folded constants) against the NumPy oracle and the golden
// vectors in a container without a GPU. It is not a fallback: the product path always calls libstg.so's CUDA kernels.
#include <stdint.h>
#include <xmmintrin.h>

#include "../../include/stg.h"
#include "../../spin_torque_rl_gym_b200/csrc/llgs_core.cuh"
#include "../../spin_torque_rl_gym_b200/csrc/stt_env_core.cuh"
#include "../../spin_torque_rl_gym_b200/csrc/rk45_core.cuh"
#include "../../spin_torque_rl_gym_b200/csrc/array_core.cuh"
#include <vector>

using namespace stg;

static void store_rows(const StgSttStepArgs& a, int64_t e, const EnvStepResult& r) {
    for (int q = 0; q < kObs; ++q) a.out.obs[e * kObs + q] = r.obs[q];
    if ((a.flags & STG_F_AUTORESET) && a.out.final_obs && r.did_reset)      // rows of running envs are left untouched
        for (int q = 0; q < kObs; ++q) a.out.final_obs[e * kObs + q] = r.final_obs[q];
}
// second pass of the FP32 entry point (stt_env_redo_kernel on the device): envs whose FP32 trajectory is ill-conditioned are
// repeated with FP64 stages, status bit 2 set
template <bool AXIS_Z, int NOISE>
static void redo_f64(const StgSttStepArgs& a, int64_t e) {
    EnvStepResult r;
    env_step_body<double, AXIS_Z, NOISE, false>(a, e, r, STG_STATUS_REDONE_F64);
    store_rows(a, e, r);
}
template <typename R, bool AXIS_Z, int NOISE>
static void step_all(const StgSttStepArgs& a) {
    for (int64_t s = 0; s < a.n_envs; ++s) {
        const int64_t e = (a.flags & STG_F_SORTED) ? a.d_perm[s] : s;
        EnvStepResult r;
        bool done;
        if ((a.flags & STG_F_EULER) && NOISE != 3)
            done = env_step_body<R, AXIS_Z, NOISE == 3 ? 1 : NOISE, true>(a, e, r);
        else
            done = env_step_body<R, AXIS_Z, NOISE, false>(a, e, r);
        if (done) store_rows(a, e, r);
        else redo_f64<AXIS_Z, NOISE>(a, e);
    }
}
template <typename R, bool AXIS_Z>
static void step_noise(const StgSttStepArgs& a) {
    if (a.flags & STG_F_THERMAL_INJECT) step_all<R, AXIS_Z, 2>(a);
    else if ((a.flags & STG_F_THERMAL_PHILOX) && (a.flags & STG_F_STREAM_PHILOX10) && !(a.flags & STG_F_EULER)) step_all<R, AXIS_Z, 3>(a);
    else if (a.flags & STG_F_THERMAL_PHILOX) step_all<R, AXIS_Z, 1>(a);
    else step_all<R, AXIS_Z, 0>(a);
}

// x86 handles denormal operands in microcode (~100x slower); the GPU does not care. Flush them on the host for the f32 runs:
// they are < 1e-38, far below every tolerance.
struct FtzScope {
    unsigned old;
    explicit FtzScope(bool on) : old(_mm_getcsr()) { if (on) _mm_setcsr(old | 0x8040u); }
    ~FtzScope() { _mm_setcsr(old); }
};

// the two-envs-per-thread body with the host emulation of the FP32x2 pack
template <int NOISE>
static void step_pairs(const StgSttStepArgs& a) {
    for (int64_t s = 0; s < a.n_envs; s += 2) {
        const int64_t eA = (a.flags & STG_F_SORTED) ? a.d_perm[s] : s;
        EnvStepResult rA, rB;
        int redo = 0;
        if (s + 1 < a.n_envs) {
            const int64_t eB = (a.flags & STG_F_SORTED) ? a.d_perm[s + 1] : s + 1;
            redo = env_step_pair_body<NOISE>(a, eA, eB, rA, rB);
            if (redo & 2) redo_f64<true, NOISE>(a, eB); else store_rows(a, eB, rB);
        } else {
            redo = env_step_body<float, true, NOISE, false>(a, eA, rA) ? 0 : 1;
        }
        if (redo & 1) redo_f64<true, NOISE>(a, eA); else store_rows(a, eA, rA);
    }
}

extern "C" int hostsim_stt_step(const StgSttStepArgs* a, int f64) {
    if (a->flags & STG_F_EULER) f64 = 1;      // as launch_step<> dispatches: Euler always runs FP64 stages
    FtzScope ftz(!f64);
    if (!f64 && (a->flags & STG_F_AXIS_Z) && !(a->flags & (STG_F_THERMAL_INJECT | STG_F_EULER | STG_F_NO_PAIR))) {
        if ((a->flags & STG_F_THERMAL_PHILOX) && (a->flags & STG_F_STREAM_PHILOX10)) step_pairs<3>(*a);      // as launch_step<> dispatches
        else if (a->flags & STG_F_THERMAL_PHILOX) step_pairs<1>(*a);
        else step_pairs<0>(*a);
        return 0;
    }
    const bool z = (a->flags & STG_F_AXIS_Z) != 0;
    if (f64) { if (z) step_noise<double, true>(*a); else step_noise<double, false>(*a); }
    else     { if (z) step_noise<float, true>(*a);  else step_noise<double, false>(*a); }   // as launch_step<> dispatches
    return 0;
}

extern "C" int hostsim_stt_reset(const StgSttResetArgs* a) {
    for (int64_t e = 0; e < a->n_envs; ++e)
        if (!a->d_mask || a->d_mask[e]) env_reset_body(*a, e);
    return 0;
}

template <typename R, bool AXIS_Z, int NOISE>
static void solve_all(const StgSttSolveArgs& a) {
    for (int64_t e = 0; e < a.n_envs; ++e) {
        if (a.flags & STG_F_EULER) solve_body<R, AXIS_Z, NOISE, true>(a, e);
        else solve_body<R, AXIS_Z, NOISE, false>(a, e);
    }
}
template <typename R, bool AXIS_Z>
static void solve_noise(const StgSttSolveArgs& a) {
    if (a.flags & STG_F_THERMAL_INJECT) solve_all<R, AXIS_Z, 2>(a);
    else if (a.flags & STG_F_THERMAL_PHILOX) solve_all<R, AXIS_Z, 1>(a);
    else solve_all<R, AXIS_Z, 0>(a);
}
template <int NOISE>
static void solve_grid_all(const StgSttSolveArgs& a) {      // mirrors launch_solve_grid in stt_kernels.cu
    for (int64_t e = 0; e < a.n_envs; ++e) {
        if (a.flags & STG_F_EULER) solve_grid_body<NOISE, true>(a, e);
        else solve_grid_body<NOISE, false>(a, e);
    }
}
extern "C" int hostsim_stt_solve(const StgSttSolveArgs* a, int f64) {
    if (a->d_current_grid || a->d_field_grid) {             // host-sampled callables: FP64 general stages
        if (a->flags & STG_F_THERMAL_INJECT) solve_grid_all<2>(*a);
        else if (a->flags & STG_F_THERMAL_PHILOX) solve_grid_all<1>(*a);
        else solve_grid_all<0>(*a);
        return 0;
    }
    FtzScope ftz(!f64);
    const bool z = (a->flags & STG_F_AXIS_Z) != 0;
    if (f64) { if (z) solve_noise<double, true>(*a); else solve_noise<double, false>(*a); }
    else     { if (z) solve_noise<float, true>(*a);  else solve_noise<double, false>(*a); }
    return 0;
}

extern "C" void hostsim_sort_bins(const StgSttFolded* table, const int32_t* pidx, const float* action, int32_t* bins,
                                  int64_t n) {
    for (int64_t e = 0; e < n; ++e) bins[e] = action_bin(table, pidx, action, e);
}

// Philox known-answer access for tests
extern "C" void hostsim_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out) {
    Philox ph{k0, k1};
    ph(c0, c1, c2, c3, out);
}
// the 12 samples of substep `sub` of an env-step's thermal stream (sequential: the draws before it are made and dropped)
template <int GEN>
static void normals12(uint64_t seed, uint64_t gid, uint32_t episode, uint32_t step, uint32_t sub, float* out) {
    const float unit = -1.3862943611198906f;
    const NoiseStream ns = make_stream(seed, gid, episode, step);
    ThermalSource<float, GEN> src;
    src.init(&ns, unit);
    for (uint32_t i = 0; i <= sub; ++i) {
        if (i & 1) src.second(i >> 1, out); else src.first(i >> 1, out);
    }
}
extern "C" void hostsim_normals12(uint64_t seed, uint64_t gid, uint32_t episode, uint32_t step, uint32_t sub, float* out) {
    normals12<0>(seed, gid, episode, step, sub, out);
}
// the same for the all-Philox stream (STG_F_STREAM_PHILOX10)
extern "C" void hostsim_normals12_philox(uint64_t seed, uint64_t gid, uint32_t episode, uint32_t step, uint32_t sub, float* out) {
    normals12<1>(seed, gid, episode, step, sub, out);
}
// xoshiro128++ known-answer access for tests: n outputs from the given state
extern "C" void hostsim_xoshiro(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, int n, uint32_t* out) {
    Xoshiro128pp g{s0, s1, s2, s3};
    for (int i = 0; i < n; ++i) out[i] = g.next();
}
// the seed state of an env-step's stream
extern "C" void hostsim_stream_seed(uint64_t seed, uint64_t gid, uint32_t episode, uint32_t step, uint32_t* out) {
    const Xoshiro128pp g = seed_xoshiro(make_stream(seed, gid, episode, step));
    out[0] = g.s0; out[1] = g.s1; out[2] = g.s2; out[3] = g.s3;
}

extern "C" int hostsim_llgs_rk45(const StgRk45Args* a) {
    for (int64_t e = 0; e < a->n_envs; ++e) {
        if (a->n_seg > 0) rk45_body<true>(*a, e); else rk45_body<false>(*a, e);
    }
    return 0;
}

// One step of n arrays through the K3 helpers (the CUDA kernel runs the same sequence with thread 0 of each CTA).
extern "C" int hostsim_array_step(const StgArrayStepArgs* pa) {
    const StgArrayStepArgs& a = *pa;
    const StgArrayParams& p = a.params;
    const int nd = p.n_rows * p.n_cols;
    std::vector<double> scratch(nd);
    for (int64_t arr = 0; arr < a.n_arrays; ++arr) {
        double* pattern = a.d_pattern + arr * nd * 3;
        const double* target = a.d_target + arr * nd * 3;
        const double prev = array_similarity(pattern, target, nd, scratch.data());
        const ArrayAction act = array_parse_action(p, a.d_action + arr * a.action_stride);
        const double energy = array_apply_action(p, a.d_coupling, pattern, act);
        const double sim = array_similarity(pattern, target, nd, scratch.data());
        const bool success = sim >= p.success_threshold;
        const double sd = array_magnitude_std(pattern, nd, scratch.data());
        a.d_reward[arr] = array_reward(p, success, sim, energy, dadd(sim, -prev), sd);
        a.d_step_count[arr] += 1;
        a.d_total_energy[arr] = dadd(a.d_total_energy[arr], energy);
        a.d_terminated[arr] = success;
        a.d_truncated[arr] = a.d_step_count[arr] >= p.max_steps;
        if (a.d_step_energy) a.d_step_energy[arr] = energy;
        if (a.d_similarity) a.d_similarity[arr] = sim;
        for (int q = 0; q < nd * 6; ++q)
            a.d_obs[arr * nd * 6 + q] = (float)((q % 6) < 3 ? pattern[3 * (q / 6) + q % 6] : target[3 * (q / 6) + q % 6 - 3]);
    }
    return 0;
}

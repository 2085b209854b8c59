"""SpinTorqueVectorEnv — N SpinTorque-v0 environments stepped by ONE kernel launch on a B200.

Host-side mirror of the reference's SpinTorqueEnv (spin_torque_gym/envs/spin_torque_env.py:26-554) with the same
constructor kwargs, spaces, reset options and step outputs, batched over `num_envs` and backed by the C-ABI in
include/stg.h (libstg.so). State, actions, observations, rewards and flags are torch tensors resident in HBM; numpy in /
numpy out is supported for drop-in use (pinned staging buffers).

There is no CPU fallback: constructing the env without CUDA or without libstg.so raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np

from .. import _lib, params as _params
from ..spaces import Box, batch_box


class SpinTorqueVectorEnv:
    """Batched drop-in for `gym.make('SpinTorque-v0', ...)` (gymnasium.vector.VectorEnv-style API).

    Reference kwargs (envs/spin_torque_env.py:36-53) are accepted unchanged. Extra kwargs:
      num_envs, device ('cuda:k'), dtype (torch.float32 | torch.float64: arithmetic of the integrator stages),
      integrator ('rk4' | 'euler'), rng_seed (Philox key), env_offset (global id of env 0: multi-GPU sharding),
      autoreset (same-step reset of finished episodes, SB3 VecEnv convention), param_index (per-env index into a list of
      device_params dicts = device mix), sort_by_substeps ('auto' | True | False).
    """

    metadata = {"render_modes": [], "autoreset_mode": "same_step"}

    def __init__(
        self,
        num_envs: int = 1,
        device_type: Union[str, Sequence[str]] = "stt_mram",
        device_params: Optional[Union[Dict[str, Any], Sequence[Dict[str, Any]]]] = None,
        target_states: Optional[List[np.ndarray]] = None,
        max_steps: int = 100,
        max_current: float = 2e6,
        max_duration: float = 5e-9,
        temperature: float = 300.0,
        include_thermal_fluctuations: bool = True,
        reward_components: Optional[Dict[str, Dict]] = None,
        action_mode: str = "continuous",
        observation_mode: str = "vector",
        success_threshold: float = 0.9,
        energy_penalty_weight: float = 0.1,
        render_mode: Optional[str] = None,
        seed: Optional[int] = None,
        *,
        device: Union[str, Any] = "cuda",
        dtype: Any = None,
        integrator: str = "rk4",
        rng_seed: Optional[int] = None,
        env_offset: int = 0,
        autoreset: bool = True,
        param_index: Optional[Any] = None,
        sort_by_substeps: Union[str, bool] = "auto",
        collect_stats: bool = True,
        pair_kernel: Any = True,
        host_outputs: bool = False,
        thermal_stream: str = "xoshiro",
    ):
        torch = _lib.require_cuda()
        self._torch = torch
        self._lib = _lib.load()
        if action_mode != "continuous":
            # the reference's discrete mode raises inside step() and returns the error tuple; not supported
            raise ValueError("only action_mode='continuous' is implemented (the reference's 'discrete' mode is non-functional)")
        if observation_mode != "vector":
            raise ValueError("only observation_mode='vector' is implemented (the reference's 'dict' mode is non-functional)")
        if reward_components is not None:
            raise ValueError("custom reward_components are Python callables and cannot run in the kernel; "
                             "the default CompositeReward (envs/spin_torque_env.py:184-207) is fused")
        if integrator not in ("rk4", "euler"):
            raise ValueError(f"unknown integrator {integrator!r}")
        self._step_args_cache = None
        self.num_envs = int(num_envs)
        if self.num_envs <= 0:
            raise ValueError("num_envs must be positive")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.StgError("SpinTorqueVectorEnv requires a CUDA device (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = torch.float32 if dtype is None else dtype
        if self.dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be torch.float32 or torch.float64")
        self.integrator = integrator
        self.max_steps = int(max_steps)
        self.max_current = float(max_current)
        self.max_duration = float(max_duration)
        self.temperature = float(temperature)
        self.include_thermal = bool(include_thermal_fluctuations)
        self.action_mode = action_mode
        self.observation_mode = observation_mode
        self.success_threshold = float(success_threshold)
        self.energy_penalty_weight = float(energy_penalty_weight)
        self.render_mode = render_mode
        self.autoreset = bool(autoreset)
        self.env_offset = int(env_offset)
        self.collect_stats = bool(collect_stats)
        # FP32 / e=z / RK4: two envs per thread on packed FFMA2 (same results per env). True: where it is the faster kernel (always
        # without the thermal stream, from 262,144 envs with it); 'always': at every batch size; False: one env per thread
        # in-kernel thermal stream of the RK4 paths. 'xoshiro' (default): one xoshiro128++ state per env-step, seeded from block 0
        # of the env-step's Philox4x32-10 stream; 'philox': every word from Philox4x32-10 (counter-based down to the substep,
        # ~20 % slower). Both are pure functions of (rng_seed, global env id, episode, step).
        if thermal_stream not in ("xoshiro", "philox"):
            raise ValueError("thermal_stream must be 'xoshiro' or 'philox'")
        self.thermal_stream = thermal_stream
        if pair_kernel not in (True, False, "always"):
            raise ValueError("pair_kernel must be True, False or 'always'")
        self.pair_kernel = pair_kernel
        # host_outputs: obs / final_obs / reward / terminated / truncated live in PINNED HOST memory and the kernels write
        # them there directly (posted PCIe writes while the kernel runs) - for consumers on the CPU (SB3, NumPy policies).
        # step() then returns CPU tensors and synchronises the stream before returning. Default: CUDA tensors, no sync.
        self.host_outputs = bool(host_outputs)
        if rng_seed is None:
            # no seed at all: fresh entropy like gymnasium / the reference (np_random(None)); state_dict() records the value
            rng_seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        self.rng_seed = int(rng_seed) & 0xFFFFFFFFFFFFFFFF

        # ---- device parameter sets ---------------------------------------------------------------------------------
        types = [device_type] if isinstance(device_type, str) else list(device_type)
        if device_params is None or isinstance(device_params, dict):
            plist = [device_params] * len(types)
        else:
            plist = list(device_params)
            if len(types) == 1:
                types = types * len(plist)
        if len(types) != len(plist):
            raise ValueError("device_type and device_params lists must have the same length")
        self.device_types = types
        self.device_type = types[0]
        self.device_params_list = []
        structs = []
        for t, p in zip(types, plist):
            if p is None:
                p = _params.env_default_device_params(t)          # envs/spin_torque_env.py:112-114
            _params.check_device_constructible(t, p)
            self.device_params_list.append(dict(p))
            structs.append(_params.make_param_struct(
                t, p, max_steps=self.max_steps, max_current=self.max_current, max_duration=self.max_duration,
                temperature=self.temperature, thermal=self.include_thermal, success_threshold=self.success_threshold,
                energy_penalty_weight=self.energy_penalty_weight))
        self.device_params = self.device_params_list[0]
        self._folded_host = _params.fold(structs)
        self._axis_z = _params.all_axis_z(self._folded_host)
        self._n_sets = len(structs)

        N = self.num_envs
        dev = self.device
        with torch.cuda.device(dev):
            self._table = torch.from_numpy(self._folded_host).to(dev)
            if param_index is not None:
                pi = torch.as_tensor(param_index, dtype=torch.int32).to(dev).contiguous()
                if pi.shape != (N,):
                    raise ValueError("param_index must have shape (num_envs,)")
                if int(pi.min()) < 0 or int(pi.max()) >= self._n_sets:
                    raise ValueError("param_index out of range")
                self._param_index = pi
            else:
                if self._n_sets != 1:
                    raise ValueError("several device parameter sets need a param_index")
                self._param_index = None

            # ---- target states (envs/spin_torque_env.py:117-120) ----------------------------------------------------
            if target_states is None:
                tt = np.array([[0.0, 0.0, 1.0], [0.0, 0.0, -1.0]])
            else:
                tt = np.array([np.asarray(t, dtype=float) / np.linalg.norm(np.asarray(t, dtype=float))
                               for t in target_states])
            self.target_states = [t.copy() for t in tt]
            self._target_table = torch.from_numpy(np.ascontiguousarray(tt)).to(dev)

            # ---- state planes (FP64 SoA) and step outputs -----------------------------------------------------------
            f64, i32 = torch.float64, torch.int32
            self._m = torch.zeros(3, N, dtype=f64, device=dev)
            self._m[2] = 1.0
            self._target = torch.zeros(3, N, dtype=f64, device=dev)
            self._target[2] = 1.0
            self._total_energy = torch.zeros(N, dtype=f64, device=dev)
            self._last_action = torch.zeros(2, N, dtype=f64, device=dev)
            self._step_count = torch.zeros(N, dtype=i32, device=dev)
            self._episode = torch.zeros(N, dtype=i32, device=dev)
            def out_buf(shape, dtype):
                if self.host_outputs:
                    return torch.zeros(shape, dtype=dtype).pin_memory()
                return torch.zeros(shape, dtype=dtype, device=dev)
            self._obs = out_buf((N, _lib.OBS_DIM), torch.float32)
            self._final_obs = out_buf((N, _lib.OBS_DIM), torch.float32)
            self._reward = out_buf((N,), f64)
            self._terminated = out_buf((N,), torch.uint8)
            self._truncated = out_buf((N,), torch.uint8)
            self._step_energy = torch.zeros(N, dtype=f64, device=dev)
            self._n_sub = torch.zeros(N, dtype=i32, device=dev)
            self._status = torch.zeros(N, dtype=i32, device=dev)
            self._stats = torch.zeros(_lib.STAT_REPLICAS, _lib.NSTATS, dtype=f64, device=dev)
            self._stats_folded = torch.zeros(_lib.NSTATS, dtype=f64, device=dev)
            self._action_dev = torch.zeros(N, 2, dtype=torch.float32, device=dev)
            self._perm = torch.zeros(N, dtype=i32, device=dev)
            self._sort_work = torch.zeros(_lib.SORT_WORK_INTS, dtype=i32, device=dev)
            # second-pass list of stg_stt_step_f32 (include/stg.h, d_redo): envs the FP32 stages decline are repeated with FP64 stages
            self._redo = torch.zeros(_lib.REDO_HEADER + N, dtype=i32, device=dev)
            # NumPy / list actions are staged through two pinned buffers used in turn: the host may only overwrite a buffer after
            # the asynchronous H2D copy that read it has completed (event recorded behind each copy, waited on before reuse)
            self._action_pinned = [torch.zeros(N, 2, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._action_copied = [None, None]
            self._action_slot = 0

        self._sort_mode = sort_by_substeps
        self._needs_reset = True
        self.gpu_launches = 0     # kernels of libstg launched so far (bench.py reports it)

        # ---- spaces (envs/spin_torque_env.py:209-248) ------------------------------------------------------------------
        self.single_action_space = Box(low=np.array([-self.max_current, 0.0]),
                                       high=np.array([self.max_current, self.max_duration]), dtype=np.float32)
        self.single_observation_space = Box(low=-np.inf, high=np.inf, shape=(_lib.OBS_DIM,), dtype=np.float32)
        self.action_space = batch_box(self.single_action_space, N)
        self.observation_space = batch_box(self.single_observation_space, N)

    # ------------------------------------------------------------------------------------------------------------------
    def _state_struct(self) -> _lib.StgSttState:
        s = _lib.StgSttState()
        s.m = self._m.data_ptr()
        s.target = self._target.data_ptr()
        s.total_energy = self._total_energy.data_ptr()
        s.last_action = self._last_action.data_ptr()
        s.step_count = self._step_count.data_ptr()
        s.episode = self._episode.data_ptr()
        return s

    def _stream(self) -> int:
        return self._torch.cuda.current_stream(self.device).cuda_stream

    def _step_args(self) -> _lib.StgSttStepArgs:
        """Argument block of stg_stt_step_*. Every buffer is allocated once in the constructor (load_state_dict copies in
        place), so the block is built once; step() rewrites only the per-call fields (action, noise, permutation, seed, flags)."""
        a = self._step_args_cache
        if a is not None:
            return a
        torch = self._torch
        a = _lib.StgSttStepArgs()
        a.d_table = self._table.data_ptr()
        a.d_param_index = _lib.ptr(self._param_index)
        a.state = self._state_struct()
        o = a.out
        o.obs = _lib.ptr(self._obs)
        o.reward = _lib.ptr(self._reward)
        o.terminated = _lib.ptr(self._terminated)
        o.truncated = _lib.ptr(self._truncated)
        o.step_energy = self._step_energy.data_ptr()
        o.n_sub = self._n_sub.data_ptr()
        o.status = self._status.data_ptr()
        o.final_obs = _lib.ptr(self._final_obs) if self.autoreset else None
        o.stats = self._stats.data_ptr() if self.collect_stats else None
        a.d_redo = self._redo.data_ptr()
        a.d_target_table = self._target_table.data_ptr()
        a.n_targets = self._target_table.shape[0]
        a.env_offset = self.env_offset
        a.n_envs = self.num_envs
        a.n_sets = self._n_sets
        self._step_fn = self._lib.stg_stt_step_f32 if self.dtype == torch.float32 else self._lib.stg_stt_step_f64
        # zero-copy bool views of the u8 flag buffers (a .bool() per step is an extra kernel launch each)
        self._terminated_b = self._terminated.view(torch.bool)
        self._truncated_b = self._truncated.view(torch.bool)
        self._step_args_cache = a
        return a

    def _as_device_f64(self, x, shape):
        torch = self._torch
        t = torch.as_tensor(x, dtype=torch.float64) if not isinstance(x, torch.Tensor) else x.to(torch.float64)
        t = t.to(self.device)
        if t.dim() == len(shape) - 1:
            t = t.unsqueeze(0).expand(*shape)
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {shape} (or one row), got {tuple(t.shape)}")
        return t.contiguous()

    # ------------------------------------------------------------------------------------------------------------------
    def reset(self, *, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None, mask=None):
        """Reset all envs (or the envs selected by `mask`). options: 'initial_state', 'target_state' ([3] or [N,3]),
        'temperature' (accepted and ignored like the reference, which only forwards it to an unused thermal model:
        envs/spin_torque_env.py:301-303)."""
        torch = self._torch
        options = options or {}
        if seed is not None:
            if mask is not None:
                # one Philox key serves every env: re-keying under a mask would move the running envs onto another noise stream
                raise ValueError("reset(seed=..., mask=...) is not supported: reseed with a full reset")
            self.rng_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
            self._episode.zero_()
        N = self.num_envs
        a = _lib.StgSttResetArgs()
        a.d_table = self._table.data_ptr()
        a.d_param_index = _lib.ptr(self._param_index)
        a.state = self._state_struct()
        keep = []
        if mask is not None:
            mk = torch.as_tensor(mask).to(self.device).to(torch.uint8).contiguous()
            keep.append(mk)
            a.d_mask = mk.data_ptr()
        if "initial_state" in options:
            m0 = self._as_device_f64(options["initial_state"], (N, 3))
            keep.append(m0)
            a.d_m0 = m0.data_ptr()
        if "target_state" in options:
            t0 = self._as_device_f64(options["target_state"], (N, 3))
            keep.append(t0)
            a.d_target0 = t0.data_ptr()
        a.d_target_table = self._target_table.data_ptr()
        a.n_targets = self._target_table.shape[0]
        a.d_obs = _lib.ptr(self._obs)
        a.seed = self.rng_seed
        a.env_offset = self.env_offset
        a.n_envs = N
        a.n_sets = self._n_sets
        with torch.cuda.device(self.device):
            _lib.check(self._lib.stg_stt_reset(C.byref(a), self._stream()), "stg_stt_reset")
        self.gpu_launches += 1
        self._needs_reset = False
        self._keep = keep
        self._sync_host_outputs()
        return self._obs, {}

    # ------------------------------------------------------------------------------------------------------------------
    def _stage_actions(self, actions):
        torch = self._torch
        N = self.num_envs
        if isinstance(actions, torch.Tensor):
            if actions.device == self.device and actions.dtype == torch.float32 and actions.is_contiguous() \
                    and tuple(actions.shape) == (N, 2):
                return actions
            if actions.device.type == "cpu" and actions.dtype == torch.float32 and actions.is_pinned() \
                    and tuple(actions.shape) == (N, 2):
                # pinned host tensor: one async H2D, no staging copy. The CALLER owns that buffer: it must not be overwritten
                # before the copy has run (synchronise the stream, or hand over a different buffer for the next step).
                self._action_dev.copy_(actions, non_blocking=True)
                return self._action_dev
            act = actions.to(device=self.device, dtype=torch.float32).reshape(N, 2)
            self._action_dev.copy_(act)
            return self._action_dev
        arr = np.asarray(actions, dtype=np.float32).reshape(N, 2)
        k = self._action_slot
        self._action_slot = 1 - k
        if self._action_copied[k] is not None:
            self._action_copied[k].synchronize()        # the copy issued two steps ago out of this buffer has finished
        self._action_pinned[k].numpy()[...] = arr
        self._action_dev.copy_(self._action_pinned[k], non_blocking=True)
        if not torch.cuda.is_current_stream_capturing():
            ev = self._action_copied[k] or torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._action_copied[k] = ev
        return self._action_dev

    def step(self, actions, noise=None):
        """One env step for all envs. `actions`: [N,2] (J [A/m^2], T [s]) torch (device or host) or numpy.
        `noise`: optional [N, n_max, S, 3] float64 N(0,1) tensor (S=4 rk4 / 1 euler) replacing the in-kernel Philox stream.
        Returns (obs [N,12] f32, reward [N] f64, terminated [N] bool, truncated [N] bool, info dict of tensors)."""
        torch = self._torch
        if self._needs_reset:
            raise RuntimeError("Environment must be reset before calling step")
        N = self.num_envs
        act = self._stage_actions(actions)
        flags = 0
        if self.integrator == "euler":
            flags |= _lib.F_EULER
        if self._axis_z:
            flags |= _lib.F_AXIS_Z
        if self.autoreset:
            flags |= _lib.F_AUTORESET
        if not self.pair_kernel:
            flags |= _lib.F_NO_PAIR
        elif self.pair_kernel == "always":
            flags |= _lib.F_PAIR_ALWAYS
        a = self._step_args()
        a.d_noise, a.noise_stride, a.d_perm = None, 0, None
        if noise is not None:
            nz = torch.as_tensor(noise, dtype=torch.float64).to(self.device).contiguous()
            S = 1 if self.integrator == "euler" else 4
            if nz.dim() != 4 or nz.shape[0] != N or nz.shape[2] != S or nz.shape[3] != 3:
                raise ValueError(f"noise must have shape [N, n_max, {S}, 3]")
            self._noise_keep = nz
            flags |= _lib.F_THERMAL_INJECT
            a.d_noise = nz.data_ptr()
            a.noise_stride = nz.shape[1]
        elif self.include_thermal and self.temperature > 0:
            flags |= _lib.F_THERMAL_PHILOX
            if self.thermal_stream == "philox":
                flags |= _lib.F_STREAM_PHILOX10
        stream = self._stream()
        with _lib.device_guard(torch, self.device):
            # 'auto': the counting sort costs three tiny launches; ragged pulse durations run ~2x faster sorted
            do_sort = self._sort_mode is True or (self._sort_mode == "auto" and N >= 4096)
            if do_sort:
                _lib.check(self._lib.stg_stt_sort_by_substeps(
                    a.d_table, self._n_sets, a.d_param_index, act.data_ptr(),
                    self._perm.data_ptr(), self._sort_work.data_ptr(), N, stream), "stg_stt_sort_by_substeps")
                self.gpu_launches += 3
                flags |= _lib.F_SORTED
                a.d_perm = self._perm.data_ptr()
            a.d_action = act.data_ptr()
            a.seed = self.rng_seed
            a.flags = flags
            _lib.check(self._step_fn(C.byref(a), stream), "stg_stt_step")
        # FP32 stages without the Philox stream are followed by the compacted FP64 pass over the declined envs (d_redo)
        two_pass = (self.dtype == torch.float32 and self._axis_z and self.integrator == "rk4"
                    and not (flags & _lib.F_THERMAL_PHILOX))
        self.gpu_launches += 2 if two_pass else 1
        info = {
            "step_energy": self._step_energy, "n_sub": self._n_sub, "status": self._status,
            "total_energy": self._total_energy, "step_count": self._step_count,
        }
        if self.autoreset:
            info["final_observation"] = self._final_obs
        self._sync_host_outputs()
        return self._obs, self._reward, self._terminated_b, self._truncated_b, info

    def _sync_host_outputs(self) -> None:
        """host_outputs: the kernel wrote obs / reward / flags into pinned host memory; they may be read once the stream has
        drained (not during CUDA-graph capture, where the caller synchronises after replay)."""
        if self.host_outputs and not self._torch.cuda.is_current_stream_capturing():
            self._torch.cuda.current_stream(self.device).synchronize()

    # ------------------------------------------------------------------------------------------------------------------
    @property
    def magnetization(self):
        """[N,3] float64 view-copy of the current magnetisation."""
        return self._m.t().contiguous()

    @property
    def target(self):
        return self._target.t().contiguous()

    def _fold_stats(self):
        """Column sums of the replicated statistics buffer (include/stg.h, STG_STAT_REPLICAS) into one [NSTATS] vector."""
        torch = self._torch
        with _lib.device_guard(torch, self.device):
            _lib.check(self._lib.stg_stats_fold_f64(self._stats.data_ptr(), self._stats_folded.data_ptr(), 0,
                                                    torch.cuda.current_stream(self.device).cuda_stream), "stg_stats_fold_f64")
        return self._stats_folded

    def episode_stats(self, reset: bool = False) -> Dict[str, float]:
        """Accumulated episode statistics of this rank (dict of python floats; one D2H of 64 bytes)."""
        vals = self._fold_stats().cpu().tolist()
        if reset:
            self._stats.zero_()
        return dict(zip(_lib.STAT_NAMES, vals))

    def stats_tensor(self):
        """[NSTATS] float64 CUDA tensor of the statistics accumulated so far: the input of the one all-reduce per rollout.
        It is a folded copy (overwritten by the next call); use reset_stats() to start a new accumulation."""
        return self._fold_stats()

    def reset_stats(self) -> None:
        self._stats.zero_()

    def state_dict(self) -> Dict[str, Any]:
        return {"m": self._m.clone(), "target": self._target.clone(), "total_energy": self._total_energy.clone(),
                "last_action": self._last_action.clone(), "step_count": self._step_count.clone(),
                "episode": self._episode.clone(), "rng_seed": self.rng_seed, "stats": self._fold_stats().clone()}

    def load_state_dict(self, sd: Dict[str, Any]) -> None:
        for k, t in (("m", self._m), ("target", self._target), ("total_energy", self._total_energy),
                     ("last_action", self._last_action), ("step_count", self._step_count),
                     ("episode", self._episode)):
            t.copy_(sd[k])
        self._stats.zero_()
        self._stats[0].copy_(sd["stats"])               # the folded vector goes into the first copy
        self.rng_seed = int(sd["rng_seed"])
        self._needs_reset = False

    def capture_step(self, actions):
        """Capture one `step(actions)` in a CUDA graph and return a `GraphedStep`; `replay()` re-issues the captured launches
        (counting sort + step kernel) for ~one driver call instead of the Python/ctypes launch path (19 us per step). Measured
        (profiles/README.md): device time is unchanged at 1,024 envs - a step is >= ~100 dependent substeps, which already hides
        the launch path - and 10 % lower at 16,384 envs; the gain is a free host thread. `actions` must be the device tensor [N,2] f32 that step() would use without staging: the caller
        overwrites it in place before every replay, and the returned (obs, reward, terminated, truncated, info) are the env's
        persistent output tensors. Thermal noise stays fresh across replays because the Philox counters are built from the
        step / episode counters in device memory. Captured by value: rng_seed, flags and buffer addresses - capture again
        after reset(seed=...)."""
        torch = self._torch
        if not (isinstance(actions, torch.Tensor) and actions.device == self.device and actions.dtype == torch.float32
                and actions.is_contiguous() and tuple(actions.shape) == (self.num_envs, 2)):
            raise ValueError("capture_step needs a contiguous float32 CUDA tensor [num_envs, 2] on the env's device")
        if self._needs_reset:
            raise RuntimeError("Environment must be reset before calling step")
        launches0 = self.gpu_launches
        saved = self.state_dict()
        self.step(actions)                          # outside the capture: CUDA loads kernels lazily on their first launch
        self.load_state_dict(saved)
        graph = torch.cuda.CUDAGraph()
        with _lib.device_guard(torch, self.device):
            with torch.cuda.graph(graph):
                out = self.step(actions)
        per_replay = (self.gpu_launches - launches0) // 2    # warm-up + capture issued the same launches
        self.gpu_launches = launches0               # capture records the launches, it does not execute them
        return GraphedStep(self, graph, out, per_replay)

    def close(self):
        pass


class GraphedStep:
    """A captured `SpinTorqueVectorEnv.step` (see `capture_step`)."""

    def __init__(self, env, graph, outputs, launches_per_replay: int):
        self.env, self.graph, self.outputs, self.launches_per_replay = env, graph, outputs, launches_per_replay

    def replay(self):
        self.graph.replay()
        self.env.gpu_launches += self.launches_per_replay
        return self.outputs

// stt_env_core.cuh — per-env body of the SpinTorque-v0 step (action parse -> integrate -> energy/reward/obs -> auto-reset).
// Shared by the CUDA kernels (stt_kernels.cu) and by the host build used for CPU-side arithmetic checks (tests/hostsim).
#pragma once

#include "../../include/stg.h"
#include "llgs_core.cuh"

namespace stg {

constexpr int kObs = 12;
#ifndef STG_RESYNC_MASK
#define STG_RESYNC_MASK 15   // exact FP64 renormalisation of the master every 16 substeps (fast path)
#endif

// ---- action -> (J, T) ------------------------------------------------------------------------------------------------
// utils/monitoring.py:288-315 (float32 clips, NaN/Inf -> (0, 1e-12)) then envs/spin_torque_env.py:409-433 (FP64 clips)
STG_HD void parse_action(float a0, float a1, double max_current, double max_duration, double& J,
                                             double& T) {
    // np.clip keeps NaN (and clips +-Inf), so only NaN reaches the reference's NaN/Inf test; fminf/fmaxf would drop it
    const bool bad = (a0 != a0) || (a1 != a1);
    a0 = fminf(fmaxf(a0, -1e8f), 1e8f);
    a1 = fminf(fmaxf(a1, 1e-12f), 1e-6f);
    if (bad) {
        a0 = 0.0f;
        a1 = 1e-12f;
    }
    J = fmin(fmax((double)a0, -max_current), max_current);
    T = fmin(fmax((double)a1, 1e-12), max_duration);
}

// Integrate n substeps of size dt from (mx,my,mz); pulse of density J on while t <= t_pulse (envs/spin_torque_env.py:442-443).
// NOISE: 0 none, 1 Philox, 2 injected tensor [n][S][3]. traj: optional [n+1][3] FP64 rows.
template <typename R, bool AXIS_Z, int NOISE, bool EULER>
STG_HD void integrate(const double* f, double J, double& mx, double& my, double& mz, int n, double dt, double t_pulse,
                      double t_end, const Philox& ph, uint64_t gid, uint32_t step_id, const double* noise_row, double* traj,
                      int& guard) {
    constexpr bool TH = NOISE != 0;
    constexpr bool FAST = sizeof(R) == 4 && AXIS_Z && !EULER;          // substep_fast (llgs_core.cuh)
    constexpr bool SCALED = sizeof(R) == 4 && AXIS_Z && !TH;            // block-scaled transverse pair
    constexpr int NS = EULER ? 3 : 12;
    // substeps whose stage times are certainly inside the pulse run with the constant a; the few around the pulse edge
    // evaluate current_func(t) exactly in FP64 (the k4 stage of the LAST substep sees t_i+dt > T for ~16 % of f32 durations)
    int i_safe;
    if (t_pulse >= t_end) {
        i_safe = n - 1;
    } else {
        const double qd = t_pulse / dt - 2.0;
        i_safe = qd < 0.0 ? 0 : (qd > (double)n ? n : (int)qd);
    }
    if (traj) { traj[0] = mx; traj[1] = my; traj[2] = mz; }

    if constexpr (FAST) {
        StepConsts<float> c;
        make_consts<float>(f, dt, J, 1.0 / 6.0, c);
        const float nscale = -1.3862943611198906f * c.cth * c.cth;      // -2 ln2 * (G h_th / 6)^2
        FastState s;
        s.st = ScaledState{mx, my, mz, 1.0, 1.0, 1.0f};
        if (SCALED) rescale(s.st);
        fast_resync(s);
        // one substep + the periodic exact renormalisation; `edge` substeps evaluate the pulse gate per stage
        auto one = [&](int i, bool edge) {
            float aH1 = c.a_hi, aL1 = c.a_lo, aH2 = c.a_hi, aL2 = c.a_lo, aH4 = c.a_hi, aL4 = c.a_lo;
            if (edge) {
                if (!pulse_on(i, 0, dt, t_pulse)) { aH1 = 0.0f; aL1 = 0.0f; }
                if (!pulse_on(i, 1, dt, t_pulse)) { aH2 = 0.0f; aL2 = 0.0f; }
                if (!pulse_on(i, 2, dt, t_pulse)) { aH4 = 0.0f; aL4 = 0.0f; }
            }
            float nz[12];
            if (NOISE == 1) {
                philox_normals12(ph, gid, step_id, (uint32_t)i, nscale, nz);
            } else if (NOISE == 2) {
#pragma unroll
                for (int q = 0; q < 12; ++q) nz[q] = c.cth * (float)noise_row[(int64_t)i * 12 + q];
            }
            substep_fast<TH, SCALED>(c, s, aH1, aL1, aH2, aL2, aH4, aL4, TH ? nz : nullptr, guard);
            if ((i & STG_RESYNC_MASK) == STG_RESYNC_MASK || traj) {
                guard_normalise<float>(s.st, guard);      // exact FP64 renormalisation of the master
                if (SCALED) rescale(s.st);
                fast_resync(s);
            }
            if (traj) {
                traj[3 * (i + 1) + 0] = s.st.sx * s.st.inv_s; traj[3 * (i + 1) + 1] = s.st.sy * s.st.inv_s;
                traj[3 * (i + 1) + 2] = s.st.z;
            }
        };
        int i = 0;
        for (; i < i_safe; ++i) one(i, false);
        for (; i < n; ++i) one(i, true);
        guard_normalise<float>(s.st, guard);
        mx = s.st.sx * s.st.inv_s; my = s.st.sy * s.st.inv_s; mz = s.st.z;
    } else {
        StepConsts<R> c;
        make_consts<R>(f, dt, J, 1.0, c);
        const float nscale = -1.3862943611198906f * (float)c.cth * (float)c.cth;
        ScaledState st{mx, my, mz, 1.0, 1.0, 1.0f};
        if (SCALED) rescale(st);
        for (int i = 0; i < n; ++i) {
            R aH[3] = {c.a_hi, c.a_hi, c.a_hi}, aL[3] = {c.a_lo, c.a_lo, c.a_lo};
            if (i >= i_safe) {
#pragma unroll
                for (int g = 0; g < 3; ++g)
                    if (!pulse_on(i, g, dt, t_pulse)) { aH[g] = R(0); aL[g] = R(0); }
            }
            R nz[NS];
            if (NOISE == 1) {
                if (EULER) {
                    float z[4];
                    philox_normals4(ph, gid, step_id, (uint32_t)i, 0u, nscale, z);
                    nz[0] = (R)z[0]; nz[1] = (R)z[1]; nz[2] = (R)z[2];
                } else {
                    float z[12];
                    philox_normals12(ph, gid, step_id, (uint32_t)i, nscale, z);
#pragma unroll
                    for (int q = 0; q < 12; ++q) nz[q] = (R)z[q];
                }
            } else if (NOISE == 2) {
#pragma unroll
                for (int q = 0; q < NS; ++q) nz[q] = c.cth * (R)noise_row[(int64_t)i * NS + q];
            }
            substep_ref<R, AXIS_Z, TH, EULER>(c, st, aH, aL, TH ? nz : nullptr, guard);
            if (SCALED && (i & 15) == 15) rescale(st);
            if (traj) {
                traj[3 * (i + 1) + 0] = st.sx * st.inv_s; traj[3 * (i + 1) + 1] = st.sy * st.inv_s;
                traj[3 * (i + 1) + 2] = st.z;
            }
        }
        mx = st.sx * st.inv_s; my = st.sy * st.inv_s; mz = st.z;
    }
}

// observation row (envs/spin_torque_env.py:500-520) + SafetyWrapper.validate_observation (utils/monitoring.py:317-330)
STG_HD void make_obs(const double* f, double mx, double my, double mz, double tx, double ty, double tz,
                                         int step, double total_e, double J, double T, float* o) {
    double r = resistance(f, mx, my, mz);
    double max_steps = f[FI_MAXSTEPS];
    o[0] = (float)mx; o[1] = (float)my; o[2] = (float)mz;
    o[3] = (float)tx; o[4] = (float)ty; o[5] = (float)tz;
    o[6] = (float)(r / f[FI_RP]);
    o[7] = (float)(f[FI_TEMP] / 300.0);
    o[8] = (float)((max_steps - (double)step) / max_steps);
    o[9] = (float)(total_e / 1e-12);
    o[10] = (float)(J / f[FI_MAXCUR]);
    o[11] = (float)(T / f[FI_MAXDUR]);
    bool bad = false;
#pragma unroll
    for (int q = 0; q < kObs; ++q) bad |= !(fabsf(o[q]) <= 3.4028234e38f);
    if (bad) {
#pragma unroll
        for (int q = 0; q < kObs; ++q) {
            float v = o[q];
            o[q] = (v != v) ? 0.0f : (v > 3.4028234e38f ? 1e6f : (v < -3.4028234e38f ? -1e6f : v));
        }
    }
}

// normalise(N(0,1)^3) start state + uniform target pick from the Philox reset stream (envs/spin_torque_env.py:286-299)
STG_HD void draw_reset(const Philox& ph, uint64_t gid, uint32_t episode, const double* target_table,
                                           int n_targets, double& mx, double& my, double& mz, double& tx, double& ty,
                                           double& tz) {
    uint32_t o[4];
    float z0, z1, z2, z3;
    uint32_t attempt = 0;
    do {
        ph((uint32_t)gid, (uint32_t)(gid >> 32), episode, 0xFFFF0000u + attempt, o);
        box_muller(o[0], o[1], z0, z1);
        box_muller(o[2], o[3], z2, z3);
        ++attempt;
    } while (z0 * z0 + z1 * z1 + z2 * z2 < 1e-12f && attempt < 8);
    double n = sqrt((double)z0 * z0 + (double)z1 * z1 + (double)z2 * z2);
    mx = z0 / n; my = z1 / n; mz = z2 / n;
    ph((uint32_t)gid, (uint32_t)(gid >> 32), episode, 0xFFFF8000u, o);
    int k = (int)(((uint64_t)o[0] * (uint64_t)n_targets) >> 32);
    tx = target_table[3 * k + 0]; ty = target_table[3 * k + 1]; tz = target_table[3 * k + 2];
}


// Everything one env-step produces besides the state update (kept in registers until the block-level store phase).
struct EnvStepResult {
    float obs[kObs];
    float final_obs[kObs];
    double reward, energy;
    int n_sub, status;
    bool terminated, truncated, did_reset, valid;
    int step_after;     // step count of the finished step (episode length if it ended)
};

// One env (index e of a.n_envs): reads and updates the FP64 state planes, returns the outputs in `r`.
template <typename R, bool AXIS_Z, int NOISE, bool EULER>
STG_HD void env_step_body(const StgSttStepArgs& a, int64_t e, EnvStepResult& r) {
    const int64_t n = a.n_envs;
    const bool autoreset = (a.flags & STG_F_AUTORESET) != 0;
    const double* f = a.d_table[a.d_param_index ? a.d_param_index[e] : 0].v;
    double mx = a.state.m[e], my = a.state.m[n + e], mz = a.state.m[2 * n + e];
    double tx = a.state.target[e], ty = a.state.target[n + e], tz = a.state.target[2 * n + e];
    double total_e = a.state.total_energy[e];
    int step = a.state.step_count[e];
    const float a0 = a.d_action[2 * e], a1 = a.d_action[2 * e + 1];

    double J, T;
    parse_action(a0, a1, f[FI_MAXCUR], f[FI_MAXDUR], J, T);
    const double prev_align = mx * tx + my * ty + mz * tz;                     // envs/spin_torque_env.py:338-339
    const StepPlan plan = substep_plan(T, f[FI_MAXSTEP_DT]);

    // ---- integrate (utils/robust_solver.py:75 -> physics/simple_solver.py:147-179) ---------------------------------
    double nx = mx, ny = my, nz = mz;
    int guard = 0;
    int status = 0;
    const bool valid = f[FI_VALID] != 0.0 && (!AXIS_Z || f[FI_AXISZ] != 0.0);
    if (valid) {
        guard_normalise<R>(nx, ny, nz, guard);                                 // SimpleLLGSSolver.solve :119
        // Philox stream of this env-step: key = (seed_lo, seed_hi ^ episode), counter = (global env id, step, 4*substep + block)
        // -> every (env, episode, step, substep) draws from its own counter block, independent of the batch partitioning
        Philox ph{(uint32_t)a.seed, (uint32_t)(a.seed >> 32) ^ (uint32_t)a.state.episode[e]};
        const uint64_t gid = a.env_offset + (uint64_t)e;
        const double* nrow = (NOISE == 2) ? a.d_noise + (int64_t)e * a.noise_stride * (EULER ? 3 : 12) : nullptr;
        const uint32_t step_id = (uint32_t)step;
        if (f[FI_HTH] > 0.0 || NOISE == 0) {
            integrate<R, AXIS_Z, NOISE, EULER>(f, J, nx, ny, nz, plan.n, plan.dt, T, T, ph, gid, step_id, nrow, nullptr, guard);
        } else {
            integrate<R, AXIS_Z, 0, EULER>(f, J, nx, ny, nz, plan.n, plan.dt, T, T, ph, gid, step_id, nullptr, nullptr, guard);
        }
        // env-level renormalisation of the last trajectory row (envs/spin_torque_env.py:464)
        const double inv = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
        nx *= inv; ny *= inv; nz *= inv;
        if (guard) {   // A3: a trajectory row failed validation => solver result discarded, m unchanged
            nx = mx; ny = my; nz = mz;
            status |= 1;
        }
    } else {
        status |= 2;
    }

    // ---- Joule energy with the PRE-step m (envs/spin_torque_env.py:474-480) ------------------------------------------
    double energy = 0.0;
    if (fabs(J) > 1e-12) {
        const double res = resistance(f, mx, my, mz);
        const double v = J * res * f[FI_AREA];
        energy = v * v / res * T;
    }
    total_e += energy;
    step += 1;
    const double align = nx * tx + ny * ty + nz * tz;                           // :350-353
    const bool success = align >= f[FI_SUCC];
    // CompositeReward default components in dict order (:184-207), then validate_reward (utils/monitoring.py:332-348)
    double reward = 0.0;
    reward += 10.0 * (success ? 10.0 : 0.0);
    reward += (-f[FI_WE]) * (-energy / 1e-12);
    reward += 1.0 * (align - prev_align);
    if (!(fabs(reward) <= 1.7e308)) reward = -1.0;
    reward = fmin(fmax(reward, -1e6), 1e6);
    const bool truncated = step >= (int)f[FI_MAXSTEPS];                          // :371-372

    make_obs(f, nx, ny, nz, tx, ty, tz, step, total_e, J, T, r.obs);
    r.reward = reward; r.energy = energy; r.n_sub = plan.n; r.status = status;
    r.terminated = success; r.truncated = truncated; r.valid = valid; r.step_after = step;
    r.did_reset = false;

    double lj = J, lt = T;
    if (autoreset && (success || truncated)) {
        // same-call reset (SB3 VecEnv convention): the returned obs is the first obs of the next episode; the last obs of the
        // finished one goes to final_obs
        r.did_reset = true;
        for (int q = 0; q < kObs; ++q) r.final_obs[q] = r.obs[q];
        const uint32_t ep = (uint32_t)a.state.episode[e] + 1u;
        a.state.episode[e] = (int32_t)ep;
        Philox ph{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
        draw_reset(ph, a.env_offset + (uint64_t)e, ep, a.d_target_table, a.n_targets, nx, ny, nz, tx, ty, tz);
        a.state.target[e] = tx; a.state.target[n + e] = ty; a.state.target[2 * n + e] = tz;
        step = 0; total_e = 0.0; lj = 0.0; lt = 0.0;
        make_obs(f, nx, ny, nz, tx, ty, tz, step, total_e, lj, lt, r.obs);
    }
    a.state.m[e] = nx; a.state.m[n + e] = ny; a.state.m[2 * n + e] = nz;
    a.state.total_energy[e] = total_e;
    a.state.step_count[e] = step;
    a.state.last_action[e] = lj;
    a.state.last_action[n + e] = lt;

    a.out.reward[e] = reward;
    a.out.terminated[e] = success ? 1 : 0;
    a.out.truncated[e] = truncated ? 1 : 0;
    if (a.out.step_energy) a.out.step_energy[e] = energy;
    if (a.out.n_sub) a.out.n_sub[e] = plan.n;
    if (a.out.status) a.out.status[e] = status;
}

// SpinTorqueEnv.reset for one env (envs/spin_torque_env.py:250-308)
STG_HD void env_reset_body(const StgSttResetArgs& a, int64_t e) {
    const int64_t n = a.n_envs;
    const double* f = a.d_table[a.d_param_index ? a.d_param_index[e] : 0].v;
    const uint32_t ep = (uint32_t)a.state.episode[e] + 1u;
    a.state.episode[e] = (int32_t)ep;
    double mx, my, mz, tx, ty, tz;
    Philox ph{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
    if (a.d_target_table && a.n_targets > 0) {
        draw_reset(ph, a.env_offset + (uint64_t)e, ep, a.d_target_table, a.n_targets, mx, my, mz, tx, ty, tz);
    } else {
        const double zt[3] = {0.0, 0.0, 1.0};
        draw_reset(ph, a.env_offset + (uint64_t)e, ep, zt, 1, mx, my, mz, tx, ty, tz);
    }
    if (a.d_m0) {   // options['initial_state'] -> device.validate_magnetization (devices/base_device.py:94-116)
        const double x = a.d_m0[3 * e], y = a.d_m0[3 * e + 1], z = a.d_m0[3 * e + 2];
        const double nn = sqrt(x * x + y * y + z * z);
        mx = x / nn; my = y / nn; mz = z / nn;
    }
    if (a.d_target0) {
        const double x = a.d_target0[3 * e], y = a.d_target0[3 * e + 1], z = a.d_target0[3 * e + 2];
        const double nn = sqrt(x * x + y * y + z * z);
        tx = x / nn; ty = y / nn; tz = z / nn;
    }
    a.state.m[e] = mx; a.state.m[n + e] = my; a.state.m[2 * n + e] = mz;
    a.state.target[e] = tx; a.state.target[n + e] = ty; a.state.target[2 * n + e] = tz;
    a.state.total_energy[e] = 0.0;
    a.state.step_count[e] = 0;
    a.state.last_action[e] = 0.0;
    a.state.last_action[n + e] = 0.0;
    if (a.d_obs) make_obs(f, mx, my, mz, tx, ty, tz, 0, 0.0, 0.0, 0.0, a.d_obs + e * kObs);
}

// Batched SimpleLLGSSolver.solve for one env (physics/simple_solver.py:71-191)
template <typename R, bool AXIS_Z, int NOISE, bool EULER>
STG_HD void solve_body(const StgSttSolveArgs& a, int64_t e) {
    const double* f = a.d_table[a.d_param_index ? a.d_param_index[e] : 0].v;
    double mx = a.d_m0[3 * e], my = a.d_m0[3 * e + 1], mz = a.d_m0[3 * e + 2];
    const double J = a.d_pulse[3 * e], t_pulse = a.d_pulse[3 * e + 1], t_end = a.d_pulse[3 * e + 2];
    int guard = 0;
    guard_normalise<R>(mx, my, mz, guard);                       // :119
    int nsub = 0;
    if (t_end > 0.0 && (!AXIS_Z || f[FI_AXISZ] != 0.0)) {        // t_end <= t_start: trivial solution (:122-123)
        StepPlan plan = substep_plan(t_end, f[FI_MAXSTEP_DT]);
        if (a.flags & STG_F_VECTORIZED_PLAN) {   // utils/vectorized_operations.py:55-57
            const double q = ddiv(t_end, f[FI_MAXSTEP_DT]);
            int nv = (q < 2.0e9) ? (int)q : 2000000000;
            if (nv < 10) nv = 10;
            plan.n = nv;
            plan.dt = ddiv(t_end, (double)nv);
        }
        nsub = plan.n;
        Philox ph{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
        const double* nrow = (NOISE == 2) ? a.d_noise + (int64_t)e * a.noise_stride * (EULER ? 3 : 12) : nullptr;
        double* traj = a.d_traj ? a.d_traj + (int64_t)e * a.traj_stride * 3 : nullptr;
        if (f[FI_HTH] > 0.0 || NOISE == 0)
            integrate<R, AXIS_Z, NOISE, EULER>(f, J, mx, my, mz, plan.n, plan.dt, t_pulse, t_end, ph,
                                               a.env_offset + (uint64_t)e, 0u, nrow, traj, guard);
        else
            integrate<R, AXIS_Z, 0, EULER>(f, J, mx, my, mz, plan.n, plan.dt, t_pulse, t_end, ph, a.env_offset + (uint64_t)e,
                                           0u, nullptr, traj, guard);
    }
    a.d_m_out[3 * e] = mx; a.d_m_out[3 * e + 1] = my; a.d_m_out[3 * e + 2] = mz;
    if (a.d_n_sub) a.d_n_sub[e] = nsub;
    if (a.d_guard) a.d_guard[e] = guard;
}

// substep-count bin of an env's action for the counting sort (descending n_sub)
STG_HD int action_bin(const StgSttFolded* table, const int32_t* pidx, const float* action, int64_t e) {
    const double* f = table[pidx ? pidx[e] : 0].v;
    double J, T;
    parse_action(action[2 * e], action[2 * e + 1], f[FI_MAXCUR], f[FI_MAXDUR], J, T);
    const int n = substep_plan(T, f[FI_MAXSTEP_DT]).n;
    return STG_SORT_BINS - 1 - (n < STG_SORT_BINS ? n : STG_SORT_BINS - 1);
}

}  // namespace stg

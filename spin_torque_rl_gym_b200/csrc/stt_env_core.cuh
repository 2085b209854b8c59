// stt_env_core.cuh — per-env body of the SpinTorque-v0 step (action parse -> integrate -> energy/reward/obs -> auto-reset).
// Shared by the CUDA kernels (stt_kernels.cu) and by the host build used for CPU-side arithmetic checks (tests/hostsim).
#pragma once

#include "../../include/stg.h"
#include "llgs_core.cuh"

namespace stg {

#ifndef STG_SUBSTEP_UNROLL
#define STG_SUBSTEP_UNROLL 4      // measured in stt_kernels.cu (STG_MINBLOCKS_NOISE_F32)
#endif
constexpr int kSubstepUnroll = STG_SUBSTEP_UNROLL;
#ifndef STG_SUBSTEP_PAIR_UNROLL
#define STG_SUBSTEP_PAIR_UNROLL 1     // thermal fast path: substep pairs per loop iteration;
                                      // 2 measured 5 % slower (9.70 vs 9.21 ms per 1M-env step, two envs per thread)
#endif
constexpr int kSubstepPairUnroll = STG_SUBSTEP_PAIR_UNROLL;
#ifndef STG_REF_SUBSTEP_UNROLL
#define STG_REF_SUBSTEP_UNROLL 2      // FP64 thermal 262,144-env step: 4.55 -> 4.45 ms; thermal off and tilted axis unchanged
#endif
constexpr int kRefSubstepUnroll = STG_REF_SUBSTEP_UNROLL;   // same for the FP64-stage / general-geometry / Euler loop   // unroll factor of the RK4 substep loop of the FP32 fast path (tuning)

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

// One env of the thermal fast path (FP32 stages, e = z^, RK4, in-kernel noise stream)
struct ThermalEnv {
    const double* f;      // folded parameter set
    double J, dt, t_pulse, t_end;
    int n;
    NoiseStream ns;
};
STG_HD float thermal_nscale(const double* f, double dt) {      // -2 ln2 * (G h_th / 6)^2: field strength folded into Box-Muller
    const float cth = (float)(-f[FI_GEFF] * dt * (1.0 / 6.0) * f[FI_HTH]);
    return -1.3862943611198906f * cth * cth;
}
// substeps below this index are certainly inside the pulse; the few around the pulse edge evaluate current_func(t) exactly in
// FP64 (the k4 stage of the LAST substep sees t_i + dt > T for ~16 % of float32 durations)
STG_HD int pulse_safe_substeps(int n, double dt, double t_pulse, double t_end) {
    if (t_pulse >= t_end) return n - 1;
    const double qd = t_pulse / dt - 2.0;
    return qd < 0.0 ? 0 : (qd > (double)n ? n : (int)qd);
}
// The envs of one thread through the thermal fast path: P = float (one env) or F2 (two envs in the halves of 64-bit register
// pairs on packed FFMA2 / FMUL2 / FADD2). Every operation is an explicit IEEE op per lane and every lane draws from its own
// stream, so an env gets the same bits whatever it is paired with: lanes may have different substep counts - a finished lane is
// frozen (its state is restored after every further substep of its partner). m: [lanes][3] in / out; traj: P = float only.
template <typename P, typename SRC>
STG_HD void integrate_thermal(const ThermalEnv* E, SRC& src, double (*m)[3], int* guard, double* traj = nullptr,
                              int64_t traj_rows = 0x7fffffff) {
    using L = Ln<P>;
    constexpr int NL = L::N;
    ThermalConsts<P> tc;
    P aH, aL, fx, fy, fz, ex, ey, ez;
    int n_max = 0, fast_to = 0x7fffffff, i_safe[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        StepConsts<float> c;
        make_consts<float>(E[l].f, E[l].dt, E[l].J, 1.0 / 6.0, c);
        L::set(tc.c_hi, l, c.c_hi); L::set(tc.c_lo, l, c.c_lo); L::set(tc.al, l, c.al_hi);
        L::set(aH, l, c.a_hi); L::set(aL, l, c.a_lo);
        const float x = (float)m[l][0], y = (float)m[l][1], z = (float)m[l][2];
        L::set(fx, l, x); L::set(fy, l, y); L::set(fz, l, z);
        L::set(ex, l, (float)(m[l][0] - (double)x)); L::set(ey, l, (float)(m[l][1] - (double)y));
        L::set(ez, l, (float)(m[l][2] - (double)z));
        i_safe[l] = pulse_safe_substeps(E[l].n, E[l].dt, E[l].t_pulse, E[l].t_end);
        n_max = E[l].n > n_max ? E[l].n : n_max;
        fast_to = i_safe[l] < fast_to ? i_safe[l] : fast_to;
    }
    if (traj) { traj[0] = m[0][0]; traj[1] = m[0][1]; traj[2] = m[0][2]; }
    // exact FP64 renormalisation of lane l's f + e
    auto renorm = [&](int l) {
        float x = L::get(fx, l), y = L::get(fy, l), z = L::get(fz, l), a = L::get(ex, l), b = L::get(ey, l), c = L::get(ez, l);
        kahan_renorm(x, y, z, a, b, c, guard[l]);
        L::set(fx, l, x); L::set(fy, l, y); L::set(fz, l, z); L::set(ex, l, a); L::set(ey, l, b); L::set(ez, l, c);
    };
    // one substep of every lane; a1/a2/a4: aJ dt seen by stage 1, stages 2 + 3, stage 4 (pulse gate)
    auto substep = [&](int i, const P* nz, P aH1, P aL1, P aH2, P aL2, P aH4, P aL4) {
        P ix, iy, iz, cx, cy, cz, d;
        rk4_thermal<P, false>(tc, fx, fy, fz, aH1, aL1, aH2, aL2, aH4, aL4, nz, ix, iy, iz, cx, cy, cz, d);
        bool small = true;
#pragma unroll
        for (int l = 0; l < NL; ++l) small = small && (fabsf(L::get(d, l)) < 0.015625f);
        if (small) {
            kahan_add<P>(fx, ex, cx);
            kahan_add<P>(fy, ey, cy);
            kahan_add<P>(fz, ez, cz);
        } else {
            // large or non-finite norm change in some lane (diverging parameters): that lane takes the plain increment and the
            // exact FP64 renormalisation with the reference's guard
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                const bool ok = fabsf(L::get(d, l)) < 0.015625f;
                float x = L::get(fx, l), y = L::get(fy, l), z = L::get(fz, l), a = L::get(ex, l), b = L::get(ey, l), c = L::get(ez, l);
                kahan_add<float>(x, a, ok ? L::get(cx, l) : L::get(ix, l));
                kahan_add<float>(y, b, ok ? L::get(cy, l) : L::get(iy, l));
                kahan_add<float>(z, c, ok ? L::get(cz, l) : L::get(iz, l));
                if (!ok) kahan_renorm(x, y, z, a, b, c, guard[l]);
                L::set(fx, l, x); L::set(fy, l, y); L::set(fz, l, z); L::set(ex, l, a); L::set(ey, l, b); L::set(ez, l, c);
            }
        }
        if ((i & STG_RESYNC_MASK) == STG_RESYNC_MASK || traj) {
#pragma unroll
            for (int l = 0; l < NL; ++l) renorm(l);
        }
        if (traj && i + 1 < traj_rows) {
            traj[3 * (i + 1) + 0] = (double)L::get(fx, 0) + (double)L::get(ex, 0);
            traj[3 * (i + 1) + 1] = (double)L::get(fy, 0) + (double)L::get(ey, 0);
            traj[3 * (i + 1) + 2] = (double)L::get(fz, 0) + (double)L::get(ez, 0);
        }
    };
    // a substep around the pulse edge / past the end of the shorter lane: per-lane gates, finished lanes frozen
    auto substep_edge = [&](int i, const P* nz) {
        P aH1 = aH, aL1 = aL, aH2 = aH, aL2 = aL, aH4 = aH, aL4 = aL;
        const P sfx = fx, sfy = fy, sfz = fz, sex = ex, sey = ey, sez = ez;
        int gs[NL];
#pragma unroll
        for (int l = 0; l < NL; ++l) gs[l] = guard[l];
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            if (i >= i_safe[l] && i < E[l].n) {
                if (!pulse_on(i, 0, E[l].dt, E[l].t_pulse)) { L::set(aH1, l, 0.0f); L::set(aL1, l, 0.0f); }
                if (!pulse_on(i, 1, E[l].dt, E[l].t_pulse)) { L::set(aH2, l, 0.0f); L::set(aL2, l, 0.0f); }
                if (!pulse_on(i, 2, E[l].dt, E[l].t_pulse)) { L::set(aH4, l, 0.0f); L::set(aL4, l, 0.0f); }
            }
        }
        substep(i, nz, aH1, aL1, aH2, aL2, aH4, aL4);
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            if (i >= E[l].n) {      // this lane had finished: undo
                L::set(fx, l, L::get(sfx, l)); L::set(fy, l, L::get(sfy, l)); L::set(fz, l, L::get(sfz, l));
                L::set(ex, l, L::get(sex, l)); L::set(ey, l, L::get(sey, l)); L::set(ez, l, L::get(sez, l));
                guard[l] = gs[l];
            }
        }
    };
    // the substep pair (2g, 2g+1): two draws of 12 samples per lane from the thermal stream
    int g = 0;
    const int g_fast = fast_to / 2, g_end = (n_max + 1) / 2;
#pragma unroll kSubstepPairUnroll
    for (; g < g_fast; ++g) {       // every lane running and inside its pulse
        P nz[12];
        src.first((uint32_t)g, nz);
        substep(2 * g, nz, aH, aL, aH, aL, aH, aL);
        src.second((uint32_t)g, nz);
        substep(2 * g + 1, nz, aH, aL, aH, aL, aH, aL);
    }
    for (; g < g_end; ++g) {
        P nz[12];
        src.first((uint32_t)g, nz);
        substep_edge(2 * g, nz);
        if (2 * g + 1 < n_max) {
            src.second((uint32_t)g, nz);
            substep_edge(2 * g + 1, nz);
        }
    }
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        renorm(l);
        m[l][0] = (double)L::get(fx, l) + (double)L::get(ex, l);
        m[l][1] = (double)L::get(fy, l) + (double)L::get(ey, l);
        m[l][2] = (double)L::get(fz, l) + (double)L::get(ez, l);
    }
}

// Integrate n substeps of size dt from (mx,my,mz); pulse of density J on while t <= t_pulse (envs/spin_torque_env.py:442-443).
// NOISE: 0 none, 1 in-kernel stream `ns` (Philox-seeded xoshiro128++), 2 injected tensor [n][S][3], 3 in-kernel stream with every
// word from Philox4x32-10 (STG_F_STREAM_PHILOX10; env step only). traj: optional [n+1][3] FP64 rows.
template <typename R, bool AXIS_Z, int NOISE, bool EULER>
STG_HD void integrate(const double* f, double J, double& mx, double& my, double& mz, int n, double dt, double t_pulse,
                      double t_end, const NoiseStream& ns, const double* noise_row, double* traj,
                      int& guard, int64_t noise_rows = 0x7fffffff, int64_t traj_rows = 0x7fffffff, int* illcond = nullptr) {
    // illcond (FP32 fast path without the Philox stream only): set to 1 when the conditioning bound of the trajectory exceeds
    // kCondTol (llgs_core.cuh: CondTrack) - the caller then repeats the env with FP64 stages
    // injected noise: substeps beyond the caller's tensor reuse its last row instead of reading out of bounds
    auto nrow_of = [&](int i) { return (int64_t)(i < noise_rows ? i : noise_rows - 1); };
    constexpr bool TH = NOISE != 0;
    constexpr bool STREAM = NOISE == 1 || NOISE == 3;
    constexpr int GEN = NOISE == 3 ? 1 : 0;
    constexpr bool FAST = sizeof(R) == 4 && AXIS_Z && !EULER;          // rk4_fast / rk4_thermal (llgs_core.cuh)
    constexpr bool SCALED = sizeof(R) == 4 && AXIS_Z && !TH;            // block-scaled transverse pair
    constexpr bool TRACK = FAST && !STREAM;                             // with the in-kernel stream parity is statistical
    constexpr int NS = EULER ? 3 : 12;
    const int i_safe = pulse_safe_substeps(n, dt, t_pulse, t_end);
    if (traj) { traj[0] = mx; traj[1] = my; traj[2] = mz; }

    if constexpr (FAST && STREAM) {
        ThermalEnv E{f, J, dt, t_pulse, t_end, n, ns};
        ThermalSource<float, GEN> src;
        src.init(&ns, thermal_nscale(f, dt));
        double w[1][3] = {{mx, my, mz}};
        integrate_thermal<float>(&E, src, w, &guard, traj, traj_rows);
        mx = w[0][0]; my = w[0][1]; mz = w[0][2];
    } else if constexpr (FAST) {
        StepConsts<float> c;
        make_consts<float>(f, dt, J, 1.0 / 6.0, c);
        PackConsts<float> pc;
        pack_consts<float>(c, c, pc);
        const ThermalConsts<float> tc{c.c_hi, c.c_lo, c.al_hi};
        const float nscale = -1.3862943611198906f * c.cth * c.cth;      // -2 ln2 * (G h_th / 6)^2
        FastState s;
        s.st = ScaledState{mx, my, mz, 1.0, 1.0, 1.0f};
        if (SCALED) rescale(s.st);
        fast_resync(s);
        CondTrack ct{0.0f, 0.0f};
        // one substep + the periodic exact renormalisation; `edge` substeps evaluate the pulse gate per stage
        auto one = [&](int i, const float* nz, bool edge) {
            float aH1 = c.a_hi, aL1 = c.a_lo, aH2 = c.a_hi, aL2 = c.a_lo, aH4 = c.a_hi, aL4 = c.a_lo;
            if (edge) {
                if (!pulse_on(i, 0, dt, t_pulse)) { aH1 = 0.0f; aL1 = 0.0f; }
                if (!pulse_on(i, 1, dt, t_pulse)) { aH2 = 0.0f; aL2 = 0.0f; }
                if (!pulse_on(i, 2, dt, t_pulse)) { aH4 = 0.0f; aL4 = 0.0f; }
            }
            float ix, iy, iz, cx, cy, cz, d;
            if constexpr (TH)
                rk4_thermal<float, true>(tc, s.fx, s.fy, s.fz, aH1, aL1, aH2, aL2, aH4, aL4, nz, ix, iy, iz, cx, cy, cz, d);
            else
                rk4_fast<float, false, SCALED>(pc, s.fx, s.fy, s.fz, -s.q, s.q, aH1, aL1, aH2, aL2, aH4, aL4, nullptr,
                                               ix, iy, iz, cx, cy, cz, d);
            fast_apply(s, ix, iy, iz, cx, cy, cz, d, guard);
            if ((i & STG_RESYNC_MASK) == STG_RESYNC_MASK || traj) {
                if (TRACK)
                    cond_block(ct, pc.c_hi, pc.ac_hi, aH1, s.fz, transverse_of(s.fx, s.fy, (float)s.st.inv_s),
                               traj ? 1.0f : (float)(STG_RESYNC_MASK + 1));
                guard_normalise<float>(s.st, guard);      // exact FP64 renormalisation of the master
                if (SCALED) rescale(s.st);
                fast_resync(s);
            }
            if (traj && i + 1 < traj_rows) {
                traj[3 * (i + 1) + 0] = s.st.sx * s.st.inv_s; traj[3 * (i + 1) + 1] = s.st.sy * s.st.inv_s;
                traj[3 * (i + 1) + 2] = s.st.z;
            }
        };
        int i = 0;
        if constexpr (STREAM) {
            // handled by integrate_thermal above
        } else if constexpr (NOISE == 2) {
            for (; i < n; ++i) {
                float nz[12];
#pragma unroll
                for (int q = 0; q < 12; ++q) nz[q] = c.cth * (float)noise_row[nrow_of(i) * 12 + q];
                one(i, nz, i >= i_safe);
            }
        } else {
#pragma unroll kSubstepUnroll
            for (; i < i_safe; ++i) one(i, nullptr, false);
            for (; i < n; ++i) one(i, nullptr, true);
        }
        if (TRACK && illcond) {
            const float s_end = transverse_of(s.fx, s.fy, (float)s.st.inv_s);
            if (!traj && (n & STG_RESYNC_MASK))
                cond_block(ct, pc.c_hi, pc.ac_hi, t_pulse >= t_end ? c.a_hi : 0.0f, s.fz, s_end, (float)(n & STG_RESYNC_MASK));
            *illcond = cond_exceeded(ct, s_end) ? 1 : 0;
        }
        guard_normalise<float>(s.st, guard);
        mx = s.st.sx * s.st.inv_s; my = s.st.sy * s.st.inv_s; mz = s.st.z;
    } else {
        StepConsts<R> c;
        make_consts<R>(f, dt, J, 1.0, c);
        const float nscale = -1.3862943611198906f * (float)c.cth * (float)c.cth;
        ScaledState st{mx, my, mz, 1.0, 1.0, 1.0f};
        if (SCALED) rescale(st);
        // one reference-structure substep with the noise rotation components nz (already scaled by cth)
        auto one = [&](int i, const R* nz) {
            R aH[3] = {c.a_hi, c.a_hi, c.a_hi}, aL[3] = {c.a_lo, c.a_lo, c.a_lo};
            if (i >= i_safe) {
#pragma unroll
                for (int g = 0; g < 3; ++g)
                    if (!pulse_on(i, g, dt, t_pulse)) { aH[g] = R(0); aL[g] = R(0); }
            }
            substep_ref<R, AXIS_Z, TH, EULER>(c, st, aH, aL, nz, guard);
            if (SCALED && (i & 15) == 15) rescale(st);
            if (traj && i + 1 < traj_rows) {
                traj[3 * (i + 1) + 0] = st.sx * st.inv_s; traj[3 * (i + 1) + 1] = st.sy * st.inv_s;
                traj[3 * (i + 1) + 2] = st.z;
            }
        };
        int i = 0;
        if constexpr (STREAM && !EULER) {
            ThermalSource<float, GEN> src;
            src.init(&ns, nscale);
#pragma unroll 1
            for (; i < n; i += 2) {
                float z[12];
                R nz[12];
                src.first((uint32_t)i >> 1, z);
#pragma unroll
                for (int q = 0; q < 12; ++q) nz[q] = (R)z[q];
                one(i, nz);
                if (i + 1 < n) {
                    src.second((uint32_t)i >> 1, z);
#pragma unroll
                    for (int q = 0; q < 12; ++q) nz[q] = (R)z[q];
                    one(i + 1, nz);
                }
            }
        } else {
#pragma unroll kRefSubstepUnroll
            for (; i < n; ++i) {
                R nz[NS];
                if (STREAM) {       // Euler
                    float z[4];
                    philox_normals3(ns, (uint32_t)i, nscale, z);
                    nz[0] = (R)z[0]; nz[1] = (R)z[1]; nz[2] = (R)z[2];
                } else if (NOISE == 2) {
#pragma unroll
                    for (int q = 0; q < NS; ++q) nz[q] = c.cth * (R)noise_row[nrow_of(i) * NS + q];
                }
                one(i, TH ? nz : nullptr);
            }
        }
        mx = st.sx * st.inv_s; my = st.sy * st.inv_s; mz = st.z;
    }
}

// Two envs per thread through the packed FP32x2 path (FP32 stages, e = z^, RK4, no thermal field). Env A lives in the .x halves,
// env B in the .y halves. The envs may have different substep counts: the thread runs to the longer one with the finished
// env's constants zeroed (its increments are then exactly 0) and its periodic / final renormalisations skipped, so each env's
// result is bit-identical to the one-env-per-thread path whatever its partner is.
struct PairEnv {
    const double* f;
    double J, dt, t_pulse;
    int n;
};
// apply the packed rk4_fast result to ONE env (H = 0: .x halves, 1: .y halves) of a pair: FP64 master + its half of the
// packed FP32 working copy
template <int H>
STG_HD void pair_apply(ScaledState& st, F2& fx, F2& fy, F2& fz, F2& q, F2& nq, const F2& ix, const F2& iy, const F2& iz,
                       const F2& cx, const F2& cy, const F2& cz, const F2& d, int& guard) {
    const float dd = H ? d.y : d.x;
    if (fabsf(dd) < 0.015625f) {
        st.sx += (double)(H ? cx.y : cx.x);
        st.sy += (double)(H ? cy.y : cy.x);
        st.z += (double)(H ? cz.y : cz.x);
    } else {
        st.sx += (double)(H ? ix.y : ix.x);
        st.sy += (double)(H ? iy.y : iy.x);
        st.z += (double)(H ? iz.y : iz.x);
        guard_normalise<float>(st, guard);
        if (H) { q.y = st.inv_s2f; nq.y = -st.inv_s2f; } else { q.x = st.inv_s2f; nq.x = -st.inv_s2f; }
    }
    if (H) { fx.y = (float)st.sx; fy.y = (float)st.sy; fz.y = (float)st.z; }
    else { fx.x = (float)st.sx; fy.x = (float)st.sy; fz.x = (float)st.z; }
}
template <int H, bool SCALED>
STG_HD void pair_renorm(ScaledState& st, F2& fx, F2& fy, F2& fz, F2& q, F2& nq, int& guard) {
    guard_normalise<float>(st, guard);
    if (SCALED) rescale(st);
    if (H) { fx.y = (float)st.sx; fy.y = (float)st.sy; fz.y = (float)st.z; q.y = st.inv_s2f; nq.y = -st.inv_s2f; }
    else { fx.x = (float)st.sx; fy.x = (float)st.sy; fz.x = (float)st.z; q.x = st.inv_s2f; nq.x = -st.inv_s2f; }
}

STG_HD void integrate_pair(const PairEnv& A, const PairEnv& B, double* mA, double* mB, int& guardA, int& guardB, int& illA,
                           int& illB) {
    constexpr bool TH = false;
    constexpr bool TRACK = true;
    CondTrack ctA{0.0f, 0.0f}, ctB{0.0f, 0.0f};
    constexpr bool SCALED = true;
    StepConsts<float> ca, cb;
    make_consts<float>(A.f, A.dt, A.J, 1.0 / 6.0, ca);
    make_consts<float>(B.f, B.dt, B.J, 1.0 / 6.0, cb);
    PackConsts<F2> pc;
    pack_consts<F2>(ca, cb, pc);
    ScaledState sa{mA[0], mA[1], mA[2], 1.0, 1.0, 1.0f}, sb{mB[0], mB[1], mB[2], 1.0, 1.0, 1.0f};
    if (SCALED) { rescale(sa); rescale(sb); }
    // packed FP32 working copy: env A in the .x halves, env B in the .y halves
    F2 fx = mk2((float)sa.sx, (float)sb.sx), fy = mk2((float)sa.sy, (float)sb.sy), fz = mk2((float)sa.z, (float)sb.z);
    F2 q = mk2(sa.inv_s2f, sb.inv_s2f), nq = mk2(-sa.inv_s2f, -sb.inv_s2f);
    const int n_max = A.n > B.n ? A.n : B.n;
    const int edge_from = (A.n < B.n ? A.n : B.n) - 1;     // the env loop always passes t_end == t_pulse: i_safe = n - 1
    F2 aH = mk2(ca.a_hi, cb.a_hi), aL = mk2(ca.a_lo, cb.a_lo);
    auto one = [&](int i, bool edge) {
        F2 aH4 = aH, aL4 = aL;
        bool runA = true, runB = true;
        if (edge) {
            // an env that has finished contributes exactly zero; the last substep of an env gates its k4 stage in FP64
            runA = i < A.n; runB = i < B.n;
            if (!runA) { pc.c_hi.x = pc.c_lo.x = pc.ac_hi.x = pc.ac_lo.x = 0.0f; aH.x = aL.x = 0.0f; }
            if (!runB) { pc.c_hi.y = pc.c_lo.y = pc.ac_hi.y = pc.ac_lo.y = 0.0f; aH.y = aL.y = 0.0f; }
            aH4 = aH; aL4 = aL;
            if (runA && i == A.n - 1 && !pulse_on(i, 2, A.dt, A.t_pulse)) { aH4.x = 0.0f; aL4.x = 0.0f; }
            if (runB && i == B.n - 1 && !pulse_on(i, 2, B.dt, B.t_pulse)) { aH4.y = 0.0f; aL4.y = 0.0f; }
        }
        F2 ix, iy, iz, cx, cy, cz, d;
        rk4_fast<F2, TH, SCALED>(pc, fx, fy, fz, nq, q, aH, aL, aH, aL, aH4, aL4, nullptr, ix, iy, iz, cx, cy, cz, d);
        if (runA) pair_apply<0>(sa, fx, fy, fz, q, nq, ix, iy, iz, cx, cy, cz, d, guardA);
        if (runB) pair_apply<1>(sb, fx, fy, fz, q, nq, ix, iy, iz, cx, cy, cz, d, guardB);
        if ((i & STG_RESYNC_MASK) == STG_RESYNC_MASK) {
            if (TRACK) {
                if (runA) cond_block(ctA, pc.c_hi.x, pc.ac_hi.x, aH.x, fz.x, transverse_of(fx.x, fy.x, (float)sa.inv_s),
                                     (float)(STG_RESYNC_MASK + 1));
                if (runB) cond_block(ctB, pc.c_hi.y, pc.ac_hi.y, aH.y, fz.y, transverse_of(fx.y, fy.y, (float)sb.inv_s),
                                     (float)(STG_RESYNC_MASK + 1));
            }
            if (runA) pair_renorm<0, SCALED>(sa, fx, fy, fz, q, nq, guardA);
            if (runB) pair_renorm<1, SCALED>(sb, fx, fy, fz, q, nq, guardB);
        }
    };
    int i = 0;
    for (; i < edge_from; ++i) one(i, false);
    for (; i < n_max; ++i) one(i, true);
    illA = illB = 0;
    if (TRACK) {
        const float sA = transverse_of((float)sa.sx, (float)sa.sy, (float)sa.inv_s);
        const float sB = transverse_of((float)sb.sx, (float)sb.sy, (float)sb.inv_s);
        if (A.n & STG_RESYNC_MASK) cond_block(ctA, ca.c_hi, ca.ac_hi, ca.a_hi, (float)sa.z, sA, (float)(A.n & STG_RESYNC_MASK));
        if (B.n & STG_RESYNC_MASK) cond_block(ctB, cb.c_hi, cb.ac_hi, cb.a_hi, (float)sb.z, sB, (float)(B.n & STG_RESYNC_MASK));
        illA = cond_exceeded(ctA, sA) ? 1 : 0;
        illB = cond_exceeded(ctB, sB) ? 1 : 0;
    }
    guard_normalise<float>(sa, guardA);
    guard_normalise<float>(sb, guardB);
    mA[0] = sa.sx * sa.inv_s; mA[1] = sa.sy * sa.inv_s; mA[2] = sa.z;
    mB[0] = sb.sx * sb.inv_s; mB[1] = sb.sy * sb.inv_s; mB[2] = sb.z;
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

// What the first half of an env step computes for one env (loads, action parsing, step plan, normalised start state).
struct EnvStepCtx {
    const double* f;
    double mx, my, mz, tx, ty, tz, total_e, J, T, prev_align;
    double w[3];        // working magnetisation handed to / returned by the integrator
    StepPlan plan;
    int step, guard;
    int illcond;        // FP32 stages only: the conditioning bound of the trajectory exceeded kCondTol (llgs_core.cuh)
    bool valid;
};

template <typename R, bool AXIS_Z>
STG_HD void env_step_prologue(const StgSttStepArgs& a, int64_t e, EnvStepCtx& c) {
    const int64_t n = a.n_envs;
    c.f = a.d_table[a.d_param_index ? a.d_param_index[e] : 0].v;
    c.mx = a.state.m[e]; c.my = a.state.m[n + e]; c.mz = a.state.m[2 * n + e];
    c.tx = a.state.target[e]; c.ty = a.state.target[n + e]; c.tz = a.state.target[2 * n + e];
    c.total_e = a.state.total_energy[e];
    c.step = a.state.step_count[e];
    parse_action(a.d_action[2 * e], a.d_action[2 * e + 1], c.f[FI_MAXCUR], c.f[FI_MAXDUR], c.J, c.T);
    c.prev_align = dot3(c.mx, c.my, c.mz, c.tx, c.ty, c.tz);                  // envs/spin_torque_env.py:338-339
    c.plan = substep_plan(c.T, c.f[FI_MAXSTEP_DT]);
    c.w[0] = c.mx; c.w[1] = c.my; c.w[2] = c.mz;
    c.guard = 0;
    c.illcond = 0;
    c.valid = c.f[FI_VALID] != 0.0 && (!AXIS_Z || c.f[FI_AXISZ] != 0.0);
    if (c.valid) guard_normalise<R>(c.w[0], c.w[1], c.w[2], c.guard);         // SimpleLLGSSolver.solve :119
}

// Second half: env-level renormalisation, energy, reward, flags, observation, auto-reset, state / output stores.
STG_HD void env_step_epilogue(const StgSttStepArgs& a, int64_t e, EnvStepCtx& c, EnvStepResult& r, int status_bits = 0) {
    const int64_t n = a.n_envs;
    const bool autoreset = (a.flags & STG_F_AUTORESET) != 0;
    const double* f = c.f;
    const double mx = c.mx, my = c.my, mz = c.mz;
    double tx = c.tx, ty = c.ty, tz = c.tz;
    double nx = mx, ny = my, nz = mz;
    int status = status_bits;
    if (c.valid) {
        // env-level renormalisation of the last trajectory row (envs/spin_torque_env.py:464)
        const double inv = 1.0 / sqrt(dot3(c.w[0], c.w[1], c.w[2], c.w[0], c.w[1], c.w[2]));
        nx = c.w[0] * inv; ny = c.w[1] * inv; nz = c.w[2] * inv;
        if (c.guard) {   // A3: a trajectory row failed validation => solver result discarded, m unchanged
            nx = mx; ny = my; nz = mz;
            status |= 1;
        }
    } else {
        status |= 2;
    }
    const double J = c.J, T = c.T;
    // ---- Joule energy with the PRE-step m (envs/spin_torque_env.py:474-480) ------------------------------------------
    double energy = 0.0;
    if (fabs(J) > 1e-12) {
        const double res = resistance(f, mx, my, mz);
        const double v = J * res * f[FI_AREA];
        energy = v * v / res * T;
    }
    double total_e = c.total_e + energy;
    int step = c.step + 1;
    const double align = dot3(nx, ny, nz, tx, ty, tz);                          // :350-353
    const bool success = align >= f[FI_SUCC];
    // CompositeReward default components in dict order (:184-207), then validate_reward (utils/monitoring.py:332-348)
    // (explicit products and sums: see dot3 in llgs_core.cuh)
    double reward = success ? 100.0 : 0.0;                                       // 10.0 * (10.0 if success else 0.0)
    reward = dadd(reward, dmul(-f[FI_WE], -energy / 1e-12));
    reward = dadd(reward, dadd(align, -c.prev_align));
    if (!(fabs(reward) <= 1.7e308)) reward = -1.0;
    reward = fmin(fmax(reward, -1e6), 1e6);
    const bool truncated = step >= (int)f[FI_MAXSTEPS];                          // :371-372

    make_obs(f, nx, ny, nz, tx, ty, tz, step, total_e, J, T, r.obs);
    r.reward = reward; r.energy = energy; r.n_sub = c.plan.n; r.status = status;
    r.terminated = success; r.truncated = truncated; r.valid = c.valid; r.step_after = step;
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
    if (a.out.n_sub) a.out.n_sub[e] = c.plan.n;
    if (a.out.status) a.out.status[e] = status;
}

// Thermal-field stream of one env-step (llgs_core.cuh NoiseStream): key = seed, counter = (global env id, episode, step, block)
// -> every (env, episode, step, substep) draws from its own counter blocks, independent of the batch partitioning
template <typename R, bool AXIS_Z, int NOISE, bool EULER>
STG_HD void env_step_integrate(const StgSttStepArgs& a, int64_t e, EnvStepCtx& c) {
    const NoiseStream ns = make_stream(a.seed, a.env_offset + (uint64_t)e, (uint32_t)a.state.episode[e], (uint32_t)c.step);
    const double* nrow = (NOISE == 2) ? a.d_noise + (int64_t)e * a.noise_stride * (EULER ? 3 : 12) : nullptr;
    if (c.f[FI_HTH] > 0.0 || NOISE == 0)
        integrate<R, AXIS_Z, NOISE, EULER>(c.f, c.J, c.w[0], c.w[1], c.w[2], c.plan.n, c.plan.dt, c.T, c.T, ns,
                                           nrow, nullptr, c.guard, NOISE == 2 ? a.noise_stride : 0x7fffffff, 0x7fffffff,
                                           &c.illcond);
    else {
        // a parameter set without thermal field inside a launch with the noise stream (device mixes): deterministic path; such a
        // launch has no second pass, so an ill-conditioned FP32 trajectory is repeated with FP64 stages on the spot
        const double w0 = c.w[0], w1 = c.w[1], w2 = c.w[2];
        const int guard0 = c.guard;
        integrate<R, AXIS_Z, 0, EULER>(c.f, c.J, c.w[0], c.w[1], c.w[2], c.plan.n, c.plan.dt, c.T, c.T, ns,
                                       nullptr, nullptr, c.guard, 0x7fffffff, 0x7fffffff, &c.illcond);
        if (sizeof(R) == 4 && c.illcond) {
            c.w[0] = w0; c.w[1] = w1; c.w[2] = w2; c.guard = guard0; c.illcond = 0;
            integrate<double, AXIS_Z, 0, EULER>(c.f, c.J, c.w[0], c.w[1], c.w[2], c.plan.n, c.plan.dt, c.T, c.T, ns,
                                                nullptr, nullptr, c.guard);
        }
    }
}

// One env (index e of a.n_envs): reads and updates the FP64 state planes, returns the outputs in `r`.
// Returns false - with NOTHING written, state and outputs untouched - when the FP32 stages cannot guarantee the 1e-4 contract
// for this trajectory (CondTrack): the caller repeats the env through the <double> instantiation (the kernels collect such
// envs in StgSttStepArgs.d_redo and run them in a second, compacted launch; status bit 2 marks them).
template <typename R, bool AXIS_Z, int NOISE, bool EULER>
STG_HD bool env_step_body(const StgSttStepArgs& a, int64_t e, EnvStepResult& r, int status_bits = 0) {
    EnvStepCtx c;
    env_step_prologue<R, AXIS_Z>(a, e, c);
    if (c.valid) env_step_integrate<R, AXIS_Z, NOISE, EULER>(a, e, c);
    if (sizeof(R) == 4 && c.illcond) return false;
    env_step_epilogue(a, e, c, r, status_bits);
    return true;
}

// Two envs (eA, eB) through the packed FP32x2 integrators (R = float, e = z^, RK4; NOISE 0: integrate_pair, 1: integrate_thermal
// with the in-kernel stream). Returns a mask of the envs that were NOT stepped because their FP32 trajectory is ill-conditioned
// (bit 0: eA, bit 1: eB; see env_step_body; never set with the noise stream, where parity is statistical).
template <int NOISE>
STG_HD int env_step_pair_body(const StgSttStepArgs& a, int64_t eA, int64_t eB, EnvStepResult& rA, EnvStepResult& rB) {
    EnvStepCtx ca, cb;
    env_step_prologue<float, true>(a, eA, ca);
    env_step_prologue<float, true>(a, eB, cb);
    const bool noise_ok = NOISE == 0 || (ca.f[FI_HTH] > 0.0 && cb.f[FI_HTH] > 0.0);
    if (ca.valid && cb.valid && noise_ok) {
        if constexpr (NOISE == 0) {
            PairEnv A{ca.f, ca.J, ca.plan.dt, ca.T, ca.plan.n};
            PairEnv B{cb.f, cb.J, cb.plan.dt, cb.T, cb.plan.n};
            integrate_pair(A, B, ca.w, cb.w, ca.guard, cb.guard, ca.illcond, cb.illcond);
        } else {
            const ThermalEnv E[2] = {
                {ca.f, ca.J, ca.plan.dt, ca.T, ca.T, ca.plan.n,
                 make_stream(a.seed, a.env_offset + (uint64_t)eA, (uint32_t)a.state.episode[eA], (uint32_t)ca.step)},
                {cb.f, cb.J, cb.plan.dt, cb.T, cb.T, cb.plan.n,
                 make_stream(a.seed, a.env_offset + (uint64_t)eB, (uint32_t)a.state.episode[eB], (uint32_t)cb.step)}};
            double w[2][3] = {{ca.w[0], ca.w[1], ca.w[2]}, {cb.w[0], cb.w[1], cb.w[2]}};
            int g[2] = {ca.guard, cb.guard};
            ThermalSource<F2, NOISE == 3 ? 1 : 0> src;
            const NoiseStream nss[2] = {E[0].ns, E[1].ns};
            src.init(nss, mk2(thermal_nscale(ca.f, ca.plan.dt), thermal_nscale(cb.f, cb.plan.dt)));
            integrate_thermal<F2>(E, src, w, g);
            ca.w[0] = w[0][0]; ca.w[1] = w[0][1]; ca.w[2] = w[0][2]; ca.guard = g[0];
            cb.w[0] = w[1][0]; cb.w[1] = w[1][1]; cb.w[2] = w[1][2]; cb.guard = g[1];
        }
    } else {
        if (ca.valid) env_step_integrate<float, true, NOISE, false>(a, eA, ca);
        if (cb.valid) env_step_integrate<float, true, NOISE, false>(a, eB, cb);
    }
    if (!ca.illcond) env_step_epilogue(a, eA, ca, rA);
    if (!cb.illcond) env_step_epilogue(a, eB, cb, rB);
    return (ca.illcond ? 1 : 0) | (cb.illcond ? 2 : 0);
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
        const NoiseStream ns = make_stream(a.seed, a.env_offset + (uint64_t)e, 0u, 0u);
        const double* nrow = (NOISE == 2) ? a.d_noise + (int64_t)e * a.noise_stride * (EULER ? 3 : 12) : nullptr;
        double* traj = a.d_traj ? a.d_traj + (int64_t)e * a.traj_stride * 3 : nullptr;
        // the solver API repeats an ill-conditioned FP32 trajectory with FP64 stages on the spot (CondTrack, llgs_core.cuh)
        const double sx = mx, sy = my, sz = mz;
        const int guard0 = guard;
        int ill = 0;
        if (f[FI_HTH] > 0.0 || NOISE == 0)
            integrate<R, AXIS_Z, NOISE, EULER>(f, J, mx, my, mz, plan.n, plan.dt, t_pulse, t_end, ns, nrow, traj, guard,
                                               NOISE == 2 ? a.noise_stride : 0x7fffffff, traj ? a.traj_stride : 0x7fffffff, &ill);
        else
            integrate<R, AXIS_Z, 0, EULER>(f, J, mx, my, mz, plan.n, plan.dt, t_pulse, t_end, ns, nullptr, traj, guard, 0x7fffffff, traj ? a.traj_stride : 0x7fffffff, &ill);
        if (sizeof(R) == 4 && ill) {
            mx = sx; my = sy; mz = sz; guard = guard0;
            if (f[FI_HTH] > 0.0 || NOISE == 0)
                integrate<double, AXIS_Z, NOISE, EULER>(f, J, mx, my, mz, plan.n, plan.dt, t_pulse, t_end, ns, nrow, traj, guard,
                                                        NOISE == 2 ? a.noise_stride : 0x7fffffff,
                                                        traj ? a.traj_stride : 0x7fffffff);
            else
                integrate<double, AXIS_Z, 0, EULER>(f, J, mx, my, mz, plan.n, plan.dt, t_pulse, t_end, ns, nullptr, traj, guard, 0x7fffffff,
                                                    traj ? a.traj_stride : 0x7fffffff);
        }
    }
    a.d_m_out[3 * e] = mx; a.d_m_out[3 * e + 1] = my; a.d_m_out[3 * e + 2] = mz;
    if (a.d_n_sub) a.d_n_sub[e] = nsub;
    if (a.d_guard) a.d_guard[e] = guard;
}

// Fixed-step solve with current_func / field_func sampled by the host at the reference's stage times (include/stg.h,
// StgSttSolveArgs): row r of jgrid holds J at (t_i, t_i + dt/2, t_i + dt), hgrid the applied field at the same times. FP64
// general-geometry stages; a rectangular pulse (J, t_pulse) is used when only the field is sampled.
template <int NOISE, bool EULER>
STG_HD void integrate_grid(const double* f, double J, double& mx, double& my, double& mz, int n, double dt, double t_pulse,
                           const double* jgrid, const double* hgrid, int64_t grid_rows, const NoiseStream& ns,
                           const double* noise_row, double* traj, int& guard, int64_t noise_rows, int64_t traj_rows) {
    constexpr bool TH = NOISE != 0;
    constexpr int NS = EULER ? 3 : 12;
    StepConsts<double> c;
    make_consts<double>(f, dt, 0.0, 1.0, c);
    const double G = -f[FI_GEFF] * dt;
    const float nscale = -1.3862943611198906f * (float)c.cth * (float)c.cth;
    ScaledState st{mx, my, mz, 1.0, 1.0, 1.0f};
    if (traj) { traj[0] = mx; traj[1] = my; traj[2] = mz; }
    ThermalSource<float> src;
    if (NOISE == 1 && !EULER) src.init(&ns, nscale);
    for (int i = 0; i < n; ++i) {
        const int64_t r = i < grid_rows ? i : grid_rows - 1;
        double nz[NS];
        if (NOISE == 1) {
            float z[12];
            if (EULER) philox_normals3(ns, (uint32_t)i, nscale, z);
            else if (i & 1) src.second((uint32_t)i >> 1, z);
            else src.first((uint32_t)i >> 1, z);
#pragma unroll
            for (int q = 0; q < NS; ++q) nz[q] = (double)z[q];
        } else if (NOISE == 2) {
            const int64_t nr = i < noise_rows ? i : noise_rows - 1;
#pragma unroll
            for (int q = 0; q < NS; ++q) nz[q] = c.cth * noise_row[nr * NS + q];
        }
        // stage s (0..3) reads the sample of time group g: 0 -> t_i, 1 -> t_i + dt/2, 2 -> t_i + dt
        auto stage = [&](int g, int s, double x, double y, double z, double& kx, double& ky, double& kz) {
            StepConsts<double> cg = c;
            if (hgrid) {
                const double* h = hgrid + (r * 3 + g) * 3;
                cg.bax += G * h[0]; cg.bay += G * h[1]; cg.baz += G * h[2];
            }
            const double jt = jgrid ? jgrid[r * 3 + g] : (pulse_on(i, g, dt, t_pulse) ? J : 0.0);
            const double a = (fabs(jt) > 1e-12) ? f[FI_AJ_PER_J] * jt * dt : 0.0;      // physics/simple_solver.py:327-331
            stage_general<double, TH>(cg, x, y, z, a, 0.0, TH ? nz[3 * s] : 0.0, TH ? nz[3 * s + 1] : 0.0,
                                      TH ? nz[3 * s + 2] : 0.0, kx, ky, kz);
        };
        const double x = st.sx, y = st.sy, z = st.z;
        double k1x, k1y, k1z, ix, iy, iz;
        stage(0, 0, x, y, z, k1x, k1y, k1z);
        if (EULER) {
            ix = k1x; iy = k1y; iz = k1z;
        } else {
            double k2x, k2y, k2z, k3x, k3y, k3z, k4x, k4y, k4z;
            stage(1, 1, x + 0.5 * k1x, y + 0.5 * k1y, z + 0.5 * k1z, k2x, k2y, k2z);
            stage(1, 2, x + 0.5 * k2x, y + 0.5 * k2y, z + 0.5 * k2z, k3x, k3y, k3z);
            stage(2, 3, x + k3x, y + k3y, z + k3z, k4x, k4y, k4z);
            ix = (k1x + 2.0 * (k2x + k3x) + k4x) / 6.0;
            iy = (k1y + 2.0 * (k2y + k3y) + k4y) / 6.0;
            iz = (k1z + 2.0 * (k2z + k3z) + k4z) / 6.0;
        }
        st.sx += ix; st.sy += iy; st.z += iz;
        guard_normalise<double>(st, guard);
        if (traj && i + 1 < traj_rows) { traj[3 * (i + 1)] = st.sx; traj[3 * (i + 1) + 1] = st.sy; traj[3 * (i + 1) + 2] = st.z; }
    }
    mx = st.sx; my = st.sy; mz = st.z;
}

template <int NOISE, bool EULER>
STG_HD void solve_grid_body(const StgSttSolveArgs& a, int64_t e) {
    const double* f = a.d_table[a.d_param_index ? a.d_param_index[e] : 0].v;
    double mx = a.d_m0[3 * e], my = a.d_m0[3 * e + 1], mz = a.d_m0[3 * e + 2];
    const double J = a.d_pulse[3 * e], t_pulse = a.d_pulse[3 * e + 1], t_end = a.d_pulse[3 * e + 2];
    int guard = 0, nsub = 0;
    guard_normalise<double>(mx, my, mz, guard);
    if (t_end > 0.0) {
        const StepPlan plan = substep_plan(t_end, f[FI_MAXSTEP_DT]);
        nsub = plan.n;
        const NoiseStream ns = make_stream(a.seed, a.env_offset + (uint64_t)e, 0u, 0u);
        const int64_t ge = a.grid_envs == 1 ? 0 : e;
        const double* jg = a.d_current_grid ? a.d_current_grid + ge * a.grid_stride * 3 : nullptr;
        const double* hg = a.d_field_grid ? a.d_field_grid + ge * a.grid_stride * 9 : nullptr;
        const double* nrow = (NOISE == 2) ? a.d_noise + (int64_t)e * a.noise_stride * (EULER ? 3 : 12) : nullptr;
        double* traj = a.d_traj ? a.d_traj + (int64_t)e * a.traj_stride * 3 : nullptr;
        if (f[FI_HTH] > 0.0 || NOISE == 0)
            integrate_grid<NOISE, EULER>(f, J, mx, my, mz, plan.n, plan.dt, t_pulse, jg, hg, a.grid_stride, ns, nrow, traj, guard,
                                         NOISE == 2 ? a.noise_stride : 0x7fffffff, traj ? a.traj_stride : 0x7fffffff);
        else
            integrate_grid<0, EULER>(f, J, mx, my, mz, plan.n, plan.dt, t_pulse, jg, hg, a.grid_stride, ns, nullptr, traj, guard, 0x7fffffff,
                                     traj ? a.traj_stride : 0x7fffffff);
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

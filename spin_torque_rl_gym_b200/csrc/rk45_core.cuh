// rk45_core.cuh — K2: adaptive Dormand-Prince 5(4) integration of the generalised LLGS right-hand side, one trajectory
// per thread, with SciPy's RK45 step-size controller reproduced decision for decision.
//
// Replaces LLGSSolver.solve (physics/llgs_solver.py:51-180 of the reference), whose arithmetic lives in the third-party
// dependency scipy.integrate.solve_ivp (pyproject pins scipy>=1.7.0; 1.18.1 installed): scipy/integrate/_ivp/rk.py
// (rk_step :14-72, RungeKutta._step_impl :111-167, RK45 tableau :281-378), common.py (select_initial_step :68-133, norm :63).
// Controller constants: SAFETY 0.9, MIN_FACTOR 0.2, MAX_FACTOR 10, error exponent -1/5, RMS error norm against
// atol + rtol*max(|y|,|y_new|), first step from select_initial_step, min_step = 10*ulp(t), last step clipped to t_bound,
// no growth right after a rejected step. FP64 throughout (atol = 1e-9 is below FP32 resolution).
#pragma once

#include "../../include/stg.h"
#include "llgs_core.cuh"

namespace stg {

struct Rk45Tableau {
    // scipy/integrate/_ivp/rk.py:281-302 (Dormand & Prince 1980)
    static constexpr double c2 = 1.0 / 5, c3 = 3.0 / 10, c4 = 4.0 / 5, c5 = 8.0 / 9;
    static constexpr double a21 = 1.0 / 5;
    static constexpr double a31 = 3.0 / 40, a32 = 9.0 / 40;
    static constexpr double a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9;
    static constexpr double a51 = 19372.0 / 6561, a52 = -25360.0 / 2187, a53 = 64448.0 / 6561, a54 = -212.0 / 729;
    static constexpr double a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176,
                            a65 = -5103.0 / 18656;
    static constexpr double b1 = 35.0 / 384, b3 = 500.0 / 1113, b4 = 125.0 / 192, b5 = -2187.0 / 6784, b6 = 11.0 / 84;
    static constexpr double e1 = -71.0 / 57600, e3 = 71.0 / 16695, e4 = -71.0 / 1920, e5 = 17253.0 / 339200,
                            e6 = -22.0 / 525, e7 = 1.0 / 40;
};

// On the device the coefficients are read from constant memory (an FP64 literal costs two uniform moves each time it is used:
// 12.5 % of the executed instructions of the attempt loop before).
enum RkIdx { RK_c2, RK_c3, RK_c4, RK_c5, RK_a21, RK_a31, RK_a32, RK_a41, RK_a42, RK_a43, RK_a51, RK_a52, RK_a53, RK_a54, RK_a61, RK_a62, RK_a63, RK_a64, RK_a65, RK_b1, RK_b3, RK_b4, RK_b5, RK_b6, RK_e1, RK_e3, RK_e4, RK_e5, RK_e6, RK_e7 };
#if defined(__CUDACC__)
static __constant__ double kRkTab[] = {Rk45Tableau::c2, Rk45Tableau::c3, Rk45Tableau::c4, Rk45Tableau::c5, Rk45Tableau::a21, Rk45Tableau::a31, Rk45Tableau::a32, Rk45Tableau::a41, Rk45Tableau::a42, Rk45Tableau::a43, Rk45Tableau::a51, Rk45Tableau::a52, Rk45Tableau::a53, Rk45Tableau::a54, Rk45Tableau::a61, Rk45Tableau::a62, Rk45Tableau::a63, Rk45Tableau::a64, Rk45Tableau::a65, Rk45Tableau::b1, Rk45Tableau::b3, Rk45Tableau::b4, Rk45Tableau::b5, Rk45Tableau::b6, Rk45Tableau::e1, Rk45Tableau::e3, Rk45Tableau::e4, Rk45Tableau::e5, Rk45Tableau::e6, Rk45Tableau::e7};
#endif
#if defined(__CUDA_ARCH__)
#define RKT(name) kRkTab[RK_##name]
#else
#define RKT(name) Rk45Tableau::name
#endif

struct V3 {
    double x, y, z;
};
STG_HD V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
STG_HD V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
STG_HD V3 cross3(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
STG_HD double dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
STG_HD double rms3(V3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z) / 1.7320508075688772; }   // common.py:63-65

// per-trajectory inputs resolved from the parameter set + the per-env controls
struct LlgRhs {
    const StgLlgParams* p;
    // hot parameters copied out of the table once per trajectory (the RHS is evaluated ~800 times)
    double gamma, alpha, ex, ey, ez, msnx, msny, msnz, exch, hth, cdp, cfp, cds, cfs, sgx, sgy, sgz;
    STG_HD void load() {
        const StgLlgParams& q = *p;
        gamma = q.gamma; alpha = q.alpha; ex = q.easy_axis[0]; ey = q.easy_axis[1]; ez = q.easy_axis[2];
        msnx = -q.saturation_magnetization * q.demag_n[0]; msny = -q.saturation_magnetization * q.demag_n[1];
        msnz = -q.saturation_magnetization * q.demag_n[2];
        exch = q.exchange_coeff; hth = q.h_th; cdp = q.c_dl_p; cfp = q.c_fl_p; cds = q.c_dl_s; cfs = q.c_fl_s;
        sgx = q.sigma[0]; sgy = q.sigma[1]; sgz = q.sigma[2];
    }
    double hk;          // 2 K_eff /(mu0 Ms)
    double J, t_pulse;
    V3 happ;
    // current_func(t), field_func(t): rectangular pulse + constant field, or the piecewise-constant tables of the argument block
    // (StgRk45Args.d_seg_*; read through the kernel parameter, so they cost the trajectory no registers)
    // SEG: compile-time switch, so that the rectangular-pulse instantiation carries no trace of the tables
    template <bool SEG>
    STG_HD void controls(const StgRk45Args& a, int64_t e, double t, double& cur, V3& h) const {
        if constexpr (!SEG) {
            cur = (t <= t_pulse) ? J : 0.0;
            h = happ;
            return;
        }
        const int64_t row = a.seg_rows > 1 ? e : 0;
        const double* st = a.d_seg_t + row * a.n_seg;
        int k = 0;
        while (k < a.n_seg && t > st[k]) ++k;
        cur = a.d_seg_current[row * (a.n_seg + 1) + k];
        if (a.d_seg_field) {
            const double* sh = a.d_seg_field + (row * (a.n_seg + 1) + k) * 3;
            h = V3{sh[0], sh[1], sh[2]};
        } else {
            h = happ;
        }
    }
    // thermal
    int noise_mode;     // 0 none, 1 Philox, 2 injected
    Philox ph;
    uint64_t gid;
    const double* noise_row;
    int64_t noise_cap;
    int n_eval;

    // unit vector of y (physics/llgs_solver.py:96-101; y * (1/|y|) is <= 1 ulp from NumPy's y / |y|). The state stays within
    // ~1e-6 of the unit sphere between accepted steps, where four terms of (1 + d)^(-1/2) are exact to 1e-17 (|d| < 2^-13) and
    // save the FP64 reciprocal square root (MUFU seed + Newton steps) of every RHS evaluation.
    STG_HD V3 unit(V3 y) const {
        V3 m = {0.0, 0.0, 1.0};
        const double n2 = dot3(y, y);
        const double d = n2 - 1.0;
        if (fabs(d) < 1.220703125e-4) {
            const double inv = 1.0 + d * (-0.5 + d * (0.375 + d * (-0.3125 + d * 0.2734375)));
            m.x = y.x * inv; m.y = y.y * inv; m.z = y.z * inv;
        } else if (n2 > 1e-24) {
#if defined(__CUDA_ARCH__)
            const double inv = rsqrt(n2);
#else
            const double inv = 1.0 / sqrt(n2);
#endif
            m.x = y.x * inv; m.y = y.y * inv; m.z = y.z * inv;
        }
        return m;
    }
    // effective field without the thermal part (physics/llgs_solver.py:182-211 + devices/*:compute_effective_field)
    STG_HD V3 field(V3 m, V3 ha) const {
        const V3 ea = {ex, ey, ez};
        const double s = hk * dot3(m, ea);
        V3 h = {ha.x + s * ea.x, ha.y + s * ea.y, ha.z + s * ea.z};
        h.x += msnx * m.x;                                    // -Ms N (.) m
        h.y += msny * m.y;
        h.z += msnz * m.z;
        if (exch != 0.0) { h.x += exch * m.x; h.y += exch * m.y; h.z += exch * m.z; }
        return h;
    }

    // llgs_rhs (physics/llgs_solver.py:92-126) generalised with the SOT terms (devices/sot_mram.py:163-194)
    template <bool SEG>
    STG_HD V3 eval(const StgRk45Args& a, int64_t e, double t, V3 y) {
        const V3 m = unit(y);
        double cur;
        V3 ha;
        controls<SEG>(a, e, t, cur, ha);
        V3 h = field(m, ha);
        if (noise_mode != 0 && hth > 0.0) {
            double nx, ny, nz;
            if (noise_mode == 2) {
                const int64_t k = n_eval < noise_cap ? n_eval : noise_cap - 1;
                nx = noise_row[3 * k]; ny = noise_row[3 * k + 1]; nz = noise_row[3 * k + 2];
            } else {
                uint32_t o[4];
                ph((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)n_eval, 0x524b3435u, o);
                float z0, z1, z2, z3;
                box_muller(o[0], o[1], z0, z1);
                box_muller(o[2], o[3], z2, z3);
                nx = z0; ny = z1; nz = z2;
            }
            h.x += hth * nx; h.y += hth * ny; h.z += hth * nz;
        }
        ++n_eval;
        V3 tau = {0.0, 0.0, 0.0};
        if (!(fabs(cur) < 1e-12)) {                            // :221-222
            if (cdp != 0.0 || cfp != 0.0) {                    // Slonczewski pair, p = z^ (uniform branch per parameter set)
                const V3 ph3 = {p->p_hat[0], p->p_hat[1], p->p_hat[2]};
                const V3 mxp = cross3(m, ph3);
                tau = tau + (cdp * cur) * cross3(m, mxp) + (cfp * cur) * mxp;
            }
            if (cds != 0.0 || cfs != 0.0) {                    // spin-orbit pair
                const V3 sg = {sgx, sgy, sgz};
                tau = tau + (cds * cur) * cross3(sg, m) + (cfs * cur) * sg;
            }
        }
        V3 dm = (-gamma) * cross3(m, h);                       // :121-124
        dm = dm + alpha * cross3(m, dm);
        return dm + tau;
    }

    // _compute_energy (physics/llgs_solver.py:239-262) and the torque norm of :168-172 for a NORMALISED m
    template <bool SEG>
    STG_HD void diagnostics(const StgRk45Args& a, int64_t env, double t, V3 m, double& energy, double& torque) const {
        const StgLlgParams& q = *p;
        const V3 e = {q.easy_axis[0], q.easy_axis[1], q.easy_axis[2]};
        const double msv = q.saturation_magnetization * q.volume;
        double cur;
        V3 ha;
        controls<SEG>(a, env, t, cur, ha);
        const double ez = -q.mu0 * msv * dot3(m, ha);
        const double c = dot3(m, e);
        const double ku = hk * q.mu0 * q.saturation_magnetization * 0.5;
        const double ea = -ku * q.volume * c * c;
        const double ed = 0.5 * q.mu0 * q.saturation_magnetization * msv *
                          (q.demag_n[0] * m.x * m.x + q.demag_n[1] * m.y * m.y + q.demag_n[2] * m.z * m.z);
        energy = ez + ea + ed;
        torque = 0.0;
        if (!(fabs(cur) < 1e-12)) {
            const V3 ph3 = {q.p_hat[0], q.p_hat[1], q.p_hat[2]};
            const V3 mxp = cross3(m, ph3);
            const V3 dl = (q.c_dl_p * cur) * cross3(m, mxp);
            const V3 fl = (q.c_fl_p * cur) * mxp;
            const V3 sg = {q.sigma[0], q.sigma[1], q.sigma[2]};
            const V3 sdl = (q.c_dl_s * cur) * cross3(sg, m) + dl;
            const V3 sfl = (q.c_fl_s * cur) * sg + fl;
            torque = sqrt(dot3(sdl, sdl)) + sqrt(dot3(sfl, sfl));
        }
    }
};

STG_HD double ulp_above(double t) {   // |nextafter(t, +inf) - t| (the integration always runs forward)
#if defined(__CUDA_ARCH__)
    if (t > 0.0) return __longlong_as_double(__double_as_longlong(t) + 1LL) - t;
    if (t < 0.0) return __longlong_as_double(__double_as_longlong(t) - 1LL) - t;
    return 4.9406564584124654e-324;
#else
    return nextafter(t, (double)INFINITY) - t;
#endif
}
// en^(-1/5) of the step-size controller (rk.py:102) from en^2 (the RMS norm's square root is never taken): x = (en^2)^(-1/10)
// by one Newton step on x^-10 = en^2 from a single-precision seed (lg2 / ex2, relative error ~1e-7 -> ~1e-13 after the step,
// far below anything that changes an accept / reject decision or a step count). pow() and exp(log()) in FP64 were 12 % of the
// kernel's stall samples.
STG_HD double pow_m01(double en2) {
#if defined(__CUDA_ARCH__)
    if (!(en2 > 1e-30 && en2 < 1e30)) return exp(-0.1 * log(en2));
    double x = (double)exp2f(-0.1f * __log2f((float)en2));
    const double x2 = x * x, x4 = x2 * x2, x5 = x4 * x;
    return x * (1.1 - 0.1 * en2 * (x5 * x5));
#else
    return pow(en2, -0.1);
#endif
}
STG_HD V3 vabs_max(V3 a, V3 b) { return {fmax(fabs(a.x), fabs(b.x)), fmax(fabs(a.y), fabs(b.y)), fmax(fabs(a.z), fabs(b.z))}; }

// One trajectory of LLGSSolver.solve. Returns through the StgRk45Args output arrays of env e.
// One trajectory's integration state. rk45_init / rk45_attempt / rk45_finish are the three phases of SciPy's solve_ivp loop
// for one trajectory; rk45_body chains them (host build, one-trajectory-per-thread kernel), the refill kernel interleaves them
// across trajectories so that a lane whose trajectory has finished starts the next one instead of idling.
struct Rk45State {
    LlgRhs f;
    V3 y, fk;
    double t, tb, h_abs, min_step;
    double* traj;
    int64_t e, attempts;
    int n_acc, n_rej, status, overflow;
    bool new_step, rejected;
    STG_HD bool running() const { return t != tb && status == 0; }
};

template <bool SEG>
STG_HD void rk45_record(const StgRk45Args& a, Rk45State& S, int row, double tt, V3 yy) {
    if (!S.traj) return;
    if (row >= a.traj_stride) { S.overflow = 2; return; }   // keeps integrating; only the recording stops
    const double n = sqrt(dot3(yy, yy));
    const V3 m = {yy.x / n, yy.y / n, yy.z / n};            // :152-153
    double en, tq;
    S.f.template diagnostics<SEG>(a, S.e, tt, m, en, tq);
    double* r = S.traj + 6 * (int64_t)row;
    r[0] = tt; r[1] = m.x; r[2] = m.y; r[3] = m.z; r[4] = en; r[5] = tq;
}

template <bool SEG>
STG_HD void rk45_init(const StgRk45Args& a, int64_t e, Rk45State& S) {
    const StgLlgParams& q = a.d_table[a.d_param_index ? a.d_param_index[e] : 0];
    LlgRhs& f = S.f;
    f.p = &q;
    f.load();
    f.J = a.d_current ? a.d_current[e] : 0.0;
    f.t_pulse = a.d_t_pulse ? a.d_t_pulse[e] : 1.0e300;
    f.happ = a.d_happ ? V3{a.d_happ[3 * e], a.d_happ[3 * e + 1], a.d_happ[3 * e + 2]} : V3{0.0, 0.0, 0.0};
    double ku = q.uniaxial_anisotropy;
    if (q.use_vcma) {   // devices/vcma_mram.py:122-147
        double v = a.d_voltage ? a.d_voltage[e] : 0.0;
        v = fmin(fmax(v, -q.breakdown_voltage), q.breakdown_voltage);
        const double k = ku + (-q.vcma_coefficient * fabs(v) / (q.dielectric_thickness * q.dielectric_thickness));
        ku = fmax(k, -0.5 * ku);
    }
    f.hk = 2.0 * ku / (q.mu0 * q.saturation_magnetization);
    f.noise_mode = (a.flags & STG_F_THERMAL_INJECT) ? 2 : ((a.flags & STG_F_THERMAL_PHILOX) ? 1 : 0);
    f.ph = Philox{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
    f.gid = a.env_offset + (uint64_t)e;
    f.noise_row = a.d_noise ? a.d_noise + (int64_t)e * a.noise_stride * 3 : nullptr;
    f.noise_cap = a.noise_stride;
    f.n_eval = 0;

    const double t0 = a.d_t_start ? a.d_t_start[e] : 0.0;
    S.e = e;
    S.tb = a.d_t_end[e];
    const double rtol = a.rtol, atol = a.atol, max_step = a.max_step;
    V3 y = {a.d_m0[3 * e], a.d_m0[3 * e + 1], a.d_m0[3 * e + 2]};
    {   // m_initial / ||m_initial|| (physics/llgs_solver.py:75)
        const double n0 = sqrt(dot3(y, y));
        y = {y.x / n0, y.y / n0, y.z / n0};
    }
    S.y = y;
    S.t = t0;
    S.n_acc = 0; S.n_rej = 0; S.status = 0; S.overflow = 0;
    S.attempts = 0;
    S.new_step = true; S.rejected = false;
    S.min_step = 0.0;
    S.h_abs = 0.0;
    S.fk = V3{0.0, 0.0, 0.0};
    S.traj = a.d_traj ? a.d_traj + (int64_t)e * a.traj_stride * 6 : nullptr;
    rk45_record<SEG>(a, S, 0, S.t, y);
    if (S.tb > t0) {
        const V3 fk = f.template eval<SEG>(a, e, S.t, y);                          // rk.py:94
        // select_initial_step (common.py:68-133), order = 4
        const double interval = fabs(S.tb - t0);
        const V3 sc = {atol + fabs(y.x) * rtol, atol + fabs(y.y) * rtol, atol + fabs(y.z) * rtol};
        const double d0 = rms3({y.x / sc.x, y.y / sc.y, y.z / sc.z});
        const double d1 = rms3({fk.x / sc.x, fk.y / sc.y, fk.z / sc.z});
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, interval);
        const V3 y1 = y + h0 * fk;
        const V3 f1 = f.template eval<SEG>(a, e, t0 + h0, y1);
        const double d2 = rms3({(f1.x - fk.x) / sc.x, (f1.y - fk.y) / sc.y, (f1.z - fk.z) / sc.z}) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
        S.h_abs = fmin(fmin(100.0 * h0, h1), fmin(interval, max_step));
        S.fk = fk;
    } else {
        S.tb = S.t;                                             // t_end <= 0: nothing to integrate (running() is false)
    }
}

// ONE attempted step of a running trajectory: solve_ivp main loop (ivp.py), RungeKutta._step_impl (rk.py:111-167).
// SciPy nests "while not step_accepted" inside the stepping loop; here one flat loop performs one attempt per iteration for
// every lane that is still integrating, so the lanes of a warp stay converged on the six RHS evaluations whether their previous
// attempt was accepted or rejected (nested loops make the whole warp pay for every lane's rejection).
template <bool SEG>
STG_HD void rk45_attempt(const StgRk45Args& a, Rk45State& S) {
    const double rtol = a.rtol, atol = a.atol, max_step = a.max_step;
    const int64_t max_attempts = a.max_attempts > 0 ? a.max_attempts : 1000000;
    LlgRhs& f = S.f;
    if (S.new_step) {                                   // head of _step_impl
        S.min_step = 10.0 * ulp_above(S.t);
        if (S.h_abs > max_step) S.h_abs = max_step;
        else if (S.h_abs < S.min_step) S.h_abs = S.min_step;
        S.rejected = false;
        S.new_step = false;
    }
    if (S.h_abs < S.min_step) { S.status |= 1; return; }     // TOO_SMALL_STEP
    if (++S.attempts > max_attempts) { S.status |= 4; return; }
    const double t = S.t;
    const V3 y = S.y;
    double h = S.h_abs;
    double t_new = t + h;
    if (t_new - S.tb > 0.0) t_new = S.tb;
    h = t_new - t;
    S.h_abs = fabs(h);
    // rk_step (rk.py:14-72)
    const V3 k1 = S.fk;
    const V3 k2 = f.template eval<SEG>(a, S.e, t + RKT(c2) * h, y + h * (RKT(a21) * k1));
    const V3 k3 = f.template eval<SEG>(a, S.e, t + RKT(c3) * h, y + h * (RKT(a31) * k1 + RKT(a32) * k2));
    const V3 k4 = f.template eval<SEG>(a, S.e, t + RKT(c4) * h, y + h * (RKT(a41) * k1 + RKT(a42) * k2 + RKT(a43) * k3));
    const V3 k5 = f.template eval<SEG>(a, S.e, t + RKT(c5) * h, y + h * (RKT(a51) * k1 + RKT(a52) * k2 + RKT(a53) * k3 + RKT(a54) * k4));
    const V3 k6 = f.template eval<SEG>(a, S.e, t + h, y + h * (RKT(a61) * k1 + RKT(a62) * k2 + RKT(a63) * k3 + RKT(a64) * k4 + RKT(a65) * k5));
    const V3 y_new = y + h * (RKT(b1) * k1 + RKT(b3) * k3 + RKT(b4) * k4 + RKT(b5) * k5 + RKT(b6) * k6);
    const V3 f_new = f.template eval<SEG>(a, S.e, t + h, y_new);
    const V3 err = h * (RKT(e1) * k1 + RKT(e3) * k3 + RKT(e4) * k4 + RKT(e5) * k5 + RKT(e6) * k6 + RKT(e7) * f_new);
    const V3 mx = vabs_max(y, y_new);
    // en^2 = mean((err / scale)^2) with ONE division: sum_i (err_i prod_{j != i} s_j)^2 / (3 (s_x s_y s_z)^2)
    const double sx = atol + mx.x * rtol, sy = atol + mx.y * rtol, sz = atol + mx.z * rtol;
    const double sxy = sx * sy, ax = err.x * (sy * sz), ay = err.y * (sx * sz), az = err.z * sxy, sp = sxy * sz;
    const double en2 = (ax * ax + ay * ay + az * az) / (3.0 * (sp * sp));
    if (en2 < 1.0) {
        double factor = (en2 == 0.0) ? 10.0 : fmin(10.0, 0.9 * pow_m01(en2));
        if (S.rejected) factor = fmin(1.0, factor);
        S.h_abs *= factor;
        S.t = t_new; S.y = y_new; S.fk = f_new;
        ++S.n_acc;
        S.new_step = true;
        rk45_record<SEG>(a, S, S.n_acc, S.t, S.y);
    } else if (en2 >= 1.0) {
        S.h_abs *= fmax(0.2, 0.9 * pow_m01(en2));
        S.rejected = true;
        ++S.n_rej;
    } else {               // NaN error norm: SciPy would never terminate; flag and stop
        S.status |= 8;
    }
}

STG_HD void rk45_finish(const StgRk45Args& a, const Rk45State& S) {
    const int64_t e = S.e;
    a.d_y_out[3 * e] = S.y.x; a.d_y_out[3 * e + 1] = S.y.y; a.d_y_out[3 * e + 2] = S.y.z;
    if (a.d_n_accepted) a.d_n_accepted[e] = S.n_acc;
    if (a.d_n_rejected) a.d_n_rejected[e] = S.n_rej;
    if (a.d_n_rhs) a.d_n_rhs[e] = S.f.n_eval;
    if (a.d_status) a.d_status[e] = S.status | S.overflow;
    if (a.d_t_reached) a.d_t_reached[e] = S.t;
}

// Estimated number of attempted steps of trajectory e, for grouping trajectories of similar cost into warps (a lane idles once
// its trajectory has ended): the controller's step is capped by max_step and otherwise shrinks with the precession frequency
// gamma |H_eff(m0)| (+ the torque rate). Only the ORDER matters; measured on the SOT / VCMA mix the lanes of a warp are 95 % busy
// when sorted by this key against 59 % in caller order (oracle sort by the true count: 97 %).
template <bool SEG>
STG_HD double rk45_cost_estimate(const StgRk45Args& a, int64_t e) {
    Rk45State S;
    const int64_t saved_stride = a.traj_stride;
    (void)saved_stride;
    LlgRhs& f = S.f;
    const StgLlgParams& q = a.d_table[a.d_param_index ? a.d_param_index[e] : 0];
    f.p = &q;
    f.load();
    f.J = a.d_current ? a.d_current[e] : 0.0;
    f.t_pulse = a.d_t_pulse ? a.d_t_pulse[e] : 1.0e300;
    f.happ = a.d_happ ? V3{a.d_happ[3 * e], a.d_happ[3 * e + 1], a.d_happ[3 * e + 2]} : V3{0.0, 0.0, 0.0};
    double ku = q.uniaxial_anisotropy;
    if (q.use_vcma) {
        double v = a.d_voltage ? a.d_voltage[e] : 0.0;
        v = fmin(fmax(v, -q.breakdown_voltage), q.breakdown_voltage);
        ku = fmax(ku + (-q.vcma_coefficient * fabs(v) / (q.dielectric_thickness * q.dielectric_thickness)), -0.5 * ku);
    }
    f.hk = 2.0 * ku / (q.mu0 * q.saturation_magnetization);
    const double t0 = a.d_t_start ? a.d_t_start[e] : 0.0;
    const double len = a.d_t_end[e] - t0;
    if (!(len > 0.0)) return 0.0;
    const V3 m = f.unit(V3{a.d_m0[3 * e], a.d_m0[3 * e + 1], a.d_m0[3 * e + 2]});
    double cur;
    V3 ha;
    f.template controls<SEG>(a, e, t0, cur, ha);
    const V3 h = f.field(m, ha);
    const double omega = f.gamma * sqrt(dot3(h, h)) + fabs(cur) * (fabs(f.cdp) + fabs(f.cfp) + fabs(f.cds) + fabs(f.cfs));
    return len * fmax(1.0 / a.max_step, omega * 8.0);            // ~0.12 rad per step at rtol 1e-6
}

template <bool SEG>
STG_HD void rk45_body(const StgRk45Args& a, int64_t e) {
    Rk45State S;
    rk45_init<SEG>(a, e, S);
    while (S.running()) rk45_attempt<SEG>(a, S);
    rk45_finish(a, S);
}

}  // namespace stg

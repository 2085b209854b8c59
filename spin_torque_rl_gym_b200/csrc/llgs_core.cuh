// llgs_core.cuh — macrospin LLGS arithmetic shared by every kernel of libstg (sm_100a).
//
// One env lives in one thread. The magnetisation is carried in FP64 across all substeps of an env step; the four RK4
// stage derivatives are evaluated in the arithmetic type R (float for the `_f32` entry points, double for `_f64`), with
// all step constants pre-multiplied so that a stage is a handful of FMAs:
//
//     b  = G*H(m)            G = -gamma/(1+alpha^2) * dt        (rotation vector of this substep, dimensionless)
//     k  = dt*f(m) = m x b + alpha * m x (m x b) + (aJ*dt) * m x (m x e)
//
// which is algebraically the reference's  dt * ( -geff*(m x H + alpha m x (m x H)) + aJ m x (m x e) )
// (physics/simple_solver.py:297-344 of the reference; formulas restated in oracle/stt_oracle.py).
//
// The header compiles for the device (nvcc) and, for the CPU-side arithmetic checks in tests/hostsim, for the host (g++).
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define STG_HD __host__ __device__ __forceinline__
#else
#define STG_HD inline
#endif

namespace stg {

// ---- folded parameter-set layout (StgSttFolded::v) ------------------------------------------------------------------
enum FoldedIdx {
    FI_ALPHA = 0, FI_GEFF, FI_HK, FI_MS, FI_AJ_PER_J, FI_HTH,
    FI_EX, FI_EY, FI_EZ, FI_REFX, FI_REFY, FI_REFZ,
    FI_RP, FI_RAP, FI_AREA, FI_RSERIES, FI_TEMP,
    FI_HAX, FI_HAY, FI_HAZ,
    FI_MAXCUR, FI_MAXDUR, FI_SUCC, FI_WE, FI_MAXSTEP_DT,
    FI_MAXSTEPS, FI_KIND, FI_THERMAL, FI_VALID, FI_AXISZ, FI_TMR,
    FI_COUNT
};

constexpr double kGamma = 2.21e5;                    // physics/simple_solver.py:59
constexpr double kMu0 = 4.0 * 3.141592653589793 * 1e-7;  // physics/simple_solver.py:60 (4*np.pi*1e-7)
constexpr double kKbSolver = 1.38e-23;               // physics/simple_solver.py:377

// ---- Philox4x32-10 counter-based generator (Salmon et al., SC'11) ---------------------------------------------------
struct Philox {
    uint32_t k0, k1;
    STG_HD static void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
        lo = a * b;
        hi = __umulhi(a, b);
#else
        uint64_t p = (uint64_t)a * (uint64_t)b;
        lo = (uint32_t)p;
        hi = (uint32_t)(p >> 32);
#endif
    }
    STG_HD void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t h0, l0, h1, l1;
            mulhilo(0xD2511F53u, c0, h0, l0);
            mulhilo(0xCD9E8D57u, c2, h1, l1);
            uint32_t n0 = h1 ^ c1 ^ a;
            uint32_t n2 = h0 ^ c3 ^ b;
            c0 = n0; c1 = l1; c2 = n2; c3 = l0;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// two uniforms -> two N(0,1) (Box-Muller, single precision; the noise only has to be statistically N(0,1))
STG_HD void box_muller(uint32_t u0, uint32_t u1, float& n0, float& n1) {
    const float two_m32 = 2.3283064365386963e-10f;
    float a = fmaf((float)u0, two_m32, 0.5f * two_m32);   // (0, 1]
    float ang = (float)u1 * (two_m32 * 6.283185307179586f);
#if defined(__CUDA_ARCH__)
    float r = sqrtf(-2.0f * __logf(a));
    float s, c;
    __sincosf(ang, &s, &c);
#else
    float r = sqrtf(-2.0f * logf(a));
    float s = sinf(ang), c = cosf(ang);
#endif
    n0 = r * c;
    n1 = r * s;
}

// 12 normals for the 4 stages of RK4 substep `sub` of env-step `step` of env `gid` (3 Philox calls)
STG_HD void philox_normals12(const Philox& ph, uint64_t gid, uint32_t step, uint32_t sub, float xi[12]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        uint32_t o[4];
        ph((uint32_t)gid, (uint32_t)(gid >> 32), step, sub * 4u + (uint32_t)c, o);
        box_muller(o[0], o[1], xi[4 * c + 0], xi[4 * c + 1]);
        box_muller(o[2], o[3], xi[4 * c + 2], xi[4 * c + 3]);
    }
}
STG_HD void philox_normals4(const Philox& ph, uint64_t gid, uint32_t step, uint32_t sub, uint32_t lane, float xi[4]) {
    uint32_t o[4];
    ph((uint32_t)gid, (uint32_t)(gid >> 32), step, sub * 4u + lane, o);
    box_muller(o[0], o[1], xi[0], xi[1]);
    box_muller(o[2], o[3], xi[2], xi[3]);
}

// ---- per-step constants of one env (everything the substep loop needs, in the arithmetic type R) ---------------------
template <typename R>
struct StepConsts {
    R ck;        // G*hk
    R cd;        // -G*Ms            (demag  -Ms*m_z z^)
    R cth;       // G*h_th
    R alpha;
    R a_on;      // aJ*dt while the pulse is on
    R ex, ey, ez;
    R bax, bay, baz;  // G*H_app
};

// One stage: k = dt*f(m). AXIS_Z: easy axis == z^ exactly and H_app == 0, so every structurally-zero term is dropped.
template <typename R, bool AXIS_Z, bool THERMAL>
STG_HD void stage(const StepConsts<R>& c, R mx, R my, R mz, R a, R xx, R xy, R xz, R& kx, R& ky, R& kz) {
    if (AXIS_Z) {
        if (!THERMAL) {
            R bz = (c.ck + c.cd) * mz;
            R px = my * bz;
            R py = -(mx * bz);
            R qx = -(mz * py);
            R qy = mz * px;
            R qz = mx * py - my * px;
            R tx = mz * mx;
            R ty = mz * my;
            R tz = -(mx * mx + my * my);
            kx = px + c.alpha * qx + a * tx;
            ky = py + c.alpha * qy + a * ty;
            kz = c.alpha * qz + a * tz;
        } else {
            R bx = c.cth * xx;
            R by = c.cth * xy;
            R bz = (c.ck + c.cd) * mz + c.cth * xz;
            R px = my * bz - mz * by;
            R py = mz * bx - mx * bz;
            R pz = mx * by - my * bx;
            R qx = my * pz - mz * py;
            R qy = mz * px - mx * pz;
            R qz = mx * py - my * px;
            R tx = mz * mx;
            R ty = mz * my;
            R tz = -(mx * mx + my * my);
            kx = px + c.alpha * qx + a * tx;
            ky = py + c.alpha * qy + a * ty;
            kz = pz + c.alpha * qz + a * tz;
        }
    } else {
        R s = mx * c.ex + my * c.ey + mz * c.ez;
        R cks = c.ck * s;
        R bx = cks * c.ex + c.bax;
        R by = cks * c.ey + c.bay;
        R bz = cks * c.ez + c.baz + c.cd * mz;
        if (THERMAL) {
            bx += c.cth * xx;
            by += c.cth * xy;
            bz += c.cth * xz;
        }
        R px = my * bz - mz * by;
        R py = mz * bx - mx * bz;
        R pz = mx * by - my * bx;
        R qx = my * pz - mz * py;
        R qy = mz * px - mx * pz;
        R qz = mx * py - my * px;
        R ux = my * c.ez - mz * c.ey;
        R uy = mz * c.ex - mx * c.ez;
        R uz = mx * c.ey - my * c.ex;
        R tx = my * uz - mz * uy;
        R ty = mz * ux - mx * uz;
        R tz = mx * uy - my * ux;
        kx = px + c.alpha * qx + a * tx;
        ky = py + c.alpha * qy + a * ty;
        kz = pz + c.alpha * qz + a * tz;
    }
}

// reciprocal norm of an FP64 vector whose norm^2 is n2
template <typename R>
STG_HD double inv_norm(double n2);
template <>
STG_HD double inv_norm<double>(double n2) {
    return 1.0 / sqrt(n2);
}
template <>
STG_HD double inv_norm<float>(double n2) {
    // f32 seed + one FP64 Newton step: relative error ~1e-14, far below the f32 stage arithmetic
#if defined(__CUDA_ARCH__)
    double y = (double)rsqrtf((float)n2);
#else
    double y = (double)(1.0f / sqrtf((float)n2));
#endif
    return y * (1.5 - 0.5 * n2 * y * y);
}

// Guard + normalise (physics/simple_solver.py:208-229): non-finite or |m| < 1e-12 -> (0,0,1) and the guard flag.
template <typename R>
STG_HD void guard_normalise(double& mx, double& my, double& mz, int& guard) {
    double n2 = mx * mx + my * my + mz * mz;
    if (n2 >= 1e-24 && n2 <= 1.0e300) {
        double inv = inv_norm<R>(n2);
        mx *= inv; my *= inv; mz *= inv;
    } else {
        mx = 0.0; my = 0.0; mz = 1.0;
        guard = 1;
    }
}

// One fixed-step substep on the FP64 state. a1..a4: aJ*dt seen by the four stages (pulse gating). xi: 12 (rk4) or 3 (euler)
// N(0,1) samples in stage order.
template <typename R, bool AXIS_Z, bool THERMAL, bool EULER>
STG_HD void substep(const StepConsts<R>& c, double& mdx, double& mdy, double& mdz, R a1, R a2, R a3, R a4, const R* xi,
                    int& guard) {
    R mx = (R)mdx, my = (R)mdy, mz = (R)mdz;
    R k1x, k1y, k1z;
    stage<R, AXIS_Z, THERMAL>(c, mx, my, mz, a1, THERMAL ? xi[0] : R(0), THERMAL ? xi[1] : R(0), THERMAL ? xi[2] : R(0),
                              k1x, k1y, k1z);
    R ix, iy, iz;
    if (EULER) {
        ix = k1x; iy = k1y; iz = k1z;
    } else {
        const R h = R(0.5);
        R k2x, k2y, k2z, k3x, k3y, k3z, k4x, k4y, k4z;
        stage<R, AXIS_Z, THERMAL>(c, mx + h * k1x, my + h * k1y, mz + h * k1z, a2, THERMAL ? xi[3] : R(0),
                                  THERMAL ? xi[4] : R(0), THERMAL ? xi[5] : R(0), k2x, k2y, k2z);
        stage<R, AXIS_Z, THERMAL>(c, mx + h * k2x, my + h * k2y, mz + h * k2z, a3, THERMAL ? xi[6] : R(0),
                                  THERMAL ? xi[7] : R(0), THERMAL ? xi[8] : R(0), k3x, k3y, k3z);
        stage<R, AXIS_Z, THERMAL>(c, mx + k3x, my + k3y, mz + k3z, a4, THERMAL ? xi[9] : R(0), THERMAL ? xi[10] : R(0),
                                  THERMAL ? xi[11] : R(0), k4x, k4y, k4z);
        const R sixth = R(1.0 / 6.0);
        ix = (k1x + R(2) * (k2x + k3x) + k4x) * sixth;
        iy = (k1y + R(2) * (k2y + k3y) + k4y) * sixth;
        iz = (k1z + R(2) * (k2z + k3z) + k4z) * sixth;
    }
    mdx += (double)ix;
    mdy += (double)iy;
    mdz += (double)iz;
    guard_normalise<R>(mdx, mdy, mdz, guard);
}

// ---- FP32 stages with a block-scaled transverse state (easy axis == z^, no thermal field) ------------------------------
// Without noise the magnetisation converges onto a pole exponentially (transverse components reach 1e-100 and below within
// a few env steps) and the reference, being FP64, regrows them from there when the current reverses. A float cannot hold
// anything below ~1e-38, so the FP32 variant carries the transverse pair multiplied by a power of two S = 1/inv_s:
// every term of k_x, k_y is linear in (m_x, m_y) (so it comes out scaled by S for free) and every term of k_z is quadratic
// (so it is multiplied back by inv_s^2, which simply underflows to 0 when the pair is negligible against m_z = +-1).
// With inv_s == 1 the arithmetic is identical to substep<float, true, false, EULER>.
struct ScaledState {
    double sx, sy, z;      // S*m_x, S*m_y, m_z
    double inv_s;          // 1/S, a power of two <= 1
    double inv_s2d;        // inv_s^2 (FP64; may underflow to 0)
    float inv_s2f;         // inv_s^2 (FP32; may underflow to 0)
};

STG_HD double pow2(int e) {   // 2^e for e in [-1022, 1023]
#if defined(__CUDA_ARCH__)
    return __hiloint2double((e + 1023) << 20, 0);
#else
    return ldexp(1.0, e);
#endif
}
STG_HD int exponent_of(double t) {   // floor(log2(t)) for a normal positive double
#if defined(__CUDA_ARCH__)
    return ((__double2hiint(t) >> 20) & 0x7ff) - 1023;
#else
    int e;
    frexp(t, &e);
    return e - 1;
#endif
}
// Keep max(|sx|,|sy|) within [2^-60, 2^20] by moving powers of two between the pair and inv_s (never below S = 1).
STG_HD void rescale(ScaledState& st) {
    const double t = fmax(fabs(st.sx), fabs(st.sy));
    if (!(t > 0.0) || !(t < 1.0e300)) return;
    int shift = 0;
    if (t < 8.673617379884035e-19) {                 // 2^-60: scale up
        const int ex = t < 2.2250738585072014e-308 ? -1022 : exponent_of(t);
        shift = -10 - ex;
        if (shift > 900) shift = 900;
    } else if (t > 1048576.0 && st.inv_s < 1.0) {    // grew back: scale down, but not below S = 1
        const int ex = exponent_of(t);
        shift = -10 - ex;
        const int have = -exponent_of(st.inv_s);     // log2(S)
        if (-shift > have) shift = -have;
    }
    if (shift == 0) return;
    const int cur = -exponent_of(st.inv_s);
    int tot = cur + shift;
    if (tot > 1000) { shift -= tot - 1000; tot = 1000; }
    const double up = pow2(shift);
    st.sx *= up;
    st.sy *= up;
    st.inv_s = pow2(-tot);
    st.inv_s2d = (tot <= 511) ? pow2(-2 * tot) : 0.0;
    st.inv_s2f = (float)st.inv_s2d;
}

template <bool EULER>
STG_HD void substep_scaled(const StepConsts<float>& c, ScaledState& st, float a1, float a2, float a3, float a4, int& guard) {
    const float mx = (float)st.sx, my = (float)st.sy, mz = (float)st.z;
    const float q = st.inv_s2f;
    float k1x, k1y, k1z;
    stage<float, true, false>(c, mx, my, mz, a1, 0.f, 0.f, 0.f, k1x, k1y, k1z);
    k1z *= q;
    float ix, iy, iz;
    if (EULER) {
        ix = k1x; iy = k1y; iz = k1z;
    } else {
        float k2x, k2y, k2z, k3x, k3y, k3z, k4x, k4y, k4z;
        stage<float, true, false>(c, mx + 0.5f * k1x, my + 0.5f * k1y, mz + 0.5f * k1z, a2, 0.f, 0.f, 0.f, k2x, k2y, k2z);
        k2z *= q;
        stage<float, true, false>(c, mx + 0.5f * k2x, my + 0.5f * k2y, mz + 0.5f * k2z, a3, 0.f, 0.f, 0.f, k3x, k3y, k3z);
        k3z *= q;
        stage<float, true, false>(c, mx + k3x, my + k3y, mz + k3z, a4, 0.f, 0.f, 0.f, k4x, k4y, k4z);
        k4z *= q;
        const float sixth = 1.0f / 6.0f;
        ix = (k1x + 2.0f * (k2x + k3x) + k4x) * sixth;
        iy = (k1y + 2.0f * (k2y + k3y) + k4y) * sixth;
        iz = (k1z + 2.0f * (k2z + k3z) + k4z) * sixth;
    }
    st.sx += (double)ix;
    st.sy += (double)iy;
    st.z += (double)iz;
    const double n2 = st.z * st.z + (st.sx * st.sx + st.sy * st.sy) * st.inv_s2d;
    if (n2 >= 1e-24 && n2 <= 1.0e300) {
        const double inv = inv_norm<float>(n2);
        st.sx *= inv; st.sy *= inv; st.z *= inv;
    } else {
        st.sx = 0.0; st.sy = 0.0; st.z = 1.0;
        st.inv_s = 1.0; st.inv_s2d = 1.0; st.inv_s2f = 1.0f;
        guard = 1;
    }
}

// ---- step plan (physics/simple_solver.py:137-139), evaluated exactly like NumPy does in FP64 -------------------------
struct StepPlan {
    int n;
    double dt;
};
STG_HD double ddiv(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
STG_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
STG_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
STG_HD StepPlan substep_plan(double t_end, double max_step) {
    double dt0 = ddiv(t_end, 100.0);
    if (max_step < dt0) dt0 = max_step;
    double q = ddiv(t_end, dt0);
    int n = (q < 2.0e9) ? (int)q : 2000000000;
    if (n < 10) n = 10;
    StepPlan p;
    p.n = n;
    p.dt = ddiv(t_end, (double)n);
    return p;
}
// current_func(t) = J if t <= t_pulse else 0 (envs/spin_torque_env.py:442-443) at t = t_i + frac*dt with
// t_i = i*dt as np.linspace produces it (physics/simple_solver.py:142); frac in {0, 0.5, 1}.
STG_HD bool pulse_on(int i, int stage_kind, double dt, double t_pulse) {
    double ti = dmul((double)i, dt);
    double t = ti;
    if (stage_kind == 1) t = dadd(ti, dmul(dt, 0.5));   // dt/2 is exact
    if (stage_kind == 2) t = dadd(ti, dt);
    return t <= t_pulse;
}

// ---- resistance (devices/stt_mram.py:78-94, devices/sot_mram.py:196-228, devices/vcma_mram.py:236-257) ---------------
STG_HD double resistance(const double* f, double mx, double my, double mz) {
    int kind = (int)f[FI_KIND];
    double rp = f[FI_RP], rap = f[FI_RAP];
    if (kind == 0) {
        double inv = 1.0 / sqrt(mx * mx + my * my + mz * mz);   // validate_magnetization renormalises (base_device.py:94-116)
        double c = (mx * f[FI_REFX] + my * f[FI_REFY] + mz * f[FI_REFZ]) * inv;
        double r = rp * (1.0 + f[FI_TMR] * (1.0 - c) / 2.0);
        double lo = rp * 0.5;
        return r > lo ? r : lo;
    }
    double c = mx * f[FI_REFX] + my * f[FI_REFY] + mz * f[FI_REFZ];
    double r = rp + (rap - rp) * (1.0 - c) / 2.0 + f[FI_RSERIES];
    return r > 1.0 ? r : 1.0;
}

}  // namespace stg

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

#ifndef STG_FAST_COPY_FROM_MASTER
#define STG_FAST_COPY_FROM_MASTER 1
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
#ifndef STG_PHILOX_ROUNDS
#define STG_PHILOX_ROUNDS 10
#endif
        for (int r = 0; r < STG_PHILOX_ROUNDS; ++r) {
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

// ---- fast device intrinsics (MUFU) with libm fallbacks for the host build ---------------------------------------------
STG_HD float fast_lg2(float x) {
#if defined(__CUDA_ARCH__) && defined(STG_EXP_NO_LGSQRT)
    return x - 1.0f;                    // timing experiment only
#elif defined(__CUDA_ARCH__)
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // x >= 2^-33 here: no denormal fix-up needed
    return r;
#else
    return log2f(x);
#endif
}
STG_HD float fast_sqrt(float x) {
#if defined(__CUDA_ARCH__) && defined(STG_EXP_NO_LGSQRT)
    return x * 0.5f;                    // timing experiment only
#elif defined(__CUDA_ARCH__)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return sqrtf(x);
#endif
}
STG_HD void fast_sincos(float ang, float& s, float& c) {   // angle in radians, [0, 2 pi)
#if defined(__CUDA_ARCH__) && defined(STG_EXP_NO_SINCOS)
    s = ang * 0.1f; c = 1.0f - s;      // timing experiment only
#elif defined(__CUDA_ARCH__)
    s = __sinf(ang);
    c = __cosf(ang);
#else
    s = sinf(ang);
    c = cosf(ang);
#endif
}
STG_HD float bits_to_angle(uint32_t x) {   // top 23 bits -> [0, 2 pi) without an int->float conversion (one FFMA)
    const float twopi = 6.283185307179586f;
#if defined(__CUDA_ARCH__)
    return fmaf(__uint_as_float(0x3f800000u | (x >> 9)), twopi, -twopi);
#else
    union { uint32_t u; float f; } v;
    v.u = 0x3f800000u | (x >> 9);
    return fmaf(v.f, twopi, -twopi);
#endif
}

// two uniforms -> two N(0, scale^2) (Box-Muller in single precision; the noise only has to be statistically Gaussian).
// The radius uses all 32 bits of u0 (tail out to 6.66 sigma), the angle the top 23 bits of u1. `neg2ln2_scale2` is
// -2*ln(2)*scale^2 so that scaling the field strength into the samples costs nothing.
STG_HD void box_muller_scaled(uint32_t u0, uint32_t u1, float neg2ln2_scale2, float& n0, float& n1) {
    const float two_m32 = 2.3283064365386963e-10f;
    const float a = fmaf((float)u0, two_m32, 0.5f * two_m32);   // (0, 1]
    const float r = fast_sqrt(neg2ln2_scale2 * fast_lg2(a));
    float s, c;
    fast_sincos(bits_to_angle(u1), s, c);
    n0 = r * c;
    n1 = r * s;
}
STG_HD void box_muller(uint32_t u0, uint32_t u1, float& n0, float& n1) {
    box_muller_scaled(u0, u1, -1.3862943611198906f, n0, n1);
}

// ---- thermal-field stream identity ---------------------------------------------------------------------------------------------
// (the RK4 paths seed a xoshiro128++ state from block 0 of this stream, see ThermalSource below; Euler and the resets draw
// their samples from the blocks themselves)
// Philox4x32-10, key = the 64-bit seed (uniform over a launch: the key schedule lives in uniform registers), counter =
// (global env id low word, episode, env step, block index | global env id bits 32..42 << 20): every (env, episode, step, block)
// owns one counter value, independent of how the batch is partitioned over launches or GPUs (global ids < 2^43, checked by the
// entry points). Only the last counter word changes inside a step, so the first Philox round and half of the second are
// loop-invariant.
//
// Bit budget of the RK4 path: one 32-bit word per Box-Muller pair (16-bit radius uniform, tail to 4.85 sigma; 16-bit angle =
// 65,536 directions): six words -> six pairs -> the 12 normals of one substep (4 stages x 3 components, drawn in the reference's
// order: substep, stage, xyz). Round 1 spent two Philox blocks per substep (23-bit radius + 19-bit angle), see profiles/README.md.
struct NoiseStream {
    Philox ph;
    uint32_t c0, c1, c2, c3;      // c3: (gid >> 32) << 20, the block index is added to it
};
STG_HD NoiseStream make_stream(uint64_t seed, uint64_t gid, uint32_t episode, uint32_t step) {
    NoiseStream s;
    s.ph = Philox{(uint32_t)seed, (uint32_t)(seed >> 32)};
    s.c0 = (uint32_t)gid; s.c1 = episode; s.c2 = step; s.c3 = (uint32_t)(gid >> 32) << 20;
    return s;
}
STG_HD float bits_as_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { uint32_t u; float f; } v;
    v.u = u;
    return v.f;
#endif
}
// Euler: 3 samples per substep from its own block (full 32-bit uniforms; not a hot path)
STG_HD void philox_normals3(const NoiseStream& ns, uint32_t sub, float neg2ln2_scale2, float xi[4]) {
    uint32_t o[4];
    ns.ph(ns.c0, ns.c1, ns.c2, ns.c3 + sub, o);
    box_muller_scaled(o[0], o[1], neg2ln2_scale2, xi[0], xi[1]);
    box_muller_scaled(o[2], o[3], neg2ln2_scale2, xi[2], xi[3]);
}

// ---- per-step constants of one env --------------------------------------------------------------------------------------
// Every product of physical constants and dt is formed in FP64 and split into a (hi, lo) pair of R so that FP32 stages see
// the constants to ~48 bits: a constant rounded to 24 bits is a SYSTEMATIC error of the precession / relaxation rate that
// grows linearly (phase) or quadratically (phase through m_z) with the 10..5000 substeps of a step. For R = double lo == 0.
// `sc` pre-scales every increment (1/6 for the RK4 fast path, so the RK4 weights become exact small integers).
template <typename R>
struct StepConsts {
    R c_hi, c_lo;       // sc * G*(hk - Ms)          axis-z:  b_z = c * m_z
    R ac_hi, ac_lo;     // sc * alpha*G*(hk - Ms)     axis-z:  w = ac*m_z + a
    R a_hi, a_lo;       // sc * aJ*dt (pulse on)
    R al_hi, al_lo;     // alpha
    R ck_hi, ck_lo;     // sc * G*hk                  general axis
    R cd_hi, cd_lo;     // sc * -G*Ms
    R cth;              // sc * G*h_th  (field strength folded into the noise samples)
    R ex, ey, ez;
    R bax, bay, baz;    // sc * G*H_app
};

template <typename R>
STG_HD void split_const(double x, R& hi, R& lo) {
    hi = (R)x;
    lo = (sizeof(R) == 4) ? (R)(x - (double)hi) : R(0);
}

template <typename R>
STG_HD void make_consts(const double* f, double dt, double J, double sc, StepConsts<R>& c) {
    const double G = -f[FI_GEFF] * dt * sc;
    const double alpha = f[FI_ALPHA];
    split_const<R>(G * (f[FI_HK] - f[FI_MS]), c.c_hi, c.c_lo);
    split_const<R>(alpha * G * (f[FI_HK] - f[FI_MS]), c.ac_hi, c.ac_lo);
    const double a = (fabs(J) > 1e-12) ? f[FI_AJ_PER_J] * J * dt * sc : 0.0;   // physics/simple_solver.py:327-331
    split_const<R>(a, c.a_hi, c.a_lo);
    split_const<R>(alpha, c.al_hi, c.al_lo);
    split_const<R>(G * f[FI_HK], c.ck_hi, c.ck_lo);
    split_const<R>(-G * f[FI_MS], c.cd_hi, c.cd_lo);
    c.cth = (R)(G * f[FI_HTH]);
    c.ex = (R)f[FI_EX]; c.ey = (R)f[FI_EY]; c.ez = (R)f[FI_EZ];
    c.bax = (R)(G * f[FI_HAX]); c.bay = (R)(G * f[FI_HAY]); c.baz = (R)(G * f[FI_HAZ]);
}

// ---- stage derivative, easy axis == z^ and H_app == 0 (deterministic part) ------------------------------------------------
// With e = z^ the reference's  m x b + alpha m x (m x b) + a m x (m x e),  b = (0, 0, c m_z),  collapses to
//     w = alpha*c*m_z + a,   k_x = m_y b_z + m_x m_z w,   k_y = -m_x b_z + m_y m_z w,   k_z = -(m_x^2 + m_y^2) w
// (13 FMA-pipe instructions). q multiplies k_z only for the block-scaled transverse state (ScaledState below).
template <typename R, bool SCALED>
STG_HD void stage_z(const StepConsts<R>& c, R mx, R my, R mz, R aH, R aL, R q, R& kx, R& ky, R& kz) {
    R bz, w;
    if (sizeof(R) == 4) {
        bz = c.c_lo * mz + c.c_hi * mz;
        w = (c.ac_hi * mz + aH) + (c.ac_lo * mz + aL);
    } else {
        bz = c.c_hi * mz;
        w = c.ac_hi * mz + aH;
    }
    const R u = mz * w;
    const R s2 = mx * mx + my * my;
    kx = mx * u + my * bz;
    ky = my * u - mx * bz;
    kz = SCALED ? (s2 * -w) * q : s2 * -w;
}
// thermal part for e = z^: k += m x bn + alpha m x (m x bn), bn = (scaled) noise rotation vector
template <typename R>
STG_HD void stage_noise(const StepConsts<R>& c, R mx, R my, R mz, R bx, R by, R bz, R& kx, R& ky, R& kz) {
    const R px = my * bz - mz * by;
    const R py = mz * bx - mx * bz;
    const R pz = mx * by - my * bx;
    const R qx = my * pz - mz * py;
    const R qy = mz * px - mx * pz;
    const R qz = mx * py - my * px;
    kx += px + c.al_hi * qx;
    ky += py + c.al_hi * qy;
    kz += pz + c.al_hi * qz;
}

// ---- stage derivative, general easy axis / applied field (cross-product form of the reference) ----------------------------
template <typename R, bool THERMAL>
STG_HD void stage_general(const StepConsts<R>& c, R mx, R my, R mz, R aH, R aL, R nx, R ny, R nz, R& kx, R& ky, R& kz) {
    const R s = mx * c.ex + my * c.ey + mz * c.ez;
    R cks, cdz;
    if (sizeof(R) == 4) {
        cks = c.ck_lo * s + c.ck_hi * s;
        cdz = c.cd_lo * mz + c.cd_hi * mz;
    } else {
        cks = c.ck_hi * s;
        cdz = c.cd_hi * mz;
    }
    R bx = cks * c.ex + c.bax;
    R by = cks * c.ey + c.bay;
    R bz = cks * c.ez + c.baz + cdz;
    if (THERMAL) { bx += nx; by += ny; bz += nz; }
    const R px = my * bz - mz * by;
    const R py = mz * bx - mx * bz;
    const R pz = mx * by - my * bx;
    const R qx = my * pz - mz * py;
    const R qy = mz * px - mx * pz;
    const R qz = mx * py - my * px;
    const R ux = my * c.ez - mz * c.ey;
    const R uy = mz * c.ex - mx * c.ez;
    const R uz = mx * c.ey - my * c.ex;
    const R tx = my * uz - mz * uy;
    const R ty = mz * ux - mx * uz;
    const R tz = mx * uy - my * ux;
    if (sizeof(R) == 4) {
        kx = px + (c.al_hi * qx + (aH * tx + (aL * tx + c.al_lo * qx)));
        ky = py + (c.al_hi * qy + (aH * ty + (aL * ty + c.al_lo * qy)));
        kz = pz + (c.al_hi * qz + (aH * tz + (aL * tz + c.al_lo * qz)));
    } else {
        kx = px + c.al_hi * qx + aH * tx;
        ky = py + c.al_hi * qy + aH * ty;
        kz = pz + c.al_hi * qz + aH * tz;
    }
}

// explicit FP64 products / sums / fused multiply-adds: the compiler may not contract or split them, so the same env gets the same
// bits from every kernel and from either half of a two-envs-per-thread pack (definitions below)
STG_HD double dmul(double a, double b);
STG_HD double dadd(double a, double b);
STG_HD double dfma(double a, double b, double c) { return fma(a, b, c); }

// reciprocal norm of an FP64 vector whose norm^2 is n2
template <typename R>
STG_HD double inv_norm(double n2);
template <>
STG_HD double inv_norm<double>(double n2) {
    return 1.0 / sqrt(n2);
}
template <>
STG_HD double inv_norm<float>(double n2) {
    // f32 seed + two FP64 Newton steps (relative error ~1e-14 after the first)
#if defined(__CUDA_ARCH__)
    double y = (double)rsqrtf((float)n2);
#else
    double y = (double)(1.0f / sqrtf((float)n2));
#endif
    y = dmul(y, dfma(-0.5, dmul(dmul(n2, y), y), 1.5));
    return dmul(y, dfma(-0.5, dmul(dmul(n2, y), y), 1.5));
}

// ---- block-scaled state ----------------------------------------------------------------------------------------------------
// Without noise the magnetisation converges onto a pole exponentially (transverse components reach 1e-100 and below within
// a few env steps) and the reference, being FP64, regrows them from there when the current reverses. A float cannot hold
// anything below ~1e-38, so the FP32 axis-z variants carry the transverse pair multiplied by a power of two S = 1/inv_s:
// every term of k_x, k_y is linear in (m_x, m_y) (it comes out scaled by S for free) and k_z is quadratic (it is multiplied
// back by inv_s^2, which simply underflows to 0 when the pair is negligible against m_z = +-1). With inv_s == 1 nothing
// changes. Every other variant keeps inv_s == 1.
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

// Guard + normalise of the FP64 master (physics/simple_solver.py:208-229): non-finite or |m| < 1e-12 -> (0,0,1) + guard flag.
template <typename R>
STG_HD void guard_normalise(ScaledState& st, int& guard) {
    const double n2 = dfma(st.z, st.z, dmul(dfma(st.sx, st.sx, dmul(st.sy, st.sy)), st.inv_s2d));
    if (n2 >= 1e-24 && n2 <= 1.0e300) {
        const double inv = inv_norm<R>(n2);
        st.sx *= inv; st.sy *= inv; st.z *= inv;
    } else {
        st.sx = 0.0; st.sy = 0.0; st.z = 1.0;
        st.inv_s = 1.0; st.inv_s2d = 1.0; st.inv_s2f = 1.0f;
        guard = 1;
    }
}
template <typename R>
STG_HD void guard_normalise(double& mx, double& my, double& mz, int& guard) {
    ScaledState st{mx, my, mz, 1.0, 1.0, 1.0f};
    guard_normalise<R>(st, guard);
    mx = st.sx; my = st.sy; mz = st.z;
}

// ---- reference-structure substep: FP64 master renormalised exactly after every substep ----------------------------------
// Used by the FP64 variants, the general-axis variants and Euler. aH/aL: aJ*dt (hi, lo) seen by the stages (pulse gating:
// [0] stage 1, [1] stages 2 and 3, [2] stage 4). nz: 12 (rk4) / 3 (euler) noise rotation components, already scaled by cth.
template <typename R, bool AXIS_Z, bool THERMAL, bool EULER>
STG_HD void substep_ref(const StepConsts<R>& c, ScaledState& st, const R* aH, const R* aL, const R* nz, int& guard) {
    constexpr bool SCALED = AXIS_Z && !THERMAL && sizeof(R) == 4;
    const R mx = (R)st.sx, my = (R)st.sy, mz = (R)st.z;
    const R q = SCALED ? (R)st.inv_s2f : R(1);
    auto f = [&](R x, R y, R z, int g, int s, R& kx, R& ky, R& kz) {
        if (AXIS_Z) {
            stage_z<R, SCALED>(c, x, y, z, aH[g], aL[g], q, kx, ky, kz);
            if (THERMAL) stage_noise<R>(c, x, y, z, nz[3 * s], nz[3 * s + 1], nz[3 * s + 2], kx, ky, kz);
        } else {
            stage_general<R, THERMAL>(c, x, y, z, aH[g], aL[g], THERMAL ? nz[3 * s] : R(0), THERMAL ? nz[3 * s + 1] : R(0),
                                      THERMAL ? nz[3 * s + 2] : R(0), kx, ky, kz);
        }
    };
    R k1x, k1y, k1z;
    f(mx, my, mz, 0, 0, k1x, k1y, k1z);
    R ix, iy, iz;
    if (EULER) {
        ix = k1x; iy = k1y; iz = k1z;
    } else {
        const R h = R(0.5);
        R k2x, k2y, k2z, k3x, k3y, k3z, k4x, k4y, k4z;
        f(mx + h * k1x, my + h * k1y, mz + h * k1z, 1, 1, k2x, k2y, k2z);
        f(mx + h * k2x, my + h * k2y, mz + h * k2z, 1, 2, k3x, k3y, k3z);
        f(mx + k3x, my + k3y, mz + k3z, 2, 3, k4x, k4y, k4z);
        ix = k1x + R(2) * (k2x + k3x) + k4x;
        iy = k1y + R(2) * (k2y + k3y) + k4y;
        iz = k1z + R(2) * (k2z + k3z) + k4z;
    }
    const double w = EULER ? 1.0 : (1.0 / 6.0);
    st.sx += w * (double)ix;
    st.sy += w * (double)iy;
    st.z += w * (double)iz;
    guard_normalise<R>(st, guard);
}

// ---- FP32 lane packs --------------------------------------------------------------------------------------------------
// The fast path is written once over a pack type P: `float` (one env per thread) or `F2` (two envs per thread in the two
// halves of a 64-bit register pair, executed by Blackwell's packed FFMA2 / FMUL2 / FADD2: one issue slot for two FMAs).
// Every operation is an explicit IEEE fma / mul / add per component, so both packs produce bit-identical results.
#if defined(__CUDA_ARCH__)
typedef float2 F2;
#else
struct F2 { float x, y; };
#endif
STG_HD F2 mk2(float a, float b) {
#if defined(__CUDA_ARCH__)
    return make_float2(a, b);
#else
    return F2{a, b};
#endif
}
template <typename P> struct Pk;
template <> struct Pk<float> {
    static STG_HD float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static STG_HD float mul(float a, float b) {
#if defined(__CUDA_ARCH__)
        return __fmul_rn(a, b);
#else
        volatile float r = a * b; return r;
#endif
    }
    static STG_HD float add(float a, float b) {
#if defined(__CUDA_ARCH__)
        return __fadd_rn(a, b);
#else
        volatile float r = a + b; return r;
#endif
    }
    static STG_HD float neg(float a) { return -a; }
    static STG_HD float bc(float v) { return v; }
};
template <> struct Pk<F2> {
    static STG_HD F2 fma(F2 a, F2 b, F2 c) {
#if defined(__CUDA_ARCH__)
        return __ffma2_rn(a, b, c);
#else
        return F2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)};
#endif
    }
    static STG_HD F2 mul(F2 a, F2 b) {
#if defined(__CUDA_ARCH__)
        return __fmul2_rn(a, b);
#else
        volatile float x = a.x * b.x, y = a.y * b.y; return F2{x, y};
#endif
    }
    static STG_HD F2 add(F2 a, F2 b) {
#if defined(__CUDA_ARCH__)
        return __fadd2_rn(a, b);
#else
        volatile float x = a.x + b.x, y = a.y + b.y; return F2{x, y};
#endif
    }
    static STG_HD F2 neg(F2 a) { return mk2(-a.x, -a.y); }   // folded into the operand modifier of the consuming FFMA2
    static STG_HD F2 bc(float v) { return mk2(v, v); }
};

// K independent F2 packs per thread (2K envs): every pack operation becomes K independent packed instructions, which gives the
// in-order issue of one warp K dependent chains to interleave (the stage chain of one pack is ~5 cycles per instruction deep).
template <int K>
struct FN {
    F2 p[K];
};
template <int K> struct Pk<FN<K>> {
    typedef FN<K> T;
    static STG_HD T fma(const T& a, const T& b, const T& c) {
        T r;
#pragma unroll
        for (int i = 0; i < K; ++i) r.p[i] = Pk<F2>::fma(a.p[i], b.p[i], c.p[i]);
        return r;
    }
    static STG_HD T mul(const T& a, const T& b) {
        T r;
#pragma unroll
        for (int i = 0; i < K; ++i) r.p[i] = Pk<F2>::mul(a.p[i], b.p[i]);
        return r;
    }
    static STG_HD T add(const T& a, const T& b) {
        T r;
#pragma unroll
        for (int i = 0; i < K; ++i) r.p[i] = Pk<F2>::add(a.p[i], b.p[i]);
        return r;
    }
    static STG_HD T neg(const T& a) {
        T r;
#pragma unroll
        for (int i = 0; i < K; ++i) r.p[i] = Pk<F2>::neg(a.p[i]);
        return r;
    }
    static STG_HD T bc(float v) {
        T r;
#pragma unroll
        for (int i = 0; i < K; ++i) r.p[i] = mk2(v, v);
        return r;
    }
};
// lane access of a pack (compile-time lane index after unrolling)
template <typename P> struct Ln;
template <> struct Ln<float> {
    static constexpr int N = 1;
    static STG_HD float get(float v, int) { return v; }
    static STG_HD void set(float& v, int, float x) { v = x; }
};
template <> struct Ln<F2> {
    static constexpr int N = 2;
    static STG_HD float get(const F2& v, int l) { return l ? v.y : v.x; }
    static STG_HD void set(F2& v, int l, float x) { if (l) v.y = x; else v.x = x; }
};
template <int K> struct Ln<FN<K>> {
    static constexpr int N = 2 * K;
    static STG_HD float get(const FN<K>& v, int l) { return Ln<F2>::get(v.p[l >> 1], l & 1); }
    static STG_HD void set(FN<K>& v, int l, float x) { Ln<F2>::set(v.p[l >> 1], l & 1, x); }
};

// ---- thermal-field samples over a pack -----------------------------------------------------------------------------------------
// P = float: one env per thread; F2: two envs per thread with one stream per lane. The integer Philox rounds and the MUFU
// evaluations run per lane, the FP32 arithmetic of Box-Muller on whole packs (FFMA2 / FMUL2); lane l of xi[k] is sample k of
// stream l, bit-identical to what the one-env form produces for that stream.
//
// Box-Muller pair from ONE 32-bit word per lane: radius uniform U = (hi16 + 1/2) / 65536 in (0, 1), angle = 2 pi lo16 / 65536.
// Both are formed by dropping the 16 bits into the mantissa of 1.0f (one PRMT / LOP3) and one exact FFMA, no int->float
// conversion. `nscale` = -2 ln(2) scale^2 folds the field strength into the radius.
STG_HD uint32_t lo16_over_one(uint32_t w) {       // 0x3f800000 | (w & 0xffff) as ONE byte permute (the AND + OR are two LOP3)
#if defined(__CUDA_ARCH__)
    return __byte_perm(w, 0x3f800000u, 0x7610);
#else
    return 0x3f800000u | (w & 0xffffu);
#endif
}
template <typename P>
STG_HD void box_muller16(const uint32_t* w, P nscale, P& n0, P& n1) {
    using K = Pk<P>;
    using L = Ln<P>;
    P vr, va, lg, r, sn, cs;
#pragma unroll
    for (int l = 0; l < L::N; ++l) {
        L::set(vr, l, bits_as_float(0x3f800000u | (w[l] >> 16)));        // 1 + hi16 2^-23
        L::set(va, l, bits_as_float(lo16_over_one(w[l])));                // 1 + lo16 2^-23
    }
    const P U = K::fma(vr, K::bc(128.0f), K::bc(-127.99999237060546875f));     // 128 vr - (128 - 2^-17) = (hi16 + 1/2) 2^-16, exact
    const P ang = K::fma(va, K::bc(804.24771931898703f), K::bc(-804.24771931898703f));   // 2 pi 128 (va - 1), one rounding
#pragma unroll
    for (int l = 0; l < L::N; ++l) L::set(lg, l, fast_lg2(L::get(U, l)));
    const P t = K::mul(nscale, lg);
#pragma unroll
    for (int l = 0; l < L::N; ++l) {
        float s_, c_;
        fast_sincos(L::get(ang, l), s_, c_);
        L::set(r, l, fast_sqrt(L::get(t, l)));
        L::set(sn, l, s_);
        L::set(cs, l, c_);
    }
    n0 = K::mul(r, cs);
    n1 = K::mul(r, sn);
}
// one Philox block of every lane's stream: w[j][l] = word j of lane l
template <int NL>
STG_HD void philox_block(const NoiseStream* ns, uint32_t idx, uint32_t (*w)[NL]) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        uint32_t o[4];
        ns[l].ph(ns[l].c0, ns[l].c1, ns[l].c2, ns[l].c3 + idx, o);
        w[0][l] = o[0]; w[1][l] = o[1]; w[2][l] = o[2]; w[3][l] = o[3];
    }
}
// ---- the in-kernel thermal stream of the RK4 paths ---------------------------------------------------------------------------
// xoshiro128++ 1.0 (Blackman & Vigna, "Scrambled linear pseudorandom number generators", ACM TOMS 2021; passes BigCrush, period
// 2^128 - 1): 4 x 32-bit state per env, output rotl(s0 + s3, 7) + s0 - nine ALU-pipe instructions per 32-bit word and no
// multiplies, where a Philox4x32-10 word costs four IMAD.WIDE on the FMA-heavy pipe that the packed FP32 stage arithmetic also
// needs (profiles/README.md: 10.44 -> 8.83 ms per 1M-env x 999-substep step). The state of an env-step is SEEDED from block 0 of
// that env-step's Philox stream (key = seed, counter = (global env id, episode, step, 0)): the stream of an env-step is a pure
// function of (seed, global id, episode, step) - independent of sharding, launch geometry and thread mapping, like the all-Philox
// stream it replaces - and sequential only inside the step (substep i consumes words 6 i .. 6 i + 5, which is the order every
// integrator runs in). 2^31 env-steps x 2^13 words from random 128-bit starting points overlap with probability ~2^-53.
struct Xoshiro128pp {
    uint32_t s0, s1, s2, s3;
    static STG_HD uint32_t rotl(uint32_t x, int k) {
#if defined(__CUDA_ARCH__)
        return __funnelshift_l(x, x, k);
#else
        return (x << k) | (x >> (32 - k));
#endif
    }
    STG_HD uint32_t next() {
        const uint32_t r = rotl(s0 + s3, 7) + s0;
        const uint32_t t = s1 << 9;
        const uint32_t n2 = s2 ^ s0, n3 = s3 ^ s1;
        s1 ^= n2;
        s0 ^= n3;
        s2 = n2 ^ t;
        s3 = rotl(n3, 11);
        return r;
    }
};
STG_HD Xoshiro128pp seed_xoshiro(const NoiseStream& ns) {
    uint32_t o[4];
    ns.ph(ns.c0, ns.c1, ns.c2, ns.c3, o);
    if ((o[0] | o[1] | o[2] | o[3]) == 0u) o[0] = 1u;        // the all-zero state is the generator's one fixed point
    return Xoshiro128pp{o[0], o[1], o[2], o[3]};
}
// Noise source of the RK4 integrators, one stream per lane of the pack. first(g, nz) / second(g, nz): the 12 samples per lane
// (already scaled) of substep 2g / 2g+1, called strictly in that order from g = 0 (a trailing odd substep calls first() only).
// GEN 0: the xoshiro stream above (default). GEN 1 (STG_F_STREAM_PHILOX10): every word from Philox4x32-10, below.
template <typename P, int GEN = 0>
struct ThermalSource {
    Xoshiro128pp st[Ln<P>::N];
    P nscale;
    STG_HD void init(const NoiseStream* ns, P scale) {
        nscale = scale;
#pragma unroll
        for (int l = 0; l < Ln<P>::N; ++l) st[l] = seed_xoshiro(ns[l]);
    }
    STG_HD void draw12(P* nz) {
        constexpr int NL = Ln<P>::N;
        uint32_t w[6][NL];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
#pragma unroll
            for (int l = 0; l < NL; ++l) w[k][l] = st[l].next();
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) box_muller16<P>(w[k], nscale, nz[2 * k], nz[2 * k + 1]);
    }
    STG_HD void first(uint32_t, P* nz) { draw12(nz); }
    STG_HD void second(uint32_t, P* nz) { draw12(nz); }
};
// GEN 1: the whole stream from Philox4x32-10, counter-based down to the substep (the stream of the first half of round 2; 10.44
// against 8.64 ms per 1M-env step, profiles/README.md). One draw of three blocks serves a substep pair: blocks 3g and 3g+1 are
// evaluated for the first substep (6 of their 8 words), the two remaining words are carried to the second, which adds block 3g+2.
template <typename P>
struct ThermalSource<P, 1> {
    NoiseStream ns[Ln<P>::N];
    P nscale;
    uint32_t carry[2][Ln<P>::N];
    STG_HD void init(const NoiseStream* s, P scale) {
        nscale = scale;
#pragma unroll
        for (int l = 0; l < Ln<P>::N; ++l) ns[l] = s[l];
    }
    STG_HD void first(uint32_t g, P* nz) {
        constexpr int NL = Ln<P>::N;
        uint32_t w[8][NL];
        philox_block<NL>(ns, 3u * g, w);
        philox_block<NL>(ns, 3u * g + 1u, w + 4);
#pragma unroll
        for (int k = 0; k < 6; ++k) box_muller16<P>(w[k], nscale, nz[2 * k], nz[2 * k + 1]);
#pragma unroll
        for (int l = 0; l < NL; ++l) { carry[0][l] = w[6][l]; carry[1][l] = w[7][l]; }
    }
    STG_HD void second(uint32_t g, P* nz) {
        constexpr int NL = Ln<P>::N;
        uint32_t w[4][NL];
        philox_block<NL>(ns, 3u * g + 2u, w);
        box_muller16<P>(carry[0], nscale, nz[0], nz[1]);
        box_muller16<P>(carry[1], nscale, nz[2], nz[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) box_muller16<P>(w[k], nscale, nz[4 + 2 * k], nz[5 + 2 * k]);
    }
};

// constants of the fast path in pack form (hi/lo pairs, see StepConsts)
template <typename P>
struct PackConsts {
    P c_hi, c_lo, ac_hi, ac_lo, al;
};

// deterministic stage, e = z^ (same algebra as stage_z): 14 pack operations (+1 when SCALED)
template <typename P, bool SCALED>
STG_HD void stage_zp(const PackConsts<P>& c, P mx, P my, P mz, P aH, P aL, P nq, P& kx, P& ky, P& kz) {
    using K = Pk<P>;
    const P bz = K::fma(c.c_lo, mz, K::mul(c.c_hi, mz));
    const P w = K::add(K::fma(c.ac_hi, mz, aH), K::fma(c.ac_lo, mz, aL));
    const P u = K::mul(mz, w);
    const P s2 = K::fma(mx, mx, K::mul(my, my));
    kx = K::fma(mx, u, K::mul(my, bz));
    ky = K::fma(my, u, K::mul(K::neg(mx), bz));
    kz = SCALED ? K::mul(K::mul(s2, w), nq) : K::mul(s2, K::neg(w));       // nq = -inv_s^2
}
// thermal part: k += m x bn + alpha m x (m x bn)
template <typename P>
STG_HD void stage_noisep(const PackConsts<P>& c, P mx, P my, P mz, P bx, P by, P bz, P& kx, P& ky, P& kz) {
    using K = Pk<P>;
    const P px = K::fma(my, bz, K::mul(K::neg(mz), by));
    const P py = K::fma(mz, bx, K::mul(K::neg(mx), bz));
    const P pz = K::fma(mx, by, K::mul(K::neg(my), bx));
    const P qx = K::fma(my, pz, K::mul(K::neg(mz), py));
    const P qy = K::fma(mz, px, K::mul(K::neg(mx), pz));
    const P qz = K::fma(mx, py, K::mul(K::neg(my), px));
    kx = K::add(kx, K::fma(c.al, qx, px));
    ky = K::add(ky, K::fma(c.al, qy, py));
    kz = K::add(kz, K::fma(c.al, qz, pz));
}

// ---- fast RK4 substep: FP32 stages, e = z^ -------------------------------------------------------------------------------
// Working copy (fx, fy, fz) in FP32, master in FP64. All constants carry the factor 1/6 (k' = k/6), so
//     stage inputs are  m + 3 k1', m + 3 k2', m + 6 k3'  and the increment is  k1' + 2 (k2' + k3') + k4'.
// The per-substep renormalisation m/|m| of the reference is applied as a first-order-exact correction computed in FP32:
//     d = |m + inc|^2 - 1 = inc . (2 m + inc),   1/sqrt(1+d) - 1 = rho(d),   m_new = m + [inc + rho (m + inc)]
// and only the bracket (small) is added to the FP64 master, so the master keeps ~1e-15 resolution while no FP64 sqrt/div
// and only F2F conversions are needed per substep. The master is renormalised exactly in FP64 every STG_RESYNC_MASK+1
// substeps (integrate() in stt_env_core.cuh), which bounds the drift.
// Outputs: the correction `corr` to add to the master and d (callers fall back to the exact path when |d| >= 2^-6).
template <typename P, bool THERMAL, bool SCALED>
STG_HD void rk4_fast(const PackConsts<P>& c, P fx, P fy, P fz, P nq, P q, P aH1, P aL1, P aH2, P aL2, P aH4, P aL4,
                     const P* nz, P& ix, P& iy, P& iz, P& cx, P& cy, P& cz, P& d) {
    using K = Pk<P>;
    const P three = K::bc(3.0f), six = K::bc(6.0f), two = K::bc(2.0f);
    P k1x, k1y, k1z, k2x, k2y, k2z, k3x, k3y, k3z, k4x, k4y, k4z;
    stage_zp<P, SCALED>(c, fx, fy, fz, aH1, aL1, nq, k1x, k1y, k1z);
    if (THERMAL) stage_noisep<P>(c, fx, fy, fz, nz[0], nz[1], nz[2], k1x, k1y, k1z);
    {
        const P x = K::fma(three, k1x, fx), y = K::fma(three, k1y, fy), z = K::fma(three, k1z, fz);
        stage_zp<P, SCALED>(c, x, y, z, aH2, aL2, nq, k2x, k2y, k2z);
        if (THERMAL) stage_noisep<P>(c, x, y, z, nz[3], nz[4], nz[5], k2x, k2y, k2z);
    }
    {
        const P x = K::fma(three, k2x, fx), y = K::fma(three, k2y, fy), z = K::fma(three, k2z, fz);
        stage_zp<P, SCALED>(c, x, y, z, aH2, aL2, nq, k3x, k3y, k3z);
        if (THERMAL) stage_noisep<P>(c, x, y, z, nz[6], nz[7], nz[8], k3x, k3y, k3z);
    }
    {
        const P x = K::fma(six, k3x, fx), y = K::fma(six, k3y, fy), z = K::fma(six, k3z, fz);
        stage_zp<P, SCALED>(c, x, y, z, aH4, aL4, nq, k4x, k4y, k4z);
        if (THERMAL) stage_noisep<P>(c, x, y, z, nz[9], nz[10], nz[11], k4x, k4y, k4z);
    }
    ix = K::add(K::fma(two, K::add(k2x, k3x), k1x), k4x);
    iy = K::add(K::fma(two, K::add(k2y, k3y), k1y), k4y);
    iz = K::add(K::fma(two, K::add(k2z, k3z), k1z), k4z);
    const P ux = K::add(fx, ix), uy = K::add(fy, iy), uz = K::add(fz, iz);      // un-normalised new vector
    const P dxy = K::fma(ix, K::add(fx, ux), K::mul(iy, K::add(fy, uy)));
    const P dz = K::mul(iz, K::add(fz, uz));
    d = SCALED ? K::fma(q, dxy, dz) : K::add(dz, dxy);
    // rho = (1+d)^(-1/2) - 1 for |d| < 2^-6: truncation error < 0.28 d^4 (1.6e-8 relative to rho)
    const P rho = K::mul(d, K::fma(d, K::fma(d, K::bc(-0.3125f), K::bc(0.375f)), K::bc(-0.5f)));
    cx = K::fma(rho, ux, ix);
    cy = K::fma(rho, uy, iy);
    cz = K::fma(rho, uz, iz);
}

// ---- thermal fast path: one merged field per stage ------------------------------------------------------------------------
// With a thermal field every stage needs the full cross products anyway, so the anisotropy / demagnetisation field c m_z z^ is
// merged into the noise rotation vector, B = (bn_x, bn_y, c m_z + bn_z), and the two damping-like terms share one product:
//     p = m x B,   t = alpha p + a (m x z^) = alpha p + a (m_y, -m_x, 0),   k = p + m x t
// which is  m x B + alpha m x (m x B) + a m x (m x z^)  in 18 FMA-pipe instructions (32 for stage_zp + stage_noisep).
// COMP: compensated (hi, lo) constants c and a, as the deterministic path carries them - the injected-noise mode, whose parity
// with the FP64 reference is per trajectory (1e-4). With the in-kernel stream parity is statistical and the constants are plain
// FP32 (a relative rate error of 3e-8).
template <typename P>
struct ThermalConsts {
    P c_hi, c_lo, al;
};
template <typename P, bool COMP>
STG_HD void stage_th(const ThermalConsts<P>& c, P mx, P my, P mz, P aH, P aL, P bx, P by, P bz, P& kx, P& ky, P& kz) {
    using K = Pk<P>;
    const P Bz = COMP ? K::fma(c.c_hi, mz, K::fma(c.c_lo, mz, bz)) : K::fma(c.c_hi, mz, bz);
    const P px = K::fma(my, Bz, K::mul(K::neg(mz), by));
    const P py = K::fma(mz, bx, K::mul(K::neg(mx), Bz));
    const P pz = K::fma(mx, by, K::mul(K::neg(my), bx));
    P tx, ty;
    if (COMP) {
        tx = K::fma(c.al, px, K::fma(aH, my, K::mul(aL, my)));
        ty = K::fma(c.al, py, K::fma(K::neg(aH), mx, K::mul(K::neg(aL), mx)));
    } else {
        tx = K::fma(aH, my, K::mul(c.al, px));
        ty = K::fma(K::neg(aH), mx, K::mul(c.al, py));
    }
    const P tz = K::mul(c.al, pz);
    kx = K::fma(my, tz, K::fma(K::neg(mz), ty, px));
    ky = K::fma(mz, tx, K::fma(K::neg(mx), tz, py));
    kz = K::fma(mx, ty, K::fma(K::neg(my), tx, pz));
}
// RK4 substep of the thermal fast path; same conventions as rk4_fast (constants carry 1/6; outputs increment, correction, d).
// nz: the 12 noise rotation components of the substep (stage, xyz).
template <typename P, bool COMP>
STG_HD void rk4_thermal(const ThermalConsts<P>& c, P fx, P fy, P fz, P aH1, P aL1, P aH2, P aL2, P aH4, P aL4, const P* nz,
                        P& ix, P& iy, P& iz, P& cx, P& cy, P& cz, P& d) {
    using K = Pk<P>;
    const P three = K::bc(3.0f), six = K::bc(6.0f), two = K::bc(2.0f);
    P k1x, k1y, k1z, k2x, k2y, k2z, k3x, k3y, k3z, k4x, k4y, k4z;
    stage_th<P, COMP>(c, fx, fy, fz, aH1, aL1, nz[0], nz[1], nz[2], k1x, k1y, k1z);
    stage_th<P, COMP>(c, K::fma(three, k1x, fx), K::fma(three, k1y, fy), K::fma(three, k1z, fz), aH2, aL2, nz[3], nz[4], nz[5],
                      k2x, k2y, k2z);
    stage_th<P, COMP>(c, K::fma(three, k2x, fx), K::fma(three, k2y, fy), K::fma(three, k2z, fz), aH2, aL2, nz[6], nz[7], nz[8],
                      k3x, k3y, k3z);
    stage_th<P, COMP>(c, K::fma(six, k3x, fx), K::fma(six, k3y, fy), K::fma(six, k3z, fz), aH4, aL4, nz[9], nz[10], nz[11],
                      k4x, k4y, k4z);
    ix = K::add(K::fma(two, K::add(k2x, k3x), k1x), k4x);
    iy = K::add(K::fma(two, K::add(k2y, k3y), k1y), k4y);
    iz = K::add(K::fma(two, K::add(k2z, k3z), k1z), k4z);
    const P ux = K::add(fx, ix), uy = K::add(fy, iy), uz = K::add(fz, iz);
    d = K::fma(iz, K::add(fz, uz), K::fma(ix, K::add(fx, ux), K::mul(iy, K::add(fy, uy))));
    const P rho = K::mul(d, K::fma(d, K::fma(d, K::bc(-0.3125f), K::bc(0.375f)), K::bc(-0.5f)));
    cx = K::fma(rho, ux, ix);
    cy = K::fma(rho, uy, iy);
    cz = K::fma(rho, uz, iz);
}

// ---- FP32 master with a running compensation (in-kernel noise stream) -----------------------------------------------------
// The thermal kicks are ~1e-9 per stage (the reference does not scale h_th with dt), far below one ulp of an O(1) FP32
// component, so the state between substeps needs more than 24 bits. Instead of an FP64 master (3 F2F + 3 DADD + 3 F2F per
// substep on the quarter-rate conversion pipe) the working copy f carries a Kahan compensation e with m = f + e to ~2^-48:
//     y = corr + e;  t = f + y;  e = y - (t - f);  f = t          (4 FADD per component, packable)
// The captured error is exact whenever |f| >= |y|; where the increment exceeds the component the loss is < 2^-24 |increment|,
// i.e. the rounding the FP32 stages put into the increment anyway. Every STG_RESYNC_MASK+1 substeps f + e is renormalised
// exactly in FP64 and split again.
template <typename P>
STG_HD void kahan_add(P& f, P& e, P corr) {
    using K = Pk<P>;
    const P y = K::add(corr, e);
    const P t = K::add(f, y);
    e = K::add(y, K::neg(K::add(t, K::neg(f))));
    f = t;
}
// exact renormalisation of one env's (f + e) in FP64; guard as in guard_normalise (non-finite / zero norm -> (0,0,1) + flag)
STG_HD void kahan_renorm(float& fx, float& fy, float& fz, float& ex, float& ey, float& ez, int& guard) {
    ScaledState st{(double)fx + (double)ex, (double)fy + (double)ey, (double)fz + (double)ez, 1.0, 1.0, 1.0f};
    guard_normalise<float>(st, guard);
    fx = (float)st.sx; fy = (float)st.sy; fz = (float)st.z;
    ex = (float)(st.sx - (double)fx); ey = (float)(st.sy - (double)fy); ez = (float)(st.z - (double)fz);
}

struct FastState {
    ScaledState st;
    float fx, fy, fz, q;
};
STG_HD void fast_resync(FastState& s) {
    s.fx = (float)s.st.sx; s.fy = (float)s.st.sy; s.fz = (float)s.st.z; s.q = s.st.inv_s2f;
}
// apply the result of rk4_fast to one env's FP64 master (exact path when the norm change is large or non-finite)
STG_HD void fast_apply(FastState& s, float ix, float iy, float iz, float cx, float cy, float cz, float d, int& guard) {
    if (fabsf(d) < 0.015625f) {
        s.st.sx += (double)cx;
        s.st.sy += (double)cy;
        s.st.z += (double)cz;
        s.fx = (float)s.st.sx; s.fy = (float)s.st.sy; s.fz = (float)s.st.z;
    } else {
        // large or non-finite norm change (diverging parameters): exact FP64 path with the reference's guard
        s.st.sx += (double)ix;
        s.st.sy += (double)iy;
        s.st.z += (double)iz;
        guard_normalise<float>(s.st, guard);
        fast_resync(s);
    }
}
// ---- conditioning of an FP32-stage trajectory (deterministic / injected-noise runs of the fast path) ---------------------
// The FP32 stages leave a relative rounding error of ~2^-24 in the polar rate of every substep. For e = z^ the polar angle obeys
// an autonomous equation,  d(theta)/dn = sin(theta) W(z),  z = cos(theta),  W = 6 (ac z + a)  (ac = alpha c; the constants carry
// the RK4 factor 1/6), so an error in theta is multiplied per substep by exp(lambda),
//     lambda = d/d(theta) [sin(theta) W] = z W - 6 ac (1 - z^2),
// and every unit of z error turns the azimuth by |6 c| per substep. A trajectory held where W ~ 0 on the unstable side
// (lambda > 0: the current balances the anisotropy) for thousands of substeps amplifies the 1e-10 per substep beyond the 1e-4
// contract of the FP32 mode (the reference, physics/simple_solver.py:278-295, is FP64). The tracker integrates that bound
// forward, once per resynchronisation block of the master (nb = 16 substeps), with s = sin(theta):
//     A   <- A e^{nb lambda} + nb eps s max(1, e^{nb lambda})     (polar-angle error bound after the block)
//     Phi <- Phi + nb |6 c| s A                                     (azimuth error bound)
//     estimate of |delta m| = A + s_end Phi
// eps = 2^-24 (|6 ac| + |6 a| + 0.01 |6 c|): rounding of the polar rates plus the share of the precession term's rounding that
// leaks into the polar angle (dominant at low damping). The caller repeats an env with FP64 stages when the estimate exceeds
// STG_COND_TOL. Calibrated on the host build of these bodies against the FP64 oracle over random (state, pulse <= 5 ns) samples
// and four parameter sets (tests/test_hostsim_parity.py::test_fp32_conditioning_flag): the worst error of an env that is NOT
// flagged stays below ~1 x STG_COND_TOL, i.e. a margin of ~20x to 1e-4; 0.3 - 1.5 % of such samples are flagged.
struct CondTrack {
    float A, Phi;
};
#ifndef STG_COND_TOL
#define STG_COND_TOL 5.0e-6f
#endif
STG_HD float fast_ex2(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return exp2f(x);
#endif
}
// c, ac, a: the per-substep constants of the stages (already divided by 6); z: current m_z; s: current transverse magnitude
// sin(theta); nb: substeps in the block
STG_HD void cond_block(CondTrack& t, float c, float ac, float a, float z, float s, float nb) {
    // every product / sum explicit: the flag must not depend on which kernel or pack half evaluates the env
    using K = Pk<float>;
    const float ac6 = K::mul(6.0f, ac), a6 = K::mul(6.0f, a);
    s = fminf(s, 1.0f);
    const float lam = fmaf(z, fmaf(ac6, z, a6), K::mul(K::mul(-ac6, s), s));
    const float g = fast_ex2(fminf(K::mul(K::mul(nb, lam), 1.4426950408889634f), 80.0f));
    const float eps = K::mul(5.9604645e-8f, fmaf(0.06f, fabsf(c), K::add(fabsf(ac6), fabsf(a6))));
    t.A = fmaf(t.A, g, K::mul(K::mul(K::mul(nb, eps), s), fmaxf(1.0f, g)));
    t.Phi = fmaf(K::mul(K::mul(nb, fabsf(K::mul(6.0f, c))), s), t.A, t.Phi);
}
STG_HD float transverse_of(float fx, float fy, float inv_s) {      // sin(theta) from the (block-scaled) FP32 working copy
    return Pk<float>::mul(fast_sqrt(fmaf(fx, fx, Pk<float>::mul(fy, fy))), inv_s);
}
STG_HD bool cond_exceeded(const CondTrack& t, float s_end) {
    return !(fmaf(fminf(s_end, 1.0f), t.Phi, t.A) <= STG_COND_TOL);        // NaN counts as exceeded
}

template <typename P>
STG_HD void pack_consts(const StepConsts<float>& a, const StepConsts<float>& b, PackConsts<P>& c);
template <>
STG_HD void pack_consts<float>(const StepConsts<float>& a, const StepConsts<float>&, PackConsts<float>& c) {
    c.c_hi = a.c_hi; c.c_lo = a.c_lo; c.ac_hi = a.ac_hi; c.ac_lo = a.ac_lo; c.al = a.al_hi;
}
template <>
STG_HD void pack_consts<F2>(const StepConsts<float>& a, const StepConsts<float>& b, PackConsts<F2>& c) {
    c.c_hi = mk2(a.c_hi, b.c_hi); c.c_lo = mk2(a.c_lo, b.c_lo);
    c.ac_hi = mk2(a.ac_hi, b.ac_hi); c.ac_lo = mk2(a.ac_lo, b.ac_lo);
    c.al = mk2(a.al_hi, b.al_hi);
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
// Multiplications and additions are explicit (no FMA contraction): the same env must get the same bits from every kernel
// that evaluates it (one- and two-envs-per-thread variants, the FP64 second pass), whatever the compiler fuses elsewhere.
STG_HD double dot3(double ax, double ay, double az, double bx, double by, double bz) {
    return dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz));
}
STG_HD double resistance(const double* f, double mx, double my, double mz) {
    int kind = (int)f[FI_KIND];
    double rp = f[FI_RP], rap = f[FI_RAP];
    if (kind == 0) {
        double inv = 1.0 / sqrt(dot3(mx, my, mz, mx, my, mz));   // validate_magnetization renormalises (base_device.py:94-116)
        double c = dmul(dot3(mx, my, mz, f[FI_REFX], f[FI_REFY], f[FI_REFZ]), inv);
        double r = dmul(rp, dadd(1.0, dmul(f[FI_TMR], dadd(1.0, -c)) / 2.0));
        double lo = rp * 0.5;
        return r > lo ? r : lo;
    }
    double c = dot3(mx, my, mz, f[FI_REFX], f[FI_REFY], f[FI_REFZ]);
    double r = dadd(dadd(rp, dmul(rap - rp, dadd(1.0, -c)) / 2.0), f[FI_RSERIES]);
    return r > 1.0 ? r : 1.0;
}

}  // namespace stg

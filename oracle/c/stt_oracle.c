/* stt_oracle.c — plain-C restatement of the reference's SpinTorque-v0 step (FP64, reference operation order).
 *
 * TEST INFRASTRUCTURE ONLY: the checker for large parity tests and the timed CPU baseline of bench.py
 * (cpu_baseline.kind = "port"). Never linked into or called from the product library.
 *
 * Parity status: PINNED — tests/test_oracle_golden.py checks this file against the golden vectors generated from the live
 * reference (tests/golden/stt_env.npz, stt_multi.npz) to <= 1e-12.
 *
 * Follows (paths relative to /root/reference/spin_torque_gym):
 *   utils/monitoring.py:288-315         SafetyWrapper.validate_action (float32 clips)
 *   envs/spin_torque_env.py:409-433      _parse_action (FP64 clips)
 *   physics/simple_solver.py:137-139     step policy;  :263-295 Euler / RK4;  :297-344 dm/dt;  :346-388 effective field
 *   physics/simple_solver.py:208-229     guard + normalise
 *   envs/spin_torque_env.py:461-480      final renormalise, Joule energy with the pre-step m
 *   devices/stt_mram.py:78-94            TMR resistance
 *   envs/spin_torque_env.py:184-207      default reward;  :500-520 observation
 * Unlike the CUDA kernels nothing is pre-folded: every constant is recomputed the way the reference writes it.
 * Build: gcc -O2 -ffp-contract=off -pthread -fPIC -shared (oracle/Makefile). Envs are split over `nthreads` pthreads
 * (libgomp is not in the image).
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>

#define GAMMA 2.21e5
#define KB 1.38e-23
static const double MU0 = 4 * 3.141592653589793 * 1e-7;

typedef struct {
    double damping, ms, ku, volume, polarization;
    double easy_axis[3], ref[3];
    double r_p, r_ap, area;
    double temperature;
    double max_current, max_duration, success_threshold, energy_weight;
    int32_t max_steps, thermal, euler, pad;
} OracleParams;

/* Stand-in for np.random.normal(0, 1, 3) (physics/simple_solver.py:381) when no noise tensor is injected: the same
 * Marsaglia polar method NumPy's legacy generator uses (second value cached), on a splitmix64 stream per env. Only the
 * distribution matters here (bench timing and statistics); bit-level noise parity uses the injected tensor. */
typedef struct { uint64_t s; int has; double cached; } Rng;
static uint64_t splitmix(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static double rng_gauss(Rng* r) {
    if (r->has) { r->has = 0; return r->cached; }
    double x1, x2, r2;
    do {
        x1 = 2.0 * ((splitmix(&r->s) >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
        x2 = 2.0 * ((splitmix(&r->s) >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    double f = sqrt(-2.0 * log(r2) / r2);
    r->cached = f * x1; r->has = 1;
    return f * x2;
}

static void cross(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double norm(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

static void guard_normalise(double* m, int* guard) {
    if (!isfinite(m[0]) || !isfinite(m[1]) || !isfinite(m[2])) { m[0] = 0; m[1] = 0; m[2] = 1; *guard = 1; return; }
    double mag = norm(m);
    if (mag < 1e-12) { m[0] = 0; m[1] = 0; m[2] = 1; *guard = 1; return; }
    double o[3] = {m[0] / mag, m[1] / mag, m[2] / mag};
    if (!isfinite(o[0]) || !isfinite(o[1]) || !isfinite(o[2])) { m[0] = 0; m[1] = 0; m[2] = 1; *guard = 1; return; }
    m[0] = o[0]; m[1] = o[1]; m[2] = o[2];
}

/* dm/dt (physics/simple_solver.py:297-388); e normalised; xi NULL or 3 N(0,1) samples */
static void rhs(const double* m, double cur, const OracleParams* p, const double* e, double hth, const double* xi,
                double* out) {
    double h_k = (2 * p->ku) / (MU0 * p->ms);
    double s = dot(m, e);
    double h[3];
    for (int k = 0; k < 3; ++k) h[k] = 0.0 + h_k * s * e[k];
    h[2] = h[2] + (-p->ms * m[2]);
    if (hth > 0 && xi)
        for (int k = 0; k < 3; ++k) h[k] = h[k] + hth * xi[k];
    double tau[3] = {0, 0, 0};
    if (fabs(cur) > 1e-12) {
        double me[3], mme[3];
        cross(m, e, me);
        cross(m, me, mme);
        double pre = p->polarization * cur / (p->ms * p->volume);
        for (int k = 0; k < 3; ++k) tau[k] = pre * mme[k];
    }
    double geff = GAMMA / (1 + p->damping * p->damping);
    double pr[3], dmp[3];
    cross(m, h, pr);
    cross(m, pr, dmp);
    for (int k = 0; k < 3; ++k) out[k] = -geff * (pr[k] + p->damping * dmp[k]) + tau[k];
}

static double stt_resistance(const double* m, const OracleParams* p, const double* refn) {
    double mag = norm(m);
    double mm[3] = {m[0] / mag, m[1] / mag, m[2] / mag};
    double tmr = (p->r_ap - p->r_p) / p->r_p;
    double r = p->r_p * (1 + tmr * (1 - dot(mm, refn)) / 2);
    double lo = p->r_p * 0.5;
    return r > lo ? r : lo;
}

static void make_obs(const double* m, const double* tgt, const OracleParams* p, const double* refn, int step,
                     double total_e, double J, double T, float* o) {
    o[0] = (float)m[0]; o[1] = (float)m[1]; o[2] = (float)m[2];
    o[3] = (float)tgt[0]; o[4] = (float)tgt[1]; o[5] = (float)tgt[2];
    o[6] = (float)(stt_resistance(m, p, refn) / p->r_p);
    o[7] = (float)(p->temperature / 300.0);
    o[8] = (float)((double)(p->max_steps - step) / (double)p->max_steps);
    o[9] = (float)(total_e / 1e-12);
    o[10] = (float)(J / p->max_current);
    o[11] = (float)(T / p->max_duration);
}

typedef struct {
    const OracleParams* p;
    int64_t lo, hi;
    double* m; const double* target; double* total_energy; int32_t* step_count; double* last_action;
    const float* action; const double* noise; int64_t noise_stride;
    float* obs; double* reward; uint8_t* terminated; uint8_t* truncated; double* step_energy; int32_t* n_sub_out;
    int64_t total_sub;
    uint64_t rng_seed;
} Job;

static void* run_range(void* arg) {
    Job* jb = (Job*)arg;
    const OracleParams* p = jb->p;
    double* m = jb->m; const double* target = jb->target; double* total_energy = jb->total_energy;
    int32_t* step_count = jb->step_count; double* last_action = jb->last_action; const float* action = jb->action;
    const double* noise = jb->noise; int64_t noise_stride = jb->noise_stride; float* obs = jb->obs;
    double* reward = jb->reward; uint8_t* terminated = jb->terminated; uint8_t* truncated = jb->truncated;
    double* step_energy = jb->step_energy; int32_t* n_sub_out = jb->n_sub_out;
    double en = norm(p->easy_axis), rn = norm(p->ref);
    double e[3] = {p->easy_axis[0] / en, p->easy_axis[1] / en, p->easy_axis[2] / en};
    double refn[3] = {p->ref[0] / rn, p->ref[1] / rn, p->ref[2] / rn};
    double hth = 0.0;
    if (p->thermal && p->temperature > 0)
        hth = sqrt(2 * p->damping * KB * p->temperature / (MU0 * p->ms * p->volume * GAMMA));
    const int S = p->euler ? 1 : 4;
    int64_t total_sub = 0;
    for (int64_t i = jb->lo; i < jb->hi; ++i) {
        float a0 = action[2 * i], a1 = action[2 * i + 1];
        a0 = a0 < -1e8f ? -1e8f : (a0 > 1e8f ? 1e8f : a0);
        a1 = a1 < 1e-12f ? 1e-12f : (a1 > 1e-6f ? 1e-6f : a1);
        if (isnan(a0) || isnan(a1) || isinf(a0) || isinf(a1)) { a0 = 0.0f; a1 = 1e-12f; }
        double J = (double)a0, T = (double)a1;
        J = J < -p->max_current ? -p->max_current : (J > p->max_current ? p->max_current : J);
        T = T < 1e-12 ? 1e-12 : (T > p->max_duration ? p->max_duration : T);
        double* mi = m + 3 * i;
        const double* tg = target + 3 * i;
        double m_old[3] = {mi[0], mi[1], mi[2]};
        double prev_align = dot(m_old, tg);
        double dt = T / 100 < 1e-12 ? T / 100 : 1e-12;
        int ns = (int)(T / dt);
        if (ns < 10) ns = 10;
        dt = T / ns;
        double mm[3] = {mi[0], mi[1], mi[2]};
        int guard = 0;
        guard_normalise(mm, &guard);
        const double* nz = noise ? noise + (int64_t)i * noise_stride * S * 3 : 0;
        /* independent stream per (env, step): the start state is a HASH of (seed, env, step). Seeding env i with
         * base + i*gamma would only shift one splitmix sequence by i positions, i.e. the envs would share their noise. */
        uint64_t h0 = jb->rng_seed ^ 0xA0761D6478BD642FULL;
        uint64_t h1 = splitmix(&h0) ^ ((uint64_t)i * 0xE7037ED1A0B428DBULL);
        uint64_t h2 = splitmix(&h1) ^ ((uint64_t)step_count[i] * 0x8EBC6AF09C88C6E3ULL);
        Rng rng = {splitmix(&h2), 0, 0.0};
        const int own_noise = (!nz && hth > 0);
        double drawn[12];
        for (int s = 0; s < ns; ++s) {
            double ti = (double)s * dt;
            double k1[3], k2[3], k3[3], k4[3], tmp[3], f[3], mn[3];
            const double* x = nz ? nz + (int64_t)s * S * 3 : 0;
            if (own_noise) {
                for (int q = 0; q < S * 3; ++q) drawn[q] = rng_gauss(&rng);
                x = drawn;
            }
            rhs(mm, ti <= T ? J : 0.0, p, e, hth, x, f);
            if (p->euler) {
                for (int k = 0; k < 3; ++k) mn[k] = mm[k] + dt * f[k];
            } else {
                for (int k = 0; k < 3; ++k) { k1[k] = dt * f[k]; tmp[k] = mm[k] + k1[k] / 2; }
                rhs(tmp, (ti + dt / 2) <= T ? J : 0.0, p, e, hth, x ? x + 3 : 0, f);
                for (int k = 0; k < 3; ++k) { k2[k] = dt * f[k]; tmp[k] = mm[k] + k2[k] / 2; }
                rhs(tmp, (ti + dt / 2) <= T ? J : 0.0, p, e, hth, x ? x + 6 : 0, f);
                for (int k = 0; k < 3; ++k) { k3[k] = dt * f[k]; tmp[k] = mm[k] + k3[k]; }
                rhs(tmp, (ti + dt) <= T ? J : 0.0, p, e, hth, x ? x + 9 : 0, f);
                for (int k = 0; k < 3; ++k) {
                    k4[k] = dt * f[k];
                    mn[k] = mm[k] + (k1[k] + 2 * k2[k] + 2 * k3[k] + k4[k]) / 6;
                }
            }
            guard_normalise(mn, &guard);
            mm[0] = mn[0]; mm[1] = mn[1]; mm[2] = mn[2];
        }
        total_sub += ns;
        double nn = norm(mm);
        double m_new[3] = {mm[0] / nn, mm[1] / nn, mm[2] / nn};
        double energy = 0.0;
        if (fabs(J) > 1e-12) {
            double r = stt_resistance(m_old, p, refn);
            double v = J * r * p->area;
            energy = v * v / r * T;
        }
        mi[0] = m_new[0]; mi[1] = m_new[1]; mi[2] = m_new[2];
        total_energy[i] += energy;
        step_count[i] += 1;
        last_action[2 * i] = J; last_action[2 * i + 1] = T;
        double align = dot(m_new, tg);
        int success = align >= p->success_threshold;
        double rew = 0.0;
        rew += 10.0 * (success ? 10.0 : 0.0);
        rew += (-p->energy_weight) * (-energy / 1e-12);
        rew += 1.0 * (align - prev_align);
        rew += -2.0 * 0.0;
        if (isnan(rew) || isinf(rew)) rew = -1.0;
        rew = rew < -1e6 ? -1e6 : (rew > 1e6 ? 1e6 : rew);
        make_obs(m_new, tg, p, refn, step_count[i], total_energy[i], J, T, obs + 12 * i);
        reward[i] = rew;
        terminated[i] = (uint8_t)success;
        truncated[i] = (uint8_t)(step_count[i] >= p->max_steps);
        if (step_energy) step_energy[i] = energy;
        if (n_sub_out) n_sub_out[i] = ns;
        (void)guard;
    }
    jb->total_sub = total_sub;
    return 0;
}

/* One step of n envs. State arrays are [n][3] row-major here (oracle layout, independent of the product's SoA planes).
 * noise: NULL or [n][noise_stride][S][3]. Envs are split into contiguous ranges over nthreads pthreads.
 * Returns total substeps. */
int64_t stt_oracle_step(const OracleParams* p, int64_t n, double* m, const double* target, double* total_energy,
                        int32_t* step_count, double* last_action, const float* action, const double* noise,
                        int64_t noise_stride, float* obs, double* reward, uint8_t* terminated, uint8_t* truncated,
                        double* step_energy, int32_t* n_sub_out, int32_t nthreads, uint64_t rng_seed) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > n) nthreads = n > 0 ? (int32_t)n : 1;
    Job jobs[256];
    pthread_t th[256];
    for (int t = 0; t < nthreads; ++t) {
        Job* jb = &jobs[t];
        jb->p = p; jb->lo = n * t / nthreads; jb->hi = n * (t + 1) / nthreads;
        jb->m = m; jb->target = target; jb->total_energy = total_energy; jb->step_count = step_count;
        jb->last_action = last_action; jb->action = action; jb->noise = noise; jb->noise_stride = noise_stride;
        jb->obs = obs; jb->reward = reward; jb->terminated = terminated; jb->truncated = truncated;
        jb->step_energy = step_energy; jb->n_sub_out = n_sub_out; jb->total_sub = 0; jb->rng_seed = rng_seed;
    }
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], 0, run_range, &jobs[t]);
    run_range(&jobs[0]);
    int64_t total = jobs[0].total_sub;
    for (int t = 1; t < nthreads; ++t) { pthread_join(th[t], 0); total += jobs[t].total_sub; }
    return total;
}

void stt_oracle_obs(const OracleParams* p, int64_t n, const double* m, const double* target, const double* total_energy,
                    const int32_t* step_count, const double* last_action, float* obs) {
    double rn = norm(p->ref);
    double refn[3] = {p->ref[0] / rn, p->ref[1] / rn, p->ref[2] / rn};
    for (int64_t i = 0; i < n; ++i)
        make_obs(m + 3 * i, target + 3 * i, p, refn, step_count[i], total_energy[i], last_action[2 * i],
                 last_action[2 * i + 1], obs + 12 * i);
}

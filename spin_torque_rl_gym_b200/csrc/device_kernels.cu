// device_kernels.cu — K4: batched device-class operations (effective field, SOT torque, resistance, VCMA anisotropy,
// thermal field) behind the C-ABI. Elementwise, one row (one device state) per thread, FP64 like the reference's NumPy code.
// Reference: devices/stt_mram.py:56-94, devices/sot_mram.py:61-132,163-228, devices/vcma_mram.py:86-166,236-257,
// physics/thermal_model.py:46-137 (paths relative to /root/reference/spin_torque_gym).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/stg.h"
#include "llgs_core.cuh"

namespace stg {

__device__ __forceinline__ double vcma_keff(const StgDeviceParams& p, double v) {
    // devices/vcma_mram.py:122-147
    v = fmin(fmax(v, -p.breakdown_voltage), p.breakdown_voltage);
    const double change = -p.vcma_coefficient * fabs(v) / (p.dielectric_thickness * p.dielectric_thickness);
    const double k = p.uniaxial_anisotropy + change;
    const double kmin = -0.5 * p.uniaxial_anisotropy;
    return k > kmin ? k : kmin;
}

__global__ void __launch_bounds__(256) device_field_kernel(const __grid_constant__ StgDeviceParams p, const double* m,
                                                           const double* happ, int happ_rows, const double* volt,
                                                           double* out, int64_t n, int* zero_rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double mx = m[3 * i], my = m[3 * i + 1], mz = m[3 * i + 2];
    const int64_t hr = happ ? (happ_rows == 1 ? 0 : i) : 0;
    const double hx = happ ? happ[3 * hr] : 0.0, hy = happ ? happ[3 * hr + 1] : 0.0, hz = happ ? happ[3 * hr + 2] : 0.0;
    const double ex = p.easy_axis[0], ey = p.easy_axis[1], ez = p.easy_axis[2];
    double ku = p.uniaxial_anisotropy;
    if (p.kind == STG_DEV_STT) {   // validate_magnetization normalises first (devices/stt_mram.py:62)
        const double nrm = sqrt(mx * mx + my * my + mz * mz);
        if (zero_rows && nrm < 1e-12) atomicAdd(zero_rows, 1);      // devices/base_device.py:112-114 raises for these
        mx /= nrm; my /= nrm; mz /= nrm;
    } else if (p.kind == STG_DEV_VCMA) {
        ku = vcma_keff(p, volt ? volt[i] : 0.0);
    }
    const double hk = (2.0 * ku / (p.mu0 * p.saturation_magnetization)) * (mx * ex + my * ey + mz * ez);
    double ox = hx + hk * ex, oy = hy + hk * ey, oz = hz + hk * ez;
    if (p.kind != STG_DEV_STT) {   // shape demag, devices/sot_mram.py:114-132 (N_z = 1 - N_x - N_y = 0)
        ox += -p.saturation_magnetization * p.demag_n[0] * mx;
        oy += -p.saturation_magnetization * p.demag_n[1] * my;
        oz += -p.saturation_magnetization * p.demag_n[2] * mz;
    }
    out[3 * i] = ox; out[3 * i + 1] = oy; out[3 * i + 2] = oz;
}

__global__ void __launch_bounds__(256) device_resistance_kernel(const __grid_constant__ StgDeviceParams p, const double* m,
                                                                double* out, int64_t n, int* zero_rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (zero_rows && p.kind == STG_DEV_STT) {
        const double x = m[3 * i], y = m[3 * i + 1], z = m[3 * i + 2];
        if (sqrt(x * x + y * y + z * z) < 1e-12) atomicAdd(zero_rows, 1);
    }
    double f[FI_COUNT];
    f[FI_KIND] = (double)p.kind;
    f[FI_RP] = p.resistance_parallel; f[FI_RAP] = p.resistance_antiparallel;
    f[FI_TMR] = (p.resistance_antiparallel - p.resistance_parallel) / p.resistance_parallel;
    f[FI_REFX] = p.reference_magnetization[0]; f[FI_REFY] = p.reference_magnetization[1];
    f[FI_REFZ] = p.reference_magnetization[2];
    f[FI_RSERIES] = p.series_resistance;
    out[i] = resistance(f, m[3 * i], m[3 * i + 1], m[3 * i + 2]);
}

// devices/sot_mram.py:163-194: sigma = z^ x J^,  tau_DL = f_dl J (sigma x m),  tau_FL = f_fl J sigma
__global__ void __launch_bounds__(256) sot_torque_kernel(const __grid_constant__ StgDeviceParams p, const double* J, int j_rows,
                                                         const double* m, double sx, double sy, double sz, double* tau_dl,
                                                         double* tau_fl, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double j = J[j_rows == 1 ? 0 : i];
    const double mx = m[3 * i], my = m[3 * i + 1], mz = m[3 * i + 2];
    const double a = p.tau_dl_factor * j, b = p.tau_fl_factor * j;
    tau_dl[3 * i] = a * (sy * mz - sz * my);
    tau_dl[3 * i + 1] = a * (sz * mx - sx * mz);
    tau_dl[3 * i + 2] = a * (sx * my - sy * mx);
    tau_fl[3 * i] = b * sx; tau_fl[3 * i + 1] = b * sy; tau_fl[3 * i + 2] = b * sz;
}

__global__ void __launch_bounds__(256) vcma_anisotropy_kernel(const __grid_constant__ StgDeviceParams p, const double* volt,
                                                              double* out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = vcma_keff(p, volt[i]);
}

// physics/thermal_model.py:75-137: white  s*xi ;  OU  x <- decay x + sqrt(1-decay^2) xi,  field = s*x
__global__ void __launch_bounds__(256) thermal_field_kernel(double strength, double decay, double* state, double* out,
                                                            uint64_t seed, uint64_t offset, uint64_t call, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Philox ph{(uint32_t)seed, (uint32_t)(seed >> 32)};
    const uint64_t gid = offset + (uint64_t)i;
    uint32_t o[4];
    ph((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)call, (uint32_t)(call >> 32) ^ 0x7e57f1e1u, o);
    float z0, z1, z2, z3;
    box_muller(o[0], o[1], z0, z1);
    box_muller(o[2], o[3], z2, z3);
    double x = z0, y = z1, z = z2;
    if (state) {
        const double w = sqrt(1.0 - decay * decay);
        x = decay * state[3 * i] + w * x;
        y = decay * state[3 * i + 1] + w * y;
        z = decay * state[3 * i + 2] + w * z;
        state[3 * i] = x; state[3 * i + 1] = y; state[3 * i + 2] = z;
    }
    out[3 * i] = strength * x; out[3 * i + 1] = strength * y; out[3 * i + 2] = strength * z;
}

// Neel-Brown analytics of ThermalFluctuations (physics/thermal_model.py:46-73, 139-258) on a (temperature x device) grid: one
// thread per grid point, four output planes [n_t][n_dev] - thermal stability factor K_u V / k_B T, switching probability
// 1 - exp(-f0 exp(-E / k_B T) t_m) (capped at 1), retention time -ln(failure_rate) / (f0 exp(-E / k_B T)) in seconds, noise
// strength sqrt(2 alpha k_B T / (gamma mu0 Ms V)) - with the reference's T <= 0 conventions (inf, 0, inf, 0).
__global__ void __launch_bounds__(256) thermal_analytics_kernel(const __grid_constant__ StgThermalAnalyticsArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)a.n_t * a.n_dev;
    if (i >= n) return;
    const int64_t it = i / a.n_dev, id = i - it * a.n_dev;
    const double T = a.d_temperature[it];
    const double ku = a.d_ku[id], vol = a.d_volume[id];
    const double barrier = a.d_barrier ? a.d_barrier[id] : ku * vol;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double delta = inf, prob = 0.0, ret = inf, noise = 0.0;
    if (T > 0.0) {
        const double kt = a.k_b * T;
        delta = ku * vol / kt;
        const double rate = a.attempt_frequency * exp(-barrier / kt);
        prob = fmin(1.0 - exp(-rate * a.measurement_time), 1.0);
        if (a.failure_rate > 0.0) ret = -log(a.failure_rate) / (a.attempt_frequency * exp(-(barrier / kt)));
        noise = sqrt(2.0 * a.d_damping[id] * a.k_b * T / (a.gamma * a.mu0 * a.d_ms[id] * vol));
    }
    a.d_out[i] = delta;
    a.d_out[n + i] = prob;
    a.d_out[2 * n + i] = ret;
    a.d_out[3 * n + i] = noise;
}

// EnergyLandscape.compute_energy / compute_energy_gradient (physics/energy_landscape.py:36-104) for n states
__global__ void __launch_bounds__(256) energy_landscape_kernel(const __grid_constant__ StgEnergyParams p, const double* m,
                                                               const double* happ, int happ_rows, double* energy,
                                                               double* grad, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double mx = m[3 * i], my = m[3 * i + 1], mz = m[3 * i + 2];
    const double nrm = sqrt(mx * mx + my * my + mz * mz);
    mx /= nrm; my /= nrm; mz /= nrm;
    const int64_t hr = happ ? (happ_rows == 1 ? 0 : i) : 0;
    const double hx = happ ? happ[3 * hr] : 0.0, hy = happ ? happ[3 * hr + 1] : 0.0, hz = happ ? happ[3 * hr + 2] : 0.0;
    const double c = mx * p.easy_axis[0] + my * p.easy_axis[1] + mz * p.easy_axis[2];
    if (energy) {
        const double ez = -p.mu0 * p.saturation_magnetization * p.volume * (mx * hx + my * hy + mz * hz);
        const double ea = -p.uniaxial_anisotropy * p.volume * (c * c);
        const double ed = 0.5 * p.mu0 * (p.saturation_magnetization * p.saturation_magnetization) * p.volume *
                          (p.demag_factors[0] * (mx * mx) + p.demag_factors[1] * (my * my) + p.demag_factors[2] * (mz * mz));
        energy[i] = ez + ea + ed + 0.0;
    }
    if (grad) {
        const double hk = (2.0 * p.uniaxial_anisotropy / (p.mu0 * p.saturation_magnetization)) * c;
        grad[3 * i] = hx + hk * p.easy_axis[0] + (-p.saturation_magnetization * p.demag_factors[0] * mx);
        grad[3 * i + 1] = hy + hk * p.easy_axis[1] + (-p.saturation_magnetization * p.demag_factors[1] * my);
        grad[3 * i + 2] = hz + hk * p.easy_axis[2] + (-p.saturation_magnetization * p.demag_factors[2] * mz);
    }
}

// VectorizedMagneticsOperations (utils/vectorized_operations.py:288-393): row-wise 3-vector helpers. Products and sums are
// left unfused (dmul/dadd) and summed left to right so every result carries NumPy's roundings bit for bit.
__global__ void __launch_bounds__(256) vec3_op_kernel(int op, const double* a, const double* b, int b_rows, const double* p0,
                                                      const double* p1, double* out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ax = a[3 * i], ay = a[3 * i + 1], az = a[3 * i + 2];
    double bx = 0.0, by = 0.0, bz = 0.0;
    if (b) {
        const int64_t r = b_rows == 1 ? 0 : i;
        bx = b[3 * r]; by = b[3 * r + 1]; bz = b[3 * r + 2];
    }
    const double dot = dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz));     // np.sum(a*b, axis=1)
    switch (op) {
    case STG_VEC3_CROSS:                                                          // np.cross: a1*b2 - a2*b1, ...
        out[3 * i] = dadd(dmul(ay, bz), -dmul(az, by));
        out[3 * i + 1] = dadd(dmul(az, bx), -dmul(ax, bz));
        out[3 * i + 2] = dadd(dmul(ax, by), -dmul(ay, bx));
        break;
    case STG_VEC3_DOT:
        out[i] = dot;
        break;
    case STG_VEC3_NORMALIZE: {                                                    // v / max(||v||, 1e-12)
        double nrm = sqrt(dadd(dadd(dmul(ax, ax), dmul(ay, ay)), dmul(az, az)));
        nrm = nrm > 1e-12 ? nrm : 1e-12;
        out[3 * i] = ddiv(ax, nrm); out[3 * i + 1] = ddiv(ay, nrm); out[3 * i + 2] = ddiv(az, nrm);
        break;
    }
    case STG_VEC3_ANIS_ENERGY:                                                    // -K_u * V * (m.e)^2
        out[i] = dmul(dmul(-p0[i], p1[i]), dmul(dot, dot));
        break;
    case STG_VEC3_TMR_RESISTANCE: {                                               // max(R_P (1 + tmr (1 - cos)/2), R_P/2)
        const double rp = p0[i], rap = p1[i];
        const double tmr = ddiv(dadd(rap, -rp), rp);
        const double r = dmul(rp, dadd(1.0, ddiv(dmul(tmr, dadd(1.0, -dot)), 2.0)));
        const double lo = dmul(rp, 0.5);
        out[i] = r > lo ? r : lo;                                                 // np.maximum; NaN rows propagate below
        if (r != r) out[i] = r;
        break;
    }
    }
}

// EnergyLandscape.generate_phase_diagram (physics/energy_landscape.py:282-340): out[i][j] = |H_i| > h_k - |beta I_j|
__global__ void __launch_bounds__(256) phase_diagram_kernel(const double* currents, const double* fields, int n_currents,
                                                            int64_t n, double beta, double h_k, double* out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const int64_t i = g / n_currents, j = g - i * n_currents;
    const double h_critical = dadd(h_k, -fabs(dmul(beta, currents[j])));
    out[g] = fabs(fields[i]) > h_critical ? 1.0 : 0.0;
}

static inline unsigned grid_for(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace stg

using namespace stg;

extern "C" int stg_device_field_f64(const StgDeviceParams* p, const double* d_m, const double* d_happ, int32_t happ_rows,
                                    const double* d_voltage, double* d_out, int64_t n, int32_t* d_zero_rows, void* stream) {
    if (!p || !d_m || !d_out) return STG_E_NULL;
    if (n < 0 || (d_happ && happ_rows != 1 && happ_rows != n)) return STG_E_SIZE;
    if (p->kind < STG_DEV_STT || p->kind > STG_DEV_VCMA) return STG_E_ENUM;
    if (n == 0) return STG_OK;
    device_field_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(*p, d_m, d_happ, happ_rows, d_voltage, d_out, n,
                                                                       d_zero_rows);
    return (int)cudaGetLastError();
}

extern "C" int stg_device_resistance_f64(const StgDeviceParams* p, const double* d_m, double* d_out, int64_t n,
                                         int32_t* d_zero_rows, void* stream) {
    if (!p || !d_m || !d_out) return STG_E_NULL;
    if (n < 0) return STG_E_SIZE;
    if (p->kind < STG_DEV_STT || p->kind > STG_DEV_VCMA) return STG_E_ENUM;
    if (n == 0) return STG_OK;
    device_resistance_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(*p, d_m, d_out, n, d_zero_rows);
    return (int)cudaGetLastError();
}

extern "C" int stg_device_sot_torque_f64(const StgDeviceParams* p, const double* d_current, int32_t current_rows,
                                         const double* d_m, const double* current_direction, double* d_tau_dl,
                                         double* d_tau_fl, int64_t n, void* stream) {
    if (!p || !d_current || !d_m || !current_direction || !d_tau_dl || !d_tau_fl) return STG_E_NULL;
    if (n < 0 || (current_rows != 1 && current_rows != n)) return STG_E_SIZE;
    if (n == 0) return STG_OK;
    const double* d = current_direction;
    const double nn = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (!(nn > 0.0)) return STG_E_SIZE;
    const double jx = d[0] / nn, jy = d[1] / nn, jz = d[2] / nn;
    (void)jz;
    // z^ x J^ = (-J_y, J_x, 0)
    sot_torque_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(*p, d_current, current_rows, d_m, 0.0 * jz - 1.0 * jy,
                                                                     1.0 * jx - 0.0 * jz, 0.0, d_tau_dl, d_tau_fl, n);
    return (int)cudaGetLastError();
}

extern "C" int stg_vcma_anisotropy_f64(const StgDeviceParams* p, const double* d_voltage, double* d_out, int64_t n,
                                       void* stream) {
    if (!p || !d_voltage || !d_out) return STG_E_NULL;
    if (n < 0) return STG_E_SIZE;
    if (n == 0) return STG_OK;
    vcma_anisotropy_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(*p, d_voltage, d_out, n);
    return (int)cudaGetLastError();
}

extern "C" int stg_thermal_field_f64(double strength, double decay, double* d_state, double* d_out, uint64_t seed,
                                     uint64_t offset, uint64_t call_index, int64_t n, void* stream) {
    if (!d_out) return STG_E_NULL;
    if (n < 0 || !(decay >= 0.0 && decay <= 1.0)) return STG_E_SIZE;
    if (n == 0) return STG_OK;
    thermal_field_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(strength, decay, d_state, d_out, seed, offset,
                                                                        call_index, n);
    return (int)cudaGetLastError();
}

extern "C" int stg_thermal_analytics_f64(const StgThermalAnalyticsArgs* args, void* stream) {
    if (!args) return STG_E_NULL;
    const StgThermalAnalyticsArgs& a = *args;
    if (!a.d_temperature || !a.d_ku || !a.d_volume || !a.d_damping || !a.d_ms || !a.d_out) return STG_E_NULL;
    if (a.n_t < 0 || a.n_dev < 0) return STG_E_SIZE;
    const int64_t n = (int64_t)a.n_t * a.n_dev;
    if (n == 0) return STG_OK;
    thermal_analytics_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

extern "C" int stg_energy_landscape_f64(const StgEnergyParams* p, const double* d_m, const double* d_happ, int32_t happ_rows,
                                        double* d_energy, double* d_gradient, int64_t n, void* stream) {
    if (!p || !d_m || (!d_energy && !d_gradient)) return STG_E_NULL;
    if (n < 0 || (d_happ && happ_rows != 1 && happ_rows != n)) return STG_E_SIZE;
    if (n == 0) return STG_OK;
    energy_landscape_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(*p, d_m, d_happ, happ_rows, d_energy, d_gradient, n);
    return (int)cudaGetLastError();
}

extern "C" int stg_vec3_op_f64(int32_t op, const double* d_a, const double* d_b, int32_t b_rows, const double* d_p0,
                               const double* d_p1, double* d_out, int64_t n, void* stream) {
    if (op < STG_VEC3_CROSS || op > STG_VEC3_TMR_RESISTANCE) return STG_E_ENUM;
    if (!d_a || !d_out || (op != STG_VEC3_NORMALIZE && !d_b)) return STG_E_NULL;
    if ((op == STG_VEC3_ANIS_ENERGY || op == STG_VEC3_TMR_RESISTANCE) && (!d_p0 || !d_p1)) return STG_E_NULL;
    if (n < 0 || (d_b && b_rows != 1 && b_rows != n)) return STG_E_SIZE;
    if (n == 0) return STG_OK;
    vec3_op_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(op, d_a, d_b, b_rows, d_p0, d_p1, d_out, n);
    return (int)cudaGetLastError();
}

extern "C" int stg_phase_diagram_f64(const double* d_currents, const double* d_fields, int32_t n_currents, int32_t n_fields,
                                     double beta, double h_k, double* d_out, void* stream) {
    if (!d_currents || !d_fields || !d_out) return STG_E_NULL;
    if (n_currents < 0 || n_fields < 0) return STG_E_SIZE;
    const int64_t n = (int64_t)n_currents * n_fields;
    if (n == 0) return STG_OK;
    phase_diagram_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(d_currents, d_fields, n_currents, n, beta, h_k, d_out);
    return (int)cudaGetLastError();
}

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
                                                           double* out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double mx = m[3 * i], my = m[3 * i + 1], mz = m[3 * i + 2];
    const int64_t hr = happ ? (happ_rows == 1 ? 0 : i) : 0;
    const double hx = happ ? happ[3 * hr] : 0.0, hy = happ ? happ[3 * hr + 1] : 0.0, hz = happ ? happ[3 * hr + 2] : 0.0;
    const double ex = p.easy_axis[0], ey = p.easy_axis[1], ez = p.easy_axis[2];
    double ku = p.uniaxial_anisotropy;
    if (p.kind == STG_DEV_STT) {   // validate_magnetization normalises first (devices/stt_mram.py:62)
        const double nrm = sqrt(mx * mx + my * my + mz * mz);
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
                                                                double* out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
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

static inline unsigned grid_for(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace stg

using namespace stg;

extern "C" int stg_device_field_f64(const StgDeviceParams* p, const double* d_m, const double* d_happ, int32_t happ_rows,
                                    const double* d_voltage, double* d_out, int64_t n, void* stream) {
    if (!p || !d_m || !d_out) return STG_E_NULL;
    if (n < 0 || (d_happ && happ_rows != 1 && happ_rows != n)) return STG_E_SIZE;
    if (p->kind < STG_DEV_STT || p->kind > STG_DEV_VCMA) return STG_E_ENUM;
    if (n == 0) return STG_OK;
    device_field_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(*p, d_m, d_happ, happ_rows, d_voltage, d_out, n);
    return (int)cudaGetLastError();
}

extern "C" int stg_device_resistance_f64(const StgDeviceParams* p, const double* d_m, double* d_out, int64_t n, void* stream) {
    if (!p || !d_m || !d_out) return STG_E_NULL;
    if (n < 0) return STG_E_SIZE;
    if (p->kind < STG_DEV_STT || p->kind > STG_DEV_VCMA) return STG_E_ENUM;
    if (n == 0) return STG_OK;
    device_resistance_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(*p, d_m, d_out, n);
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

extern "C" int stg_energy_landscape_f64(const StgEnergyParams* p, const double* d_m, const double* d_happ, int32_t happ_rows,
                                        double* d_energy, double* d_gradient, int64_t n, void* stream) {
    if (!p || !d_m || (!d_energy && !d_gradient)) return STG_E_NULL;
    if (n < 0 || (d_happ && happ_rows != 1 && happ_rows != n)) return STG_E_SIZE;
    if (n == 0) return STG_OK;
    energy_landscape_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(*p, d_m, d_happ, happ_rows, d_energy, d_gradient, n);
    return (int)cudaGetLastError();
}

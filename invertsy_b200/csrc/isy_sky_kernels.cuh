// Stand-alone sky kernels: explicit ray orientations and spectrum_influence.
#pragma once
#include "isy_layout.cuh"

namespace isy {

// ------------------------------------------------------------------------------------------
// Sky.__call__ / UniformSky.__call__ on explicit ray orientations (sky.py:99-113, 269-334)
// ------------------------------------------------------------------------------------------
__global__ void sky_dirs_kernel(const double* __restrict__ quat, int64_t n, isy_sky_params sp,
                                const SunConst* __restrict__ sun, const uint8_t* __restrict__ eta,
                                double* __restrict__ Y, double* __restrict__ P,
                                double* __restrict__ A, double* __restrict__ theta_out, double* __restrict__ phi_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double dx, dy, dz;
    quat_dir(quat[4 * i + 0], quat[4 * i + 1], quat[4 * i + 2], quat[4 * i + 3], dx, dy, dz);
    dx = fmin(fmax(dx, -1.0), 1.0), dy = fmin(fmax(dy, -1.0), 1.0), dz = fmin(fmax(dz, -1.0), 1.0);  // :101
    double phi = wrap_pi_numpy(atan2(dy, dx));   // :102, :105
    double theta = acos(dz);                     // :103
    if (theta_out) theta_out[i] = theta;
    if (phi_out) phi_out[i] = phi;
    const bool masked = eta && eta[i];
    double y = masked ? 0.0 : sp.luminance, p = 0.0, a = CUDART_NAN;   // UniformSky: sky.py:77-84
    if (sp.kind == ISY_SKY_PEREZ) {
        double f_theta = perez_f(dz, theta < kHalfPi, sp.A, sp.B);
        sky_eval<true>(sp, sun[0], dx, dy, dz, theta, f_theta, y, p, a, masked);
    }
    Y[i] = y, P[i] = p, A[i] = a;
}

// ------------------------------------------------------------------------------------------
// spectrum_influence(v, irgbu).sum(axis=1) (sky.py:671-703), one CTA per call (view).
// The float32-evaluated wavelength powers of the reference are baked in as the doubles
// NumPy produces for np.power(wl/1000., 8) and np.power(1000./wl, 8) with wl float32 (:686-692).
// ------------------------------------------------------------------------------------------
__constant__ double kWl1[5] = {4.29981803894043, 0.06830433011054993, 0.006711667403578758, 0.002591485856100917,
                               0.00022518751211464405};
__constant__ double kWl2[5] = {0.23256799578666687, 14.640357971191406, 148.9943084716797, 385.87908935546875,
                               4440.74365234375};

__device__ __forceinline__ double block_reduce(double v, bool is_max, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        double u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmax(v, u) : v + u;   // fmax ignores NaN like np.nanmax
    }
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = is_max ? -CUDART_INF : 0.0;
    for (int w = 0; w < nw; ++w) r = is_max ? fmax(r, scratch[w]) : r + scratch[w];
    return r;
}

// T = double: the drop-in call; T = float: the luminance rows of a batched render, one CTA per view, in place
// (sum_all may alias v_all: every value is read into registers before the view's results are written ... per element,
// after the three reductions that need all of them).
template <typename T>
__global__ void spectrum_kernel(const T* v_all, int64_t n, const double* __restrict__ irgbu,
                                int64_t rows, T* sum_all, T* chan_all) {
    __shared__ double scratch[32];
    const T* v = v_all + (size_t)blockIdx.x * n;
    T* out = sum_all ? sum_all + (size_t)blockIdx.x * n : nullptr;
    T* chan = chan_all ? chan_all + (size_t)blockIdx.x * n * 5 : nullptr;
    double ssq = 0.0, vmax = CUDART_NAN;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double x = v[i];
        if (x == x) ssq += x * x;       // nansum(square(v))  :692
        vmax = fmax(vmax, x);           // nanmax(v)          :694
    }
    ssq = block_reduce(ssq, false, scratch);
    double vm = block_reduce(vmax == vmax ? vmax : -CUDART_INF, true, scratch);
    const double size = (double)n;      // float(v.size), v is (n, 1)
    double wmax = -CUDART_INF;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double x = v[i];
        const double* w = irgbu + 5 * (i % rows);
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            double l1 = 10.0 * w[c] * kWl1[c] * (x * x) / size;
            double l2 = 0.001 * w[c] * kWl2[c] * ssq / size;
            double t = x + l1 + l2;
            if (t == t) wmax = fmax(wmax, t);
        }
    }
    wmax = block_reduce(wmax, true, scratch);   // nanmax(v + l1 + l2)  :695
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double x = v[i];
        const double* w = irgbu + 5 * (i % rows);
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            double l1 = 10.0 * w[c] * kWl1[c] * (x * x) / size;
            double l2 = 0.001 * w[c] * kWl2[c] * ssq / size;
            double val = vm * (x + l1 + l2) / wmax;   // :696
            if (w[c] < 0) val = x;                    // :700
            if (chan) chan[5 * i + c] = (T)val;
            acc += val;                               // .sum(axis=1)  sky.py:344
        }
        if (out) out[i] = (T)acc;
    }
}

}  // namespace isy

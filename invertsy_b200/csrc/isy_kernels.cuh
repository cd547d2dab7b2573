// Kernels of the sm_100a view renderer (persistent, one CTA per position work item).
//
//   phase A1  cull      : min-vertex distance of every world triangle to the eye, fp64,
//                         reference order (world.py:94-98)                       -> slab_vis
//   phase A2  project   : vertex (theta, phi) by atan2, seam rule, reference crosses,
//                         normalised fp32 edge functions (world.py:109-120, 482-526) -> slab
//   phase B   rays      : every ray of every orientation at this position against the entry
//                         tile staged in shared memory; fp32 filter + exact fp64 re-check near
//                         edges; argmin of distance over containing copies == the painter's
//                         far->near overwrite (world.py:106-123); ground / background rule
//                         (world.py:79-90); analytic sky per ray (sky.py:304-334); cone
//                         average over the S samples of an ommatidium; ONE write of the view.
#pragma once
#include "isy_common.cuh"

namespace isy {

constexpr int kThreads = 512;        // CTA size
constexpr int kRPT = 4;              // rays per thread per pass (register-blocked against the smem tile)
constexpr int kPassRays = kThreads * kRPT;
constexpr int kSmemEntries = 1536;   // entries per shared-memory tile (72 KB)

struct RayOut {
    double r, g, b;
    double Y, P, A;
};

template <typename T>
__device__ __forceinline__ void store_val(void* base, size_t idx, double v) {
    reinterpret_cast<T*>(base)[idx] = (T)v;
}

// eye-frame (pitch, yaw) from the ommatidia quaternions, once per eye
__global__ void eye_setup_kernel(const double* __restrict__ quat, double* __restrict__ py, int R) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    double dx, dy, dz, p, y;
    quat_dir(quat[4 * r + 0], quat[4 * r + 1], quat[4 * r + 2], quat[4 * r + 3], dx, dy, dz);
    dir_pitch_yaw(dx, dy, dz, p, y);
    py[2 * r + 0] = p;
    py[2 * r + 1] = y;
}

// per-view sun constants (sky.py:309-312, 320-321, 501-521)
__global__ void sun_setup_kernel(const double* __restrict__ sun, int n, isy_sky_params sp, SunConst* __restrict__ out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    double ths = sun[2 * b + 0], phs = sun[2 * b + 1];
    double st, ct, sp_, cp;
    sincos(ths, &st, &ct);
    sincos(phs, &sp_, &cp);
    SunConst c;
    c.sx = st * cp;
    c.sy = st * sp_;
    c.sz = ct;
    // i_00 = L(0, theta_s), i_90 = L(pi/2, |theta_s - pi/2|)   (sky.py:309-310)
    double f00 = perez_f(ct, ths < kHalfPi, sp.A, sp.B);
    c.i00 = f00 * perez_phi(0.0, 1.0, sp.C, sp.D, sp.E);
    double z90 = fabs(ths - kHalfPi);
    double f90 = perez_f(cos(z90), z90 < kHalfPi, sp.A, sp.B);
    c.i90 = f90 * perez_phi(kHalfPi, cos(kHalfPi), sp.C, sp.D, sp.E);
    c.K = c.i00 * c.i90 / (c.i00 - c.i90 + kEps);
    double chi = (4.0 / 9.0 - sp.tau_L / 120.0) * (kPi - 2 * (kHalfPi - ths));
    c.Yz = (4.0453 * sp.tau_L - 4.9710) * tan(chi) - 0.2155 * sp.tau_L + 2.4192;
    c.Mp = exp(-(sp.tau_L - sp.c1) / (sp.c2 + kEps));
    out[b] = c;
}

// ------------------------------------------------------------------------------------------
// phase A: cull + project one position into this CTA's slab. Returns the entry count.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_vertices(const float* __restrict__ planes, int T, int t, double px, double py,
                                              double pz, double v[3][3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        v[k][0] = (double)__ldg(planes + (size_t)(3 * k + 0) * T + t) - px;   // world.py:94
        v[k][1] = (double)__ldg(planes + (size_t)(3 * k + 1) * T + t) - py;
        v[k][2] = (double)__ldg(planes + (size_t)(3 * k + 2) * T + t) - pz;
    }
}

__device__ int cull_and_project(const RenderParams& rp, int p, int* vis, Entry* entries, Exact* exacts, int* s_count,
                                int* s_nent) {
    const int tid = threadIdx.x, lane = tid & 31;
    const double px = rp.pos[3 * p + 0], py = rp.pos[3 * p + 1], pz = rp.pos[3 * p + 2];
    const double h = rp.horizon;
    const double h2_hi = h * h * (1.0 + 1e-12);  // above this the rounded norm cannot be <= horizon

    if (tid == 0) {
        *s_count = 0;
        *s_nent = 0;
    }
    __syncthreads();

    // ---- A1: cull (world.py:97-98) ----
    for (int base = 0; base < rp.T; base += kThreads) {
        int t = base + tid;
        bool visible = false;
        if (t < rp.T) {
            double v[3][3];
            load_vertices(rp.tri_planes, rp.T, t, px, py, pz, v);
            double s0 = norm2_rn(v[0][0], v[0][1], v[0][2]);
            double s1 = norm2_rn(v[1][0], v[1][1], v[1][2]);
            double s2 = norm2_rn(v[2][0], v[2][1], v[2][2]);
            bool nan_any = (s0 != s0) || (s1 != s1) || (s2 != s2);   // NaN vertex -> never visible (:98, :102)
            double smin = fmin(s0, fmin(s1, s2));
            if (!nan_any && smin <= h2_hi) visible = __dsqrt_rn(smin) <= h;
        }
        unsigned m = __ballot_sync(0xffffffffu, visible);
        if (m) {
            int basei = 0;
            if (lane == 0) basei = atomicAdd(s_count, __popc(m));
            basei = __shfl_sync(0xffffffffu, basei, 0);
            if (visible) {
                int slot = basei + __popc(m & ((1u << lane) - 1));
                if (slot < rp.slab_cap) vis[slot] = t;
            }
        }
    }
    __syncthreads();
    int V = *s_count;
    if (V > rp.slab_cap) {
        if (tid == 0) atomicExch(rp.status, (int)ISY_ERR_OVERFLOW);
        V = rp.slab_cap;
    }

    // ---- A2: angular projection of the visible triangles (world.py:109-120) ----
    for (int j = tid; j < V; j += kThreads) {
        int t = vis[j];
        double v[3][3];
        load_vertices(rp.tri_planes, rp.T, t, px, py, pz, v);
        double smin = fmin(norm2_rn(v[0][0], v[0][1], v[0][2]),
                           fmin(norm2_rn(v[1][0], v[1][1], v[1][2]), norm2_rn(v[2][0], v[2][1], v[2][2])));
        double dist = __dsqrt_rn(smin);
        double th[3], ph[3];
        double phmax = -CUDART_INF, phmin = CUDART_INF;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            ph[k] = atan2(v[k][1], v[k][0]);                                              // :109
            double rxy = __dsqrt_rn(__dadd_rn(__dmul_rn(v[k][0], v[k][0]), __dmul_rn(v[k][1], v[k][1])));
            th[k] = atan2(rxy, v[k][2]) - kHalfPi;                                        // :110
            phmax = fmax(phmax, ph[k]);
            phmin = fmin(phmin, ph[k]);
        }
        bool seam = (phmax - phmin) >= kPi;                                               // :116
        int slot = atomicAdd(s_nent, seam ? 2 : 1);
        if (slot + (seam ? 2 : 1) <= rp.slab_cap) {
            if (seam) {
                double p1[3], p2[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    p1[k] = ph[k] < 0 ? ph[k] + kTwoPi : ph[k];                           // :118
                    p2[k] = ph[k] > 0 ? ph[k] - kTwoPi : ph[k];                           // :119
                }
                finish_copy(th, p1, dist, t, entries[slot], exacts[slot]);
                finish_copy(th, p2, dist, t, entries[slot + 1], exacts[slot + 1]);
            } else {
                finish_copy(th, ph, dist, t, entries[slot], exacts[slot]);
            }
        }
    }
    __syncthreads();
    int E = *s_nent;
    if (E > rp.slab_cap) {
        if (tid == 0) atomicExch(rp.status, (int)ISY_ERR_OVERFLOW);
        E = rp.slab_cap;
    }
    return E;
}

// ------------------------------------------------------------------------------------------
// phase B helpers
// ------------------------------------------------------------------------------------------
struct Ray {
    double pitch, yaw;    // query point of the containment test (world.py:112)
    float fp, fy;         // the same, rounded once to fp32 for the filter
    double best;          // smallest distance among containing copies so far
    int hit;
};

__device__ __forceinline__ void test_tile(Ray (&ray)[kRPT], const Entry* __restrict__ tile, int n,
                                          const Exact* __restrict__ exacts) {
#pragma unroll 2
    for (int j = 0; j < n; ++j) {
        const float4 a = reinterpret_cast<const float4*>(tile + j)[0];
        const float4 b = reinterpret_cast<const float4*>(tile + j)[1];
        const float4 c = reinterpret_cast<const float4*>(tile + j)[2];
#pragma unroll
        for (int k = 0; k < kRPT; ++k) {
            float e0 = fmaf(a.x, ray[k].fp, fmaf(a.y, ray[k].fy, a.z));
            float e1 = fmaf(a.w, ray[k].fp, fmaf(b.x, ray[k].fy, b.y));
            float e2 = fmaf(b.z, ray[k].fp, fmaf(b.w, ray[k].fy, c.x));
            float m = fminf(e0, fminf(e1, e2));
            if (m >= -kEdgeEps) {   // rare: inside, or within the rounding band of an edge
                double dist = __hiloint2double(__float_as_int(c.w), __float_as_int(c.z));
                int tri = __float_as_int(c.y);
                bool closer = dist < ray[k].best || (dist == ray[k].best && tri < ray[k].hit);
                if (closer && (m > kEdgeEps || inside_exact(exacts[j], ray[k].pitch, ray[k].yaw))) {
                    ray[k].best = dist;
                    ray[k].hit = tri;
                }
            }
        }
    }
}

template <bool kF64>
__device__ __forceinline__ void write_view(const RenderParams& rp, size_t b, int r, bool valid, int hit, double best,
                                           RayOut o) {
    using T = typename std::conditional<kF64, double, float>::type;
    if (valid) {
        size_t ray_idx = b * (size_t)rp.R + r;
        if (rp.out.hit) rp.out.hit[ray_idx] = hit;
        if (rp.out.depth) store_val<T>(rp.out.depth, ray_idx, hit >= 0 ? best : CUDART_NAN);
    }
    const int S = rp.S;
    if (S > 1) {
        // weighted cone average over the S consecutive lanes of one ommatidium
        float w = valid ? rp.eye_w[r] : 0.f;
        double sA, cA;
        sincos(o.A, &sA, &cA);
        double wr = w * o.r, wg = w * o.g, wb = w * o.b;
        double wY = w * o.Y, wP = w * o.P, wS = w * sA, wC = w * cA;
        for (int off = S >> 1; off > 0; off >>= 1) {
            wr += __shfl_xor_sync(0xffffffffu, wr, off);
            wg += __shfl_xor_sync(0xffffffffu, wg, off);
            wb += __shfl_xor_sync(0xffffffffu, wb, off);
            wY += __shfl_xor_sync(0xffffffffu, wY, off);
            wP += __shfl_xor_sync(0xffffffffu, wP, off);
            wS += __shfl_xor_sync(0xffffffffu, wS, off);
            wC += __shfl_xor_sync(0xffffffffu, wC, off);
        }
        o.r = wr, o.g = wg, o.b = wb, o.Y = wY, o.P = wP;
        o.A = atan2(wS, wC);
        if (!valid || (r % S) != 0) return;
        r /= S;
    } else if (!valid) {
        return;
    }
    size_t oi = b * (size_t)rp.N + r;
    if (rp.out.rgb) {
        T* q = reinterpret_cast<T*>(rp.out.rgb) + 3 * oi;
        q[0] = (T)o.r, q[1] = (T)o.g, q[2] = (T)o.b;
    }
    if (rp.out.Y) store_val<T>(rp.out.Y, oi, o.Y);
    if (rp.out.dop) store_val<T>(rp.out.dop, oi, o.P);
    if (rp.out.aop) store_val<T>(rp.out.aop, oi, o.A);
}

// ------------------------------------------------------------------------------------------
// the render kernel (brute-force variant: every ray against every visible copy)
// ------------------------------------------------------------------------------------------
template <bool kF64>
__global__ void __launch_bounds__(kThreads, 1) render_brute_kernel(const RenderParams rp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Entry* tile = reinterpret_cast<Entry*>(smem_raw);
    __shared__ int s_count, s_nent;
    __shared__ unsigned int s_item;

    const int tid = threadIdx.x;
    int* vis = rp.slab_vis + (size_t)blockIdx.x * rp.slab_cap;
    Entry* entries = rp.slab_entry + (size_t)blockIdx.x * rp.slab_cap;
    Exact* exacts = rp.slab_exact + (size_t)blockIdx.x * rp.slab_cap;
    const int raysPerPos = rp.H * rp.R;
    const bool want_sky = rp.sky.kind != ISY_SKY_NONE;
    const bool perez = rp.sky.kind == ISY_SKY_PEREZ;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(rp.queue, 1u);
        __syncthreads();
        const unsigned item = s_item;
        if (item >= (unsigned)rp.items) break;
        const int p = item / rp.splits, split = item % rp.splits;

        const int E = cull_and_project(rp, p, vis, entries, exacts, &s_count, &s_nent);
        const int nchunks = (E + kSmemEntries - 1) / kSmemEntries;
        if (nchunks <= 1) {
            // common case: the whole projected world of this position lives in shared memory
            const float4* src = reinterpret_cast<const float4*>(entries);
            float4* dst = reinterpret_cast<float4*>(tile);
            for (int i = tid; i < E * 3; i += kThreads) dst[i] = src[i];
            __syncthreads();
        }

        for (int pass = split * kPassRays; pass < raysPerPos; pass += rp.splits * kPassRays) {
            Ray ray[kRPT];
            int rr[kRPT], hh[kRPT];
            double ddx[kRPT], ddy[kRPT], ddz[kRPT];
#pragma unroll
            for (int k = 0; k < kRPT; ++k) {
                int g = pass + k * kThreads + tid;
                bool valid = g < raysPerPos;
                int gg = valid ? g : 0;
                int h = gg / rp.R, r = gg - h * rp.R;
                rr[k] = valid ? r : -1;
                hh[k] = h;
                size_t b = (size_t)p * rp.H + h;
                double pitch, yaw, dx = 0, dy = 0, dz = 0;
                if (rp.vquat) {
                    const double* qv = rp.vquat + 4 * b;
                    const double* qo = rp.eye_quat + 4 * (size_t)r;
                    double x1 = qv[0], y1 = qv[1], z1 = qv[2], w1 = qv[3];
                    double x2 = qo[0], y2 = qo[1], z2 = qo[2], w2 = qo[3];
                    double w = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2;
                    double x = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2;
                    double y = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2;
                    double z = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2;
                    quat_dir(x, y, z, w, dx, dy, dz);
                    dir_pitch_yaw(dx, dy, dz, pitch, yaw);
                } else {
                    pitch = rp.eye_py[2 * (size_t)r + 0];
                    yaw = wrap_yaw(rp.eye_py[2 * (size_t)r + 1] + (rp.yaw ? rp.yaw[b] : 0.0));
                    if (perez) {
                        double sp_, cp, sy, cy;
                        sincos(pitch, &sp_, &cp);
                        sincos(yaw, &sy, &cy);
                        dx = cp * cy, dy = cp * sy, dz = -sp_;
                    }
                }
                ddx[k] = dx, ddy[k] = dy, ddz[k] = dz;
                ray[k].pitch = pitch;
                ray[k].yaw = yaw;
                ray[k].fp = (float)pitch;
                ray[k].fy = (float)yaw;
                ray[k].best = CUDART_INF;
                ray[k].hit = -1;
            }

            if (nchunks <= 1) {
                test_tile(ray, tile, E, exacts);
            } else {
                for (int c = 0; c < nchunks; ++c) {
                    int n = min(kSmemEntries, E - c * kSmemEntries);
                    __syncthreads();
                    const float4* src = reinterpret_cast<const float4*>(entries + (size_t)c * kSmemEntries);
                    float4* dst = reinterpret_cast<float4*>(tile);
                    for (int i = tid; i < n * 3; i += kThreads) dst[i] = src[i];
                    __syncthreads();
                    test_tile(ray, tile, n, exacts + (size_t)c * kSmemEntries);
                }
            }

#pragma unroll
            for (int k = 0; k < kRPT; ++k) {
                const bool valid = rr[k] >= 0;
                const int r = valid ? rr[k] : 0;
                const size_t b = (size_t)p * rp.H + hh[k];
                RayOut o;
                o.Y = 0, o.P = 0, o.A = CUDART_NAN;
                if (perez) {
                    const SunConst sc = rp.sun[rp.sun_per_view ? b : 0];
                    sky_eval(rp.sky, sc, ddx[k], ddy[k], ddz[k], ray[k].pitch + kHalfPi, o.Y, o.P, o.A);
                } else if (want_sky) {
                    o.Y = rp.sky.luminance;   // UniformSky: y = luminance, p = 0, a = nan (sky.py:77-79)
                }
                // world colour (world.py:79-90, :123)
                double r_ = CUDART_NAN, g_ = CUDART_NAN, b_ = CUDART_NAN;
                const bool from_sky = (rp.flags & ISY_FLAG_INIT_FROM_SKY) && want_sky;
                if (rp.init_c || from_sky) {
                    double lum = rp.init_c ? rp.init_c[b * (size_t)rp.R + r] : o.Y;
                    r_ = fmin(fmax(19.0 * lum + 13.0, 0.0), 255.0);      // luminance2skycolour, world.py:479
                    g_ = fmin(fmax(19.0 * lum + 135.0, 0.0), 255.0);
                    b_ = fmin(fmax(19.0 * lum + 201.0, 0.0), 255.0);
                }
                if (ray[k].pitch >= 0) r_ = 229.0 / 255.0, g_ = 183.0 / 255.0, b_ = 90.0 / 255.0;   // :90
                if (ray[k].hit >= 0) {
                    const float* col = rp.tri_rgb + 3 * (size_t)ray[k].hit;
                    r_ = __ldg(col + 0), g_ = __ldg(col + 1), b_ = __ldg(col + 2);
                }
                o.r = r_, o.g = g_, o.b = b_;
                write_view<kF64>(rp, b, r, valid, ray[k].hit, ray[k].best, o);
            }
        }
    }
}

}  // namespace isy

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
    if (sp.kind == ISY_SKY_PEREZ) sky_eval(sp, sun[0], dx, dy, dz, theta, y, p, a, masked);
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

__global__ void spectrum_kernel(const double* __restrict__ v_all, int64_t n, const double* __restrict__ irgbu,
                                int64_t rows, double* __restrict__ sum_all, double* __restrict__ chan_all) {
    __shared__ double scratch[32];
    const double* v = v_all + (size_t)blockIdx.x * n;
    double* out = sum_all ? sum_all + (size_t)blockIdx.x * n : nullptr;
    double* chan = chan_all ? chan_all + (size_t)blockIdx.x * n * 5 : nullptr;
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
            if (chan) chan[5 * i + c] = val;
            acc += val;                               // .sum(axis=1)  sky.py:344
        }
        if (out) out[i] = acc;
    }
}

}  // namespace isy

// Per-position angular bins (general orientations), the ray record, the edge test and the generic view writer.
#pragma once
#include "isy_project.cuh"

namespace isy {

// ------------------------------------------------------------------------------------------
// phase A3: angular bins
// ------------------------------------------------------------------------------------------
struct BinGrid {
    float th_lo, th_hi, inv_dth, dth;
    float inv_dph, dph;
    int rows, cols;
};

__device__ __forceinline__ bool bin_accepts(const Entry& e, float t0, float t1, float p0, float p1) {
    // largest value of each edge function on the rectangle [t0,t1] x [p0,p1]
    float m0 = fmaf(e.A0, e.A0 > 0.f ? t1 : t0, fmaf(e.B0, e.B0 > 0.f ? p1 : p0, e.C0));
    float m1 = fmaf(e.A1, e.A1 > 0.f ? t1 : t0, fmaf(e.B1, e.B1 > 0.f ? p1 : p0, e.C1));
    float m2 = fmaf(e.A2, e.A2 > 0.f ? t1 : t0, fmaf(e.B2, e.B2 > 0.f ? p1 : p0, e.C2));
    return fminf(m0, fminf(m1, m2)) >= kBinReject;
}

// one warp per copy, lanes over the bins of its bounding box; kFill = false counts, true fills
template <bool kFill>
__device__ __forceinline__ void bin_pass(SmemLayout& sm, const BinGrid& g, int E) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = warp; e < E; e += kWarps) {
        const float4 bb = sm.bbox[e];
        int r0 = max((int)floorf((bb.x - g.th_lo) * g.inv_dth), 0);
        int r1 = min((int)floorf((bb.y - g.th_lo) * g.inv_dth), g.rows - 1);
        int c0 = max((int)floorf((bb.z + (float)kPi) * g.inv_dph), 0);
        int c1 = min((int)floorf((bb.w + (float)kPi) * g.inv_dph), g.cols - 1);
        if (r0 > r1 || c0 > c1) continue;
        const Entry ent = sm.tile[e];
        const int w = c1 - c0 + 1, n = (r1 - r0 + 1) * w;
        const float inv_w = 1.0f / (float)w;
        for (int i = lane; i < n; i += 32) {
            const int q = (int)(((float)i + 0.5f) * inv_w);   // i / w, exact for i < 2^12
            int r = r0 + q, c = c0 + (i - q * w);
            float t0 = g.th_lo + r * g.dth - kBinMargin, t1 = g.th_lo + (r + 1) * g.dth + kBinMargin;
            float p0 = -(float)kPi + c * g.dph - kBinMargin, p1 = -(float)kPi + (c + 1) * g.dph + kBinMargin;
            if (!bin_accepts(ent, t0, t1, p0, p1)) continue;
            int b = r * g.cols + c;
            if (kFill) {
                unsigned pos = atomicAdd(&sm.bins.bin_cur[b], 1u);
                sm.bins.list[pos] = (unsigned short)e;
            } else {
                atomicAdd(&sm.bins.bin_off[b], 1u);
            }
        }
    }
}

// exclusive scan of bin_off[0..nb) in place (nb multiple of kThreads); returns the total
__device__ unsigned int block_scan_bins(SmemLayout& sm, int nb, unsigned int* s_warp) {
    constexpr int kScan = 512;         // threads that take part (kThreads >= 512)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = tid < kScan ? nb / kScan : 0;     // 8 (fine) or 2 (coarse)
    unsigned int local[kBins / kScan];
    unsigned int sum = 0;
    for (int i = 0; i < per; ++i) {
        local[i] = sm.bins.bin_off[tid * per + i];
        sum += local[i];
    }
    unsigned int inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned int wbase = 0, total = 0;
    for (int w = 0; w < kWarps; ++w) {
        unsigned int v = s_warp[w];
        if (w < warp) wbase += v;
        total += v;
    }
    unsigned int run = wbase + inc - sum;
    for (int i = 0; i < per; ++i) {
        sm.bins.bin_off[tid * per + i] = run;
        sm.bins.bin_cur[tid * per + i] = run;
        run += local[i];
    }
    if (tid == kScan - 1) sm.bins.bin_off[nb] = run;
    __syncthreads();
    return total;
}

// builds the bins for the E copies in sm.tile; false if they do not fit the list budget
__device__ bool build_bins(SmemLayout& sm, BinGrid& g, int E, const int* s_band, unsigned int* s_warp) {
    g.th_lo = key_float(s_band[0]);
    g.th_hi = key_float(s_band[1]);
    if (!(g.th_hi > g.th_lo)) {   // no copies: an empty single-row grid
        g.th_lo = 0.f, g.th_hi = 0.f;
    }
    g.th_lo = fmaxf(g.th_lo, -(float)kHalfPi - 1e-3f);
    g.th_hi = fminf(g.th_hi, (float)kHalfPi + 1e-3f);
    for (int level = 0; level < 2; ++level) {
        g.rows = kBinRows >> level;
        g.cols = kBinCols >> level;
        float span = fmaxf(g.th_hi - g.th_lo, 1e-3f);
        g.dth = span / g.rows;
        g.inv_dth = g.rows / span;
        g.dph = (float)kTwoPi / g.cols;
        g.inv_dph = g.cols / (float)kTwoPi;
        const int nb = g.rows * g.cols;
        for (int i = threadIdx.x; i < nb; i += kThreads) sm.bins.bin_off[i] = 0;
        __syncthreads();
        bin_pass<false>(sm, g, E);
        __syncthreads();
        unsigned int total = block_scan_bins(sm, nb, s_warp);
        if (total <= (unsigned)kListCap) {
            bin_pass<true>(sm, g, E);
            __syncthreads();
            return true;
        }
        __syncthreads();
    }
    return false;
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

__device__ __forceinline__ void test_entry(Ray& ray, const Entry* __restrict__ ent, const Exact* __restrict__ exact) {
    const float4 a = reinterpret_cast<const float4*>(ent)[0];
    const float4 b = reinterpret_cast<const float4*>(ent)[1];
    const float4 c = reinterpret_cast<const float4*>(ent)[2];
    float e0 = fmaf(a.x, ray.fp, fmaf(a.y, ray.fy, a.z));
    float e1 = fmaf(a.w, ray.fp, fmaf(b.x, ray.fy, b.y));
    float e2 = fmaf(b.z, ray.fp, fmaf(b.w, ray.fy, c.x));
    float m = fminf(e0, fminf(e1, e2));
    if (m >= -kEdgeEps) {   // inside, or within the rounding band of an edge
        double dist = __hiloint2double(__float_as_int(c.w), __float_as_int(c.z));
        int tri = __float_as_int(c.y);
        bool closer = dist < ray.best || (dist == ray.best && tri < ray.hit);
        if (closer && (m > kEdgeEps || inside_exact(*exact, ray.pitch, ray.yaw))) {
            ray.best = dist;
            ray.hit = tri;
        }
    }
}

template <bool kF64>
__device__ __forceinline__ void write_view(const RenderParams& rp, size_t b, int r, bool valid, int hit, double best,
                                           RayOut o) {
    using T = typename std::conditional<kF64, double, float>::type;
    if (valid) {
        size_t ray_idx = b * (size_t)rp.R + r;
        if (rp.out.hit) __stcs(rp.out.hit + ray_idx, hit);
        if (rp.out.depth) store_val<T>(rp.out.depth, ray_idx, hit >= 0 ? best : CUDART_NAN);
    }
    const int S = rp.S;
    if (S > 1) {
        // weighted cone average over the S consecutive lanes of one ommatidium
        float w = valid ? rp.eye_w[r] : 0.f;
        const double sA = o.sA, cA = o.cA;
        double wr = w * o.r, wg = w * o.g, wb = w * o.b;
        double wY = w * o.Y, wP = w * o.P, wS = w * sA, wC = w * cA;
        o.sr *= w, o.sg *= w, o.sb *= w;
        for (int off = S >> 1; off > 0; off >>= 1) {
            o.sr += __shfl_xor_sync(0xffffffffu, o.sr, off);
            o.sg += __shfl_xor_sync(0xffffffffu, o.sg, off);
            o.sb += __shfl_xor_sync(0xffffffffu, o.sb, off);
            wr += __shfl_xor_sync(0xffffffffu, wr, off);
            wg += __shfl_xor_sync(0xffffffffu, wg, off);
            wb += __shfl_xor_sync(0xffffffffu, wb, off);
            wY += __shfl_xor_sync(0xffffffffu, wY, off);
            wP += __shfl_xor_sync(0xffffffffu, wP, off);
            wS += __shfl_xor_sync(0xffffffffu, wS, off);
            wC += __shfl_xor_sync(0xffffffffu, wC, off);
        }
        o.r = wr, o.g = wg, o.b = wb, o.Y = wY, o.P = wP;
        if (!valid || (r % S) != 0) return;
        o.A = kF64 ? atan2(wS, wC) : (double)atan2f((float)wS, (float)wC);   // circular mean of the samples' angles
        r /= S;
    } else if (!valid) {
        return;
    }
    size_t oi = b * (size_t)rp.N + r;
    if (rp.out.rgb) {
        T* q = reinterpret_cast<T*>(rp.out.rgb) + 3 * oi;
        __stcs(q, (T)o.r), __stcs(q + 1, (T)o.g), __stcs(q + 2, (T)o.b);
    }
    if (rp.out.Y) store_val<T>(rp.out.Y, oi, o.Y);
    if (rp.out.dop) store_val<T>(rp.out.dop, oi, o.P);
    if (rp.out.aop) store_val<T>(rp.out.aop, oi, o.A);
    if (rp.out.resp) sensor_store<T>(rp.sensor, rp.out.resp, oi, o.sr, o.sg, o.sb);
}

}  // namespace isy

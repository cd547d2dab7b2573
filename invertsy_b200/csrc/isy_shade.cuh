// Final pass of the rasteriser path: shade, sky, write.
#pragma once
#include "isy_raster.cuh"

namespace isy {

// ------------------------------------------------------------------------------------------
// final pass of the rasteriser path, float outputs, one sample per ommatidium: per ray pick the
// winner from best[], shade, evaluate the sky and write the view (the only HBM traffic).
// Warp units = (block of 32 consecutive rays) x (chunk of headings); a ray's eye data are loaded
// once per unit and reused for all its headings; each store instruction covers 32 consecutive
// rays of one view.
// ------------------------------------------------------------------------------------------
#ifndef ISY_SHADE_U
#define ISY_SHADE_U 2
#endif
constexpr int kShadeU = ISY_SHADE_U;  // headings per lane and trip in the final pass (positions with several headings)

// kOneSun: every view shares one sun, its constants are read from shared memory (a compile-time address space).
// kAllOut: all six outputs were asked for (the usual call): no per-store null checks.
// kSU: headings per lane and trip (2 hides the latency of the fp64 chain; 1 for single-heading positions, where a
// second slot would only repeat the first)
// kBand: rays [band_lo, band_lo + band_n) only -- one band of a large eye (the winner buffer then has band_n slots
// per heading and sorted positions count from the band's first); otherwise the whole eye
// kCone: S > 1 samples per ommatidium (S consecutive rays = S consecutive lanes): hit and depth stay per ray, the
// other outputs are the weighted cone averages, written by the first lane of each ommatidium
template <bool kSmem, bool kOneSun, bool kAllOut, int kSU, bool kBand, bool kCone = false>
__device__ void final_pass_fast(const RenderParams& rp, SmemLayout& sm, const unsigned int* gbest, int p, int grp,
                                int Hcur, int band_lo, int band_n) {
    const int R = rp.R, r_lo = kBand ? band_lo : 0, r_n = kBand ? band_n : R, Rrow = r_n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool perez = rp.sky.kind == ISY_SKY_PEREZ, uniform = rp.sky.kind == ISY_SKY_UNIFORM;
    const bool from_sky = (rp.flags & ISY_FLAG_INIT_FROM_SKY) && (perez || uniform);
    float* const o_rgb = reinterpret_cast<float*>(rp.out.rgb);
    float* const o_depth = reinterpret_cast<float*>(rp.out.depth);
    float* const o_Y = reinterpret_cast<float*>(rp.out.Y);
    float* const o_P = reinterpret_cast<float*>(rp.out.dop);
    float* const o_A = reinterpret_cast<float*>(rp.out.aop);
    const float nanf_ = __int_as_float(0x7fc00000);
    const int nb = (r_n + 31) >> 5;
    int nch = min(Hcur, max(1, rp.shade_nch));
    int hper = (Hcur + nch - 1) / nch;
    hper = min(Hcur, (hper + kSU - 1) / kSU * kSU);      // whole trips
    nch = (Hcur + hper - 1) / hper;
    for (int u = warp; u < nb * nch; u += kWarps) {
        const int rb = u / nch, hc = u - rb * nch;
        const int r = r_lo + rb * 32 + lane;
        const bool valid = r < r_lo + r_n;
        const int rr = valid ? r : r_lo + r_n - 1;
        const double e_pitch = rp.eye_py[2 * (size_t)rr];
        const bool ground = e_pitch >= 0;                               // world.py:90
        double4 t4v = make_double4(0, 1, 0, 1);
        double f_theta = 0;
        if (perez) {
            t4v = *reinterpret_cast<const double4*>(rp.eye_trig + 4 * (size_t)rr);
            f_theta = rp.eye_ftheta[rr];
        }
        const int pos = __ldg(rp.grid_pos_of + rr) - r_lo;
        const int hend = min(Hcur, (hc + 1) * hper);
        const float wray = kCone ? (valid ? __ldg(rp.eye_w + rr) : 0.f) : 1.f;   // cone weight of this sample

        // kSU headings per trip: their sky evaluations are independent instruction streams (the pass is
        // bound by the latency of the fp64 chain, not by issue slots)
        // element offset of (view of heading hh0, this lane's ray) in the [view][ray] outputs; consecutive
        // headings of a group are ngroups views apart, so it advances by a constant
        const size_t ostride = (size_t)rp.ngroups * (size_t)R;
        size_t off0 = ((size_t)p * rp.H + grp + (size_t)(hc * hper) * rp.ngroups) * (size_t)R + (size_t)r;
        // ... and of (view, this lane's ommatidium) in the per-ommatidium outputs of a cone-sampled eye
        const size_t nstride = (size_t)rp.ngroups * (size_t)rp.N;
        size_t offn0 = ((size_t)p * rp.H + grp + (size_t)(hc * hper) * rp.ngroups) * (size_t)rp.N + (size_t)(r / max(rp.S, 1));
        for (int hh0 = hc * hper; hh0 < hend; hh0 += kSU, off0 += kSU * ostride, offn0 += kSU * nstride) {
            double Y[kSU], P[kSU], A[kSU];
            float sA[kSU], cA[kSU];      // (cone-sampled eyes) sin / cos of the polarisation angle
            float pol[kSU];              // polarisation term of the photoreceptor (out.resp)
            size_t view[kSU];
            unsigned int id[kSU];
#pragma unroll
            for (int q = 0; q < kSU; ++q) {
                Y[q] = 0, P[q] = 0, A[q] = CUDART_NAN, pol[q] = 1.f;
                const int hh = min(hh0 + q, hend - 1);
                view[q] = (size_t)p * rp.H + grp + (size_t)hh * rp.ngroups;   // group = every ngroups-th heading
                const int hs = sm.ras.hinv[hh];
                ISY_BOUND(pos >= 0 && pos < Rrow && hs < Hcur && (!kSmem || hs * Rrow + pos < kBestCap));
                id[q] = kSmem ? (unsigned)sm.ras.best[hs * Rrow + pos] : gbest[(size_t)hs * R + pos];
                ISY_BOUND(id[q] == kNoCopy || id[q] < (unsigned)kMaxE);
                if (perez) {
                    const double hsn = sm.head_sin[hh], hcs = sm.head_cos[hh];
                    const double sy = t4v.z * hcs + t4v.w * hsn, cy = t4v.w * hcs - t4v.z * hsn;   // yaw_e + heading
                    const SunConst& sc = kOneSun ? sm.sun0 : rp.sun[rp.sun_per_view ? view[q] : 0];
                    sky_eval<false>(rp.sky, sc, t4v.y * cy, t4v.y * sy, -t4v.x, e_pitch + kHalfPi, f_theta, Y[q], P[q], A[q],
                                    false, sm.gtab);
                    if (rp.out.resp && rp.sensor.pol_op > 0.f)
                        pol[q] = sensor_pol_term(rp.sensor.pol_op, (float)P[q], (float)(t4v.y * cy - sc.sx), (float)(t4v.y * sy - sc.sy),
                                                 (float)cy, (float)sy);
                    if (kCone) {
                        // A = atan2(dy - sy, dx - sx) + pi/2 (sky.py:323)  =>  (sin A, cos A) = (X, -Y) / |(X, Y)|
                        const float X = (float)(t4v.y * cy - sc.sx), Yd = (float)(t4v.y * sy - sc.sy), h2 = X * X + Yd * Yd;
                        float ih;
                        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ih) : "f"(fmaxf(h2, 1e-36f)));
                        sA[q] = h2 > 0.f ? X * ih : 1.f;               // atan2(0, 0) = 0 -> A = pi/2
                        cA[q] = -Yd * ih;
                    }
                } else if (uniform) {
                    Y[q] = rp.sky.luminance;
                }
            }
#pragma unroll
            for (int q = 0; q < kSU; ++q) {
                if (hh0 + q >= hend) break;
                float cr = nanf_, cg = nanf_, cb = nanf_, depth = nanf_;
                int hit = -1;
                if (from_sky) {                                              // luminance2skycolour, world.py:479
                    cr = (float)fmin(fmax(19.0 * Y[q] + 13.0, 0.0), 255.0);
                    cg = (float)fmin(fmax(19.0 * Y[q] + 135.0, 0.0), 255.0);
                    cb = (float)fmin(fmax(19.0 * Y[q] + 201.0, 0.0), 255.0);
                }
                if (ground) cr = (float)(229.0 / 255.0), cg = (float)(183.0 / 255.0), cb = (float)(90.0 / 255.0);
                if (id[q] != kNoCopy) {
                    const Entry& w = sm.tile[id[q]];
                    hit = w.tri;
                    depth = (float)__hiloint2double(w.dist_hi, w.dist_lo);
                    const float* col = rp.tri_rgb + 3 * (size_t)hit;
                    cr = __ldg(col + 0), cg = __ldg(col + 1), cb = __ldg(col + 2);
                }
                float sr = 0.f, sg = 0.f, sb = 0.f;      // the sample on the 0..1 scale of the photoreceptor transform
                if (rp.out.resp) sensor_colour(rp.sensor, id[q] != kNoCopy, cr, cg, cb, ground, perez || uniform, Y[q], pol[q], sr, sg, sb);
                if (kCone) {
                    if (valid) {
                        const size_t i = off0 + q * ostride;
                        if (rp.out.hit) __stcs(rp.out.hit + i, hit);
                        if (o_depth) __stcs(o_depth + i, depth);
                    }
                    // weighted cone average over the S consecutive lanes of an ommatidium (S divides 32).
                    // Luminance (and the 0..255 colours painted from it) in fp64, the rest in fp32.
                    double wY = (double)wray * Y[q];
                    double wr = 0, wg = 0, wb = 0;
                    float fr = wray * cr, fg = wray * cg, fb = wray * cb;
                    float wP = wray * (float)P[q], wS = wray * sA[q], wC = wray * cA[q];
                    if (from_sky) wr = (double)wray * cr, wg = (double)wray * cg, wb = (double)wray * cb;
                    sr *= wray, sg *= wray, sb *= wray;
                    if (rp.out.resp)
                        for (int off = rp.S >> 1; off > 0; off >>= 1) {
                            sr += __shfl_xor_sync(0xffffffffu, sr, off);
                            sg += __shfl_xor_sync(0xffffffffu, sg, off);
                            sb += __shfl_xor_sync(0xffffffffu, sb, off);
                        }
                    for (int off = rp.S >> 1; off > 0; off >>= 1) {
                        wY += __shfl_xor_sync(0xffffffffu, wY, off);
                        wP += __shfl_xor_sync(0xffffffffu, wP, off);
                        wS += __shfl_xor_sync(0xffffffffu, wS, off);
                        wC += __shfl_xor_sync(0xffffffffu, wC, off);
                        if (from_sky) {
                            wr += __shfl_xor_sync(0xffffffffu, wr, off);
                            wg += __shfl_xor_sync(0xffffffffu, wg, off);
                            wb += __shfl_xor_sync(0xffffffffu, wb, off);
                        } else {
                            fr += __shfl_xor_sync(0xffffffffu, fr, off);
                            fg += __shfl_xor_sync(0xffffffffu, fg, off);
                            fb += __shfl_xor_sync(0xffffffffu, fb, off);
                        }
                    }
                    if (from_sky) fr = (float)wr, fg = (float)wg, fb = (float)wb;
                    if (valid && (r & (rp.S - 1)) == 0) {
                        const size_t i = offn0 + q * nstride;
                        if (o_rgb) {
                            float* qq = o_rgb + 3 * i;
                            __stcs(qq, fr), __stcs(qq + 1, fg), __stcs(qq + 2, fb);
                        }
                        if (o_Y) __stcs(o_Y + i, (float)wY);
                        if (o_P) __stcs(o_P + i, wP);
                        if (o_A) __stcs(o_A + i, perez ? atan2_fast(wS, wC) : nanf_);   // circular mean of the angles
                        if (rp.out.resp) sensor_store<float>(rp.sensor, rp.out.resp, i, sr, sg, sb);
                    }
                } else {
                    // streaming stores: the views are written once and never read back by the kernel, so they
                    // should leave L2 before the slabs, the world and the eye tables do
                    const size_t i = off0 + q * ostride;
                    if (!valid) continue;
                    if (kAllOut || o_rgb) {
                        float* qq = o_rgb + 3 * i;
                        __stcs(qq, cr), __stcs(qq + 1, cg), __stcs(qq + 2, cb);
                    }
                    if (kAllOut || rp.out.hit) __stcs(rp.out.hit + i, hit);
                    if (kAllOut || o_depth) __stcs(o_depth + i, depth);
                    if (kAllOut || o_Y) __stcs(o_Y + i, (float)Y[q]);
                    if (kAllOut || o_P) __stcs(o_P + i, (float)P[q]);
                    if (kAllOut || o_A) __stcs(o_A + i, (float)A[q]);
                    if (rp.out.resp) sensor_store<float>(rp.sensor, rp.out.resp, i, sr, sg, sb);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Row templates + patches (small eyes, yaw-only views, float outputs, one sample per ommatidium).
//
// Most of a view does not depend on where the eye stands.  A ray that hits nothing gets hit = -1, depth = NaN and
// the ground colour or NaN by its own pitch (world.py:80, :90) -- one template row per output for the whole launch
// (row_templates_kernel).  Its sky is constant under a uniform sky (sky.py:77-79), and a function of (heading, ray)
// alone when every position looks along the same row of headings under one sun (grids, heat maps):
// sky_shared_kernel evaluates that once per launch into a [3][H][R] table with the per-view pass's own sky_eval.
// So a position's views are written in two steps:
//   copy_rows        -- while the copies are rasterised, a few warps stream the template / table rows into the
//                       position's views, 16 bytes per lane.  The SM's store path (32 B/clk for full lines, 11.6 B/clk
//                       for a 12-byte-stride colour store: tools/microbench/store_bw.cu) is idle during the
//                       rasterisation, and it is what bounds a final pass that writes whole views.
//   final_pass_patch -- after the rasterisation, the winner buffer is scanned for the rays that did hit something
//                       (7.6 % on the benchmark grid) and only those get their hit, depth and colour stored, all
//                       of it looked up in shared memory.
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T zero_of();
template <>
__device__ __forceinline__ float zero_of<float>() { return 0.f; }
template <>
__device__ __forceinline__ float4 zero_of<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// rows of the views of (position p, heading group grp), by the warps [w0, w0 + nw); T = float4 when the rows are
// 16-byte aligned (R a multiple of 4), else float.
//   A. sky rows of a table (one row per heading): load + streaming store, eight loads in flight per lane.
//   B. rows that are the same for every heading (templates of hit / depth / rgb, the sky rows of a uniform sky):
//      a warp takes a column of 32 elements, loads it once and stores it into the Hcur views -- a pure store stream.
// Sky rows are never touched again: streaming stores.  The hit / depth / rgb rows will be patched by final_pass_patch a
// few tens of microseconds later: normal stores, and written as late in the rasterisation as there is time for, so
// that the patches meet them in L2 (kTemplates picks the part: the render kernel calls it twice).
#ifndef ISY_RESP_EARLY
#define ISY_RESP_EARLY 1
#endif
constexpr bool kRespEarly = ISY_RESP_EARLY;   // the per-heading response rows go out with the sky rows (1) or with the templates (0)

template <typename T, bool kSky, bool kTemplates>
__device__ void copy_rows(const RenderParams& rp, int p, int grp, int Hcur, int w0, int nw) {
    constexpr int kPer = sizeof(T) / sizeof(float);
    const int nR = rp.R / kPer;                              // elements per row of R floats
    const int lane = threadIdx.x & 31, warp = (int)(threadIdx.x >> 5) - w0;
    if (warp < 0 || warp >= nw) return;
    const size_t view0 = (size_t)p * rp.H + grp;             // views of the group: view0 + hh * ngroups
    const T* const tab = reinterpret_cast<const T*>(rp.sky_tab);
    T* const o_sky[3] = {reinterpret_cast<T*>(rp.out.Y), reinterpret_cast<T*>(rp.out.dop), reinterpret_cast<T*>(rp.out.aop)};
    if (kSky && rp.sky_rows_h > 1) {                         // A
        constexpr int kU = 8;
        const size_t plane = (size_t)rp.sky_rows_h * (size_t)nR;
        for (int it = warp; it < 3 * Hcur; it += nw) {
            const int k = it / Hcur, hh = it - k * Hcur;
            T* const dk = k == 0 ? o_sky[0] : (k == 1 ? o_sky[1] : o_sky[2]);
            if (!dk) continue;
            const T* const src = tab + (size_t)k * plane + (size_t)(grp + hh * rp.ngroups) * nR;
            T* const dst = dk + (view0 + (size_t)hh * rp.ngroups) * (size_t)nR;
            for (int c0 = lane; c0 < nR; c0 += 32 * kU) {
                T v[kU];
#pragma unroll
                for (int q = 0; q < kU; ++q) {
                    v[q] = zero_of<T>();
                    if (c0 + 32 * q < nR) v[q] = __ldg(src + c0 + 32 * q);
                }
#pragma unroll
                for (int q = 0; q < kU; ++q)
                    if (c0 + 32 * q < nR) __stcs(dst + c0 + 32 * q, v[q]);
            }
        }
    }
    // A': the per-heading rows of the photoreceptor responses (plane 3 of the table).  They are patched like the
    // template rows, so they go out with them, late, as normal stores.
    if (kRespEarly ? kSky : kTemplates)
    if (rp.sky_rows_h > 1 && rp.out.resp) {
        const size_t plane = (size_t)rp.sky_rows_h * (size_t)nR;
        T* const dk = reinterpret_cast<T*>(rp.out.resp);
        for (int hh = warp; hh < Hcur; hh += nw) {
            const T* const src = tab + 3 * plane + (size_t)(grp + hh * rp.ngroups) * nR;
            T* const dst = dk + (view0 + (size_t)hh * rp.ngroups) * (size_t)nR;
            for (int c0 = lane; c0 < nR; c0 += 32) dst[c0] = __ldg(src + c0);
        }
    }
    // B: columns of [hit | depth | rgb (3 rows long) | responses (no / uniform sky) | Y | DoP | AoP of a uniform sky]
    const int cR = (nR + 31) >> 5, cC = (3 * nR + 31) >> 5;
    const int cResp = (rp.out.resp && rp.sky_rows_h <= 1) ? cR : 0;
    const int nT = 2 * cR + cC + cResp;                      // template columns (patched later: normal stores)
    const int nB = nT + (rp.sky_rows_h == 1 ? 3 * cR : 0);
    const int cB0 = kTemplates ? 0 : nT, cB1 = kSky ? nB : nT;   // the columns of this call
    const T* const bg = reinterpret_cast<const T*>(rp.row_bg);
    const size_t vstride = (size_t)rp.ngroups;
    for (int c = cB0 + warp; c < cB1; c += nw) {
        const T* src;
        T* dst;
        int len, e;
        bool stream = false;
        if (c < 2 * cR) {
            const int k = c >= cR;                           // 0: hit, 1: depth
            e = (c - k * cR) * 32 + lane, len = nR;
            src = bg + (size_t)k * nR;
            dst = reinterpret_cast<T*>(k ? rp.out.depth : (void*)rp.out.hit);
        } else if (c < 2 * cR + cC) {
            e = (c - 2 * cR) * 32 + lane, len = 3 * nR;
            src = bg + 2 * (size_t)nR;
            dst = reinterpret_cast<T*>(rp.out.rgb);
        } else if (c < nT) {
            e = (c - 2 * cR - cC) * 32 + lane, len = nR;
            src = bg + 5 * (size_t)nR;
            dst = reinterpret_cast<T*>(rp.out.resp);
        } else {
            const int k = (c - nT) / cR;
            e = (c - nT - k * cR) * 32 + lane, len = nR;
            src = tab + (size_t)k * nR;
            dst = k == 0 ? o_sky[0] : (k == 1 ? o_sky[1] : o_sky[2]);
            stream = true;
        }
        if (!dst || e >= len) continue;
        const T v = __ldg(src + e);
        T* d = dst + view0 * (size_t)len + e;
        ISY_BOUND((view0 + (size_t)(Hcur - 1) * vstride) * (size_t)len + e < (size_t)rp.P * rp.H * (size_t)len);
        if (stream) {
            for (int hh = 0; hh < Hcur; ++hh, d += vstride * (size_t)len) __stcs(d, v);
        } else {
            for (int hh = 0; hh < Hcur; ++hh, d += vstride * (size_t)len) *d = v;
        }
    }
}

// The rays that hit something: the winner buffer (Hcur rows of R sorted rays, u16 copy slots) is scanned, a slot per
// thread and trip; an occupied slot gets its hit, depth and colour stored.  Everything it looks up -- heading of the
// row, ray of the sorted position, triangle, distance, colour -- is in shared memory: the pass issues no global load
// (they would queue behind the SM's own stores), only the five 4-byte stores per hit.
__device__ void final_pass_patch(const RenderParams& rp, SmemLayout& sm, int p, int grp, int Hcur) {
    const int R = rp.R, n = Hcur * R;
    const float inv_R = 1.0f / (float)R;
    float* const o_rgb = reinterpret_cast<float*>(rp.out.rgb);
    float* const o_depth = reinterpret_cast<float*>(rp.out.depth);
    const size_t view0 = (size_t)p * rp.H + grp;
    // consecutive threads take consecutive slots: consecutive sorted positions are mostly consecutive rays, so the
    // stores of a clump of hits share their sectors
    for (int idx = threadIdx.x; idx < n; idx += kThreads) {
        const unsigned int id = sm.ras.best[idx];
        if (id == kNoCopy) continue;
#ifdef ISY_NO_PATCH_STORES
        if (id != 0x12345u) continue;   // timing experiment: the patches are not stored
#endif
        // row = idx / R: exact in float for idx < 2^16 ((idx + 0.5) / R is at least 0.5 / R from an integer)
        const int hs = (int)(((float)idx + 0.5f) * inv_R), pos = idx - hs * R;
        ISY_BOUND(hs < Hcur && pos < R && id < (unsigned)kMaxE);
        const int hh = sm.ras.hperm[hs];
        const int ray = sm.sorted_ray16[pos];
        const Entry& e = sm.tile[id];
        const size_t i = (view0 + (size_t)hh * rp.ngroups) * (size_t)R + (size_t)ray;
        ISY_BOUND(hh < Hcur && ray < R && i < (size_t)rp.P * rp.H * (size_t)R);
        if (rp.out.hit) __stcs(rp.out.hit + i, e.tri);
        if (o_depth) __stcs(o_depth + i, (float)__hiloint2double(e.dist_hi, e.dist_lo));
        if (o_rgb) {
            float* qq = o_rgb + 3 * i;
            __stcs(qq, sm.rgb[3 * id]), __stcs(qq + 1, sm.rgb[3 * id + 1]), __stcs(qq + 2, sm.rgb[3 * id + 2]);
        }
        if (rp.out.resp) __stcs(reinterpret_cast<float*>(rp.out.resp) + i, sm.resp1[id]);
    }
}

}  // namespace isy

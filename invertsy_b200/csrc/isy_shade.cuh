// Final pass of the rasteriser path: shade, sky, write.
#pragma once
#include "isy_raster.cuh"

namespace isy {

// The colours of 32 consecutive rays of a view are 96 consecutive floats.  Written as they sit in the lanes (lane = ray,
// one instruction per channel) every store would touch each of the 12 sectors of that span -- and the final pass is
// bound by the store sectors the L1 takes per cycle.  Three shuffles regroup them so that each store instruction
// covers one contiguous 128-byte third of the span: shuffle t moves channel t of every lane, lane L = 3n + j takes
// the n-th value of third (j - t) mod 3 (the counts match: thirds hold 11/11/10 values of a channel, lane classes
// 11/11/10 lanes), and the value a lane holds for third i is the one it got from shuffle (j - i) mod 3.
// blk = address of the first colour of the 32 rays, nfl = 3 * (rays of the block that exist).  All lanes call it.
#ifndef ISY_RGB_COALESCE
#define ISY_RGB_COALESCE 0
#endif
__device__ __forceinline__ void store_rgb_block(float* blk, int nfl, float cr, float cg, float cb, int lane) {
#if !ISY_RGB_COALESCE
    if (3 * lane < nfl) __stcs(blk + 3 * lane, cr), __stcs(blk + 3 * lane + 1, cg), __stcs(blk + 3 * lane + 2, cb);
    return;
#endif
    const int j = lane % 3, n = lane / 3;
    const int b1 = j >= 1 ? j - 1 : 2, b2 = j >= 2 ? 0 : j + 1;         // third served by shuffles 1 and 2 (shuffle 0: j)
    const int s0 = (j ? (32 * j + 2) / 3 : 0) + n;                        // first source of that third, + n
    const int s1 = (b1 ? (32 * b1 + 1) / 3 : 0) + n;
    const int s2 = (b2 ? (32 * b2) / 3 : 0) + n;
    const float v0 = __shfl_sync(0xffffffffu, cr, s0);
    const float v1 = __shfl_sync(0xffffffffu, cg, s1);
    const float v2 = __shfl_sync(0xffffffffu, cb, s2);
    const int m0 = 3 * s0, m1 = 3 * s1 + 1, m2 = 3 * s2 + 2;             // where they go in the 96 floats
    // third 0 <- shuffle j, third 1 <- shuffle (j - 1) mod 3 = b1, third 2 <- shuffle (j - 2) mod 3 = b2
    const float o0 = j == 0 ? v0 : (j == 1 ? v1 : v2);
    const int i0 = j == 0 ? m0 : (j == 1 ? m1 : m2);
    const float o1 = b1 == 0 ? v0 : (b1 == 1 ? v1 : v2);
    const int i1 = b1 == 0 ? m0 : (b1 == 1 ? m1 : m2);
    const float o2 = b2 == 0 ? v0 : (b2 == 1 ? v1 : v2);
    const int i2 = b2 == 0 ? m0 : (b2 == 1 ? m1 : m2);
    if (i0 < nfl) __stcs(blk + i0, o0);
    if (i1 < nfl) __stcs(blk + i1, o1);
    if (i2 < nfl) __stcs(blk + i2, o2);
}

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
            size_t view[kSU];
            unsigned int id[kSU];
#pragma unroll
            for (int q = 0; q < kSU; ++q) {
                Y[q] = 0, P[q] = 0, A[q] = CUDART_NAN;
                const int hh = min(hh0 + q, hend - 1);
                view[q] = (size_t)p * rp.H + grp + (size_t)hh * rp.ngroups;   // group = every ngroups-th heading
                const int hs = sm.ras.hinv[hh];
                id[q] = kSmem ? (unsigned)sm.ras.best[hs * Rrow + pos] : gbest[(size_t)hs * R + pos];
                if (perez) {
                    const double hsn = sm.head_sin[hh], hcs = sm.head_cos[hh];
                    const double sy = t4v.z * hcs + t4v.w * hsn, cy = t4v.w * hcs - t4v.z * hsn;   // yaw_e + heading
                    const SunConst& sc = kOneSun ? sm.sun0 : rp.sun[rp.sun_per_view ? view[q] : 0];
                    sky_eval<false>(rp.sky, sc, t4v.y * cy, t4v.y * sy, -t4v.x, e_pitch + kHalfPi, f_theta, Y[q], P[q], A[q],
                                    false, sm.gtab);
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
                    }
                } else {
                    // streaming stores: the views are written once and never read back by the kernel, so they
                    // should leave L2 before the slabs, the world and the eye tables do
                    const size_t i = off0 + q * ostride;
                    if (kAllOut || o_rgb) store_rgb_block(o_rgb + 3 * (i - lane), 3 * (r_lo + r_n - (r - lane)), cr, cg, cb, lane);
                    if (!valid) continue;
                    if (kAllOut || rp.out.hit) __stcs(rp.out.hit + i, hit);
                    if (kAllOut || o_depth) __stcs(o_depth + i, depth);
                    if (kAllOut || o_Y) __stcs(o_Y + i, (float)Y[q]);
                    if (kAllOut || o_P) __stcs(o_P + i, (float)P[q]);
                    if (kAllOut || o_A) __stcs(o_A + i, (float)A[q]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// final pass when every position of the call looks along the same row of headings under one sun
// (grids, heat maps: the scans of the reference's data-collection simulations).  The sky of view
// (position, heading) is then a function of (heading, ray) alone: sky_shared_kernel evaluates it once
// per launch into a [3][H][R] float table (Y, DoP, AoP; L2-resident) with the same sky_eval as the
// per-view pass, and this pass copies it: per (ray, heading) one winner look-up, three coalesced
// table loads, the colour, and the stores of the view.
// ------------------------------------------------------------------------------------------
// kSky = false: Y / DoP / AoP are not this pass's business (copy_sky_rows wrote them while the copies were rasterised)
template <bool kAllOut, int kSU, bool kSky = true>
__device__ void final_pass_tab(const RenderParams& rp, SmemLayout& sm, int p, int grp, int Hcur) {
    const int R = rp.R;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* const o_rgb = reinterpret_cast<float*>(rp.out.rgb);
    float* const o_depth = reinterpret_cast<float*>(rp.out.depth);
    float* const o_Y = reinterpret_cast<float*>(rp.out.Y);
    float* const o_P = reinterpret_cast<float*>(rp.out.dop);
    float* const o_A = reinterpret_cast<float*>(rp.out.aop);
    const size_t plane = (size_t)rp.H * (size_t)R;
    const float* const tY = rp.sky_tab, * const tP = rp.sky_tab + plane, * const tA = rp.sky_tab + 2 * plane;
    const float nanf_ = __int_as_float(0x7fc00000);
    const int nb = (R + 31) >> 5;
    int nch = min(Hcur, max(1, rp.shade_nch));
    int hper = (Hcur + nch - 1) / nch;
    hper = min(Hcur, (hper + kSU - 1) / kSU * kSU);      // whole trips
    nch = (Hcur + hper - 1) / hper;
    for (int u = warp; u < nb * nch; u += kWarps) {
        const int rb = u / nch, hc = u - rb * nch;
        const int r = rb * 32 + lane;
        const bool valid = r < R;
        const int rr = valid ? r : R - 1;
        const bool ground = rp.eye_py[2 * (size_t)rr] >= 0;                 // world.py:90
        const int pos = __ldg(rp.grid_pos_of + rr);
        const int hend = min(Hcur, (hc + 1) * hper);
        // element offsets of (heading hh0, this lane's ray) in the table and of its view in the outputs;
        // consecutive headings of a group are ngroups apart, so both advance by a constant
        const size_t tstride = (size_t)rp.ngroups * (size_t)R;
        size_t toff = ((size_t)grp + (size_t)(hc * hper) * rp.ngroups) * (size_t)R + (size_t)rr;
        size_t off0 = ((size_t)p * rp.H + grp + (size_t)(hc * hper) * rp.ngroups) * (size_t)R + (size_t)r;
        for (int hh0 = hc * hper; hh0 < hend; hh0 += kSU, toff += kSU * tstride, off0 += kSU * tstride) {
            float Y[kSU], P[kSU], A[kSU];
            unsigned int id[kSU];
#pragma unroll
            for (int q = 0; q < kSU; ++q) {
                const int hh = min(hh0 + q, hend - 1);
                id[q] = (unsigned)sm.ras.best[sm.ras.hinv[hh] * R + pos];
                if (kSky) {
                    const size_t ti = toff + (size_t)(hh - hh0) * tstride;
                    Y[q] = __ldg(tY + ti), P[q] = __ldg(tP + ti), A[q] = __ldg(tA + ti);
                }
            }
#pragma unroll
            for (int q = 0; q < kSU; ++q) {
                if (hh0 + q >= hend) break;
                float cr = nanf_, cg = nanf_, cb = nanf_, depth = nanf_;
                int hit = -1;
                if (ground) cr = (float)(229.0 / 255.0), cg = (float)(183.0 / 255.0), cb = (float)(90.0 / 255.0);
                if (id[q] != kNoCopy) {
                    const Entry& w = sm.tile[id[q]];
                    hit = w.tri;
                    depth = (float)__hiloint2double(w.dist_hi, w.dist_lo);
                    const float* col = rp.tri_rgb + 3 * (size_t)hit;
                    cr = __ldg(col + 0), cg = __ldg(col + 1), cb = __ldg(col + 2);
                }
                const size_t i = off0 + q * tstride;
                if (kAllOut || o_rgb) store_rgb_block(o_rgb + 3 * (i - lane), 3 * (R - (r - lane)), cr, cg, cb, lane);
                if (valid) {
                    if (kAllOut || rp.out.hit) __stcs(rp.out.hit + i, hit);
                    if (kAllOut || o_depth) __stcs(o_depth + i, depth);
                    if (kSky) {
                        if (kAllOut || o_Y) __stcs(o_Y + i, Y[q]);
                        if (kAllOut || o_P) __stcs(o_P + i, P[q]);
                        if (kAllOut || o_A) __stcs(o_A + i, A[q]);
                    }
                }
            }
        }
    }
}

// Y / DoP / AoP of the views of (position p, heading group grp) when the launch has a sky table: the rows of the table
// ARE the rows of those views.  Copied 16 bytes per lane (R a multiple of 4: rows are 16-byte aligned), four
// independent loads in flight per lane, by the warps [w0, w0 + nw) only: it has nothing to do with the winners, so a
// few warps do it while the others rasterise -- the SM's store path (32 B/clk, microbench/store_bw.cu) is idle then,
// and it is what bounds the final pass.
__device__ void copy_sky_rows(const RenderParams& rp, int p, int grp, int Hcur, int w0, int nw) {
    const int R4 = rp.R >> 2;
    const int lane = threadIdx.x & 31, warp = (int)(threadIdx.x >> 5) - w0;
    if (warp < 0 || warp >= nw) return;
    const size_t plane4 = (size_t)rp.H * (size_t)R4;
    const float4* const tab4 = reinterpret_cast<const float4*>(rp.sky_tab);
    float4* const dY = reinterpret_cast<float4*>(rp.out.Y), * const dP = reinterpret_cast<float4*>(rp.out.dop),
                * const dA = reinterpret_cast<float4*>(rp.out.aop);
    constexpr int kU = 4;                                   // loads in flight per lane
    for (int row = warp; row < 3 * Hcur; row += nw) {       // (plane, heading): one row of R4 float4
        const int k = row / Hcur, hg = grp + (row - k * Hcur) * rp.ngroups;
        float4* const dk = k == 0 ? dY : (k == 1 ? dP : dA);
        if (!dk) continue;
        const float4* const src = tab4 + (size_t)k * plane4 + (size_t)hg * R4;
        float4* const dst = dk + ((size_t)p * rp.H + hg) * (size_t)R4;
        for (int c0 = lane; c0 < R4; c0 += 32 * kU) {
            float4 v[kU];
#pragma unroll
            for (int q = 0; q < kU; ++q) {
                v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + 32 * q < R4) v[q] = __ldg(src + c0 + 32 * q);
            }
#pragma unroll
            for (int q = 0; q < kU; ++q)
                if (c0 + 32 * q < R4) __stcs(dst + c0 + 32 * q, v[q]);
        }
    }
}

// The same for eyes whose ray count is a multiple of 4 (rows of the outputs are then 16-byte aligned): everything
// moves 16 bytes per lane.
//   1. Y / DoP / AoP: the rows of the table ARE the rows of the views -- copied float4 by float4, four independent
//      loads in flight per lane, no winner look-up involved.
//   2. hit / depth / rgb: a lane takes 4 consecutive rays of one heading (four independent winner -> tile -> colour
//      chains), a warp 128 rays: one 16-byte store each for hit and depth and three for the 12 colour floats, every
//      one of them covering contiguous memory.
// A quarter of the load/store instructions of the ray-per-lane pass, all of them fully coalesced.
template <bool kAllOut>
__device__ void final_pass_tab4(const RenderParams& rp, SmemLayout& sm, int p, int grp, int Hcur) {
    const int R = rp.R, R4 = R >> 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float nanf_ = __int_as_float(0x7fc00000);
    {
        const size_t plane4 = (size_t)rp.H * (size_t)R4;
        const float4* const tab4 = reinterpret_cast<const float4*>(rp.sky_tab);
        float4* const dY = reinterpret_cast<float4*>(rp.out.Y), * const dP = reinterpret_cast<float4*>(rp.out.dop),
                    * const dA = reinterpret_cast<float4*>(rp.out.aop);
        constexpr int kU = 4;                                   // loads in flight per lane
        for (int row = warp; row < 3 * Hcur; row += kWarps) {   // (plane, heading): one row of R4 float4
            const int k = row / Hcur, hg = grp + (row - k * Hcur) * rp.ngroups;
            float4* const dk = k == 0 ? dY : (k == 1 ? dP : dA);
            if (!kAllOut && !dk) continue;
            const float4* const src = tab4 + (size_t)k * plane4 + (size_t)hg * R4;
            float4* const dst = dk + ((size_t)p * rp.H + hg) * (size_t)R4;
            for (int c0 = lane; c0 < R4; c0 += 32 * kU) {
                float4 v[kU];
#pragma unroll
                for (int q = 0; q < kU; ++q) {
                    v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c0 + 32 * q < R4) v[q] = __ldg(src + c0 + 32 * q);
                }
#pragma unroll
                for (int q = 0; q < kU; ++q)
                    if (c0 + 32 * q < R4) __stcs(dst + c0 + 32 * q, v[q]);
            }
        }
    }
    if (!kAllOut && !rp.out.rgb && !rp.out.hit && !rp.out.depth) return;
    const int nb = (R + 127) >> 7;
    for (int u = warp; u < nb * Hcur; u += kWarps) {
        const int hh = u / nb, rb = u - hh * nb;
        const int r0 = rb * 128 + lane * 4;                     // this lane's rays r0 .. r0 + 3 (all or none exist)
        const bool valid = r0 < R;
        const int rr = valid ? r0 : R - 4;
        const int4 pos4 = __ldg(reinterpret_cast<const int4*>(rp.grid_pos_of + rr));
        const unsigned short* const brow = sm.ras.best + (int)sm.ras.hinv[hh] * R;
        const unsigned int id[4] = {brow[pos4.x], brow[pos4.y], brow[pos4.z], brow[pos4.w]};
        float c[4][3], depth[4];
        int hit[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool ground = rp.eye_py[2 * (size_t)(rr + k)] >= 0;        // world.py:90
            c[k][0] = ground ? (float)(229.0 / 255.0) : nanf_;
            c[k][1] = ground ? (float)(183.0 / 255.0) : nanf_;
            c[k][2] = ground ? (float)(90.0 / 255.0) : nanf_;
            depth[k] = nanf_, hit[k] = -1;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (id[k] != kNoCopy) {
                const Entry& w = sm.tile[id[k]];
                hit[k] = w.tri;
                depth[k] = (float)__hiloint2double(w.dist_hi, w.dist_lo);
                const float* col = rp.tri_rgb + 3 * (size_t)hit[k];
                c[k][0] = __ldg(col + 0), c[k][1] = __ldg(col + 1), c[k][2] = __ldg(col + 2);
            }
        }
        if (valid) {
            const size_t i = ((size_t)p * rp.H + grp + (size_t)hh * rp.ngroups) * (size_t)R + (size_t)r0;
            if (kAllOut || rp.out.hit) __stcs(reinterpret_cast<int4*>(rp.out.hit + i), make_int4(hit[0], hit[1], hit[2], hit[3]));
            if (kAllOut || rp.out.depth)
                __stcs(reinterpret_cast<float4*>(reinterpret_cast<float*>(rp.out.depth) + i),
                       make_float4(depth[0], depth[1], depth[2], depth[3]));
            if (kAllOut || rp.out.rgb) {
                float4* q = reinterpret_cast<float4*>(reinterpret_cast<float*>(rp.out.rgb) + 3 * i);
                __stcs(q + 0, make_float4(c[0][0], c[0][1], c[0][2], c[1][0]));
                __stcs(q + 1, make_float4(c[1][1], c[1][2], c[2][0], c[2][1]));
                __stcs(q + 2, make_float4(c[2][2], c[3][0], c[3][1], c[3][2]));
            }
        }
    }
}

}  // namespace isy

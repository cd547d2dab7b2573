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
                } else if (valid) {
                    // streaming stores: the views are written once and never read back by the kernel, so they
                    // should leave L2 before the slabs, the world and the eye tables do
                    const size_t i = off0 + q * ostride;
                    if (kAllOut || o_rgb) {
                        float* qq = o_rgb + 3 * i;
                        __stcs(qq, cr), __stcs(qq + 1, cg), __stcs(qq + 2, cb);
                    }
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

}  // namespace isy

// Rasterisers for yaw-only views: tile walk and column walk (evenly spaced headings), pitch-run walk, ray-grid walk.
#pragma once
#include "isy_bins.cuh"

namespace isy {

// ------------------------------------------------------------------------------------------
// phase A3' + B' (yaw-only views): triangle-centric rasterisation over the eye's pitch-sorted rays.
//
// For a yaw-only view the query of ray r is (pitch_r, yaw_r + heading).  A projected copy with
// bounding box [t0,t1] x [p0,p1] can only contain rays whose pitch lies in [t0,t1] -- a contiguous
// run of the eye's rays sorted by pitch (static per eye, isy_eye_create) -- and, for such a ray,
// only the headings h with yaw_r + h in [p0,p1]: a window of the group's sorted headings found
// through a small look-up table.  One warp takes a copy, its lanes the rays of the pitch run;
// each (ray, heading) that survives gets the fp32 edge test (+ exact re-check) and posts the
// copy's distance rank with an atomic min into best[heading][sorted ray].
// The work is proportional to (copies x rays in their pitch band) + (true candidate pairs) for
// ALL headings of the position at once; nothing is built per position beyond the ranks.
// ------------------------------------------------------------------------------------------
constexpr unsigned int kNoCopy = 0xffffu;

// atomicAdd on a __shared__ int through an explicit shared-space instruction (a generic-address
// atomic on shared memory is much slower)
__device__ __forceinline__ int smem_atomic_add(int* p, int v) {
    int old;
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(p);
    asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
    return old;
}
#ifndef ISY_RU
#define ISY_RU 2
#endif
constexpr int kRU = ISY_RU;           // rays per lane and iteration in the rasteriser

// is copy e (distance d, triangle tri) nearer than copy c? (ties: lower triangle index, then lower slot)
// The upper word of the fp64 distance decides almost always (positive doubles order like their bit patterns): one
// shared-memory word and an integer compare; equal upper words take the full comparison.
__device__ __forceinline__ bool nearer(const Entry* tile, double d, int tri, unsigned int e, unsigned int c) {
    const Entry& o = tile[c];
    const int dh = __double2hiint(d);
    if (dh != o.dist_hi) return dh < o.dist_hi;
    const double dc = __hiloint2double(o.dist_hi, o.dist_lo);
    return d < dc || (d == dc && (tri < o.tri || (tri == o.tri && e < c)));
}

// best[idx] = nearer(e, best[idx]) ? e : best[idx], atomically (CAS on the 32-bit word holding the u16).
// best_s is the SHARED-space address of best[0]: explicit ld.shared / atom.shared keep the address
// arithmetic to two instructions (a generic pointer makes the compiler rebuild the window base every time).
__device__ __forceinline__ void post_copy_u16(unsigned int best_s, unsigned int idx, const Entry* tile, double d, int tri,
                                              unsigned int e) {
    ISY_BOUND(idx < (unsigned)kBestCap && e < (unsigned)kMaxE);
    const unsigned int addr = best_s + ((idx >> 1) << 2);
    const bool hi = idx & 1u;
    unsigned int old;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(old) : "r"(addr) : "memory");
    for (;;) {
        const unsigned int cur = hi ? (old >> 16) : (old & 0xffffu);
        if (cur != kNoCopy && !nearer(tile, d, tri, e, cur)) return;
        const unsigned int nw = hi ? ((old & 0xffffu) | (e << 16)) : ((old & 0xffff0000u) | e);
        unsigned int prev;
        asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(prev) : "r"(addr), "r"(old), "r"(nw) : "memory");
        if (prev == old) return;
        old = prev;
    }
}

__device__ __forceinline__ void post_copy_u32(unsigned int* slot, const Entry* tile, double d, int tri, unsigned int e) {
    unsigned int old = *reinterpret_cast<volatile unsigned int*>(slot);
    for (;;) {
        if (old != kNoCopy && !nearer(tile, d, tri, e, old)) return;
        const unsigned int prev = atomicCAS(slot, old, e);
        if (prev == old) return;
        old = prev;
    }
}

// sorts the group's wrapped headings (head_val[0..Hcur)); returns true (to every thread) when they are evenly
// spaced over the full circle -- every scan of the reference's simulations, and a single heading: the column /
// tile walk applies; otherwise it also builds the azimuth look-up table of the pitch-run walk
__device__ bool sort_headings(SmemLayout& sm, const Team& tm, int Hcur) {
    const int tid = threadIdx.x;
    if (tid < Hcur) {
        const float h = (float)sm.head_val[tid];
        int pos = 0;
        for (int k = 0; k < Hcur; ++k) {
            const float hk = (float)sm.head_val[k];
            pos += (hk < h) || (hk == h && k < tid);
        }
        sm.ras.hsort[pos] = h;
        sm.ras.hsort[pos + Hcur] = h + (float)kTwoPi;
        sm.ras.hvalf[pos] = h;
        sm.ras.hvalf[pos + Hcur] = h;
        sm.ras.hperm[pos] = (unsigned char)tid;
        sm.ras.hperm[pos + Hcur] = (unsigned char)tid;
        sm.ras.hinv[tid] = (unsigned char)pos;
    }
    tm.sync();
    const float h0 = sm.ras.hsort[0];
    const float delta = (float)kTwoPi / (float)Hcur;
    const bool mine = tid >= Hcur || fabsf(sm.ras.hsort[tid] - (h0 + tid * delta)) < 5e-7f;
    const bool uniform = tm.sync_and(mine);
    if (tid == 0) sm.ras.uni[0] = h0, sm.ras.uni[1] = delta, sm.ras.uni[2] = 1.0f / delta, sm.ras.uni[3] = uniform ? 1.f : 0.f;
    if (!uniform) {
        for (int s = tid; s < kHeadLut; s += tm.n) {
            // first sorted slot whose heading is not below this azimuth slot (minus a guard band):
            // windows start their walk here; whatever they pick up below `lower` just fails the edge test
            const float start = -(float)kPi + s * ((float)kTwoPi / kHeadLut) - 1e-4f;
            int k = 0;
            while (k < 2 * Hcur && sm.ras.hsort[k] < start) ++k;
            sm.ras.hlut[s] = (unsigned short)k;
        }
    }
    return uniform;
}

// One step of the pitch-run walk: RU rays per lane (32 * RU consecutive pitch-sorted rays from jb) against copy e.
// Their window walks and edge tests are independent instruction streams, which is what hides the
// shared-memory and branch latencies here.
template <int RU>
__device__ __forceinline__ void walk_chunk(const RenderParams& rp, SmemLayout& sm, const Exact* exacts,
                                           const Entry& ent, const float4 bb, double dist, int e, unsigned jb, unsigned jend,
                                           float glo, float width, int Hcur, int R, unsigned jbase) {
    // R: rays per row of the winner buffer (the eye, or the band of a large eye); jbase: first sorted position of the band
    const int lane = threadIdx.x & 31, H2 = 2 * Hcur;
    const float pi_f = (float)kPi, twopi_f = (float)kTwoPi;
    const unsigned int best_s = (unsigned int)__cvta_generic_to_shared(sm.ras.best);
    float2 py[RU];
    int k[RU], ks[RU];    // window cursor; sorted slot of the heading under the cursor
    float upper[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
        const unsigned j = jb + u * 32 + lane;
        py[u] = make_float2(4.f, 0.f);                  // pitch 4 rad: outside every band
        ISY_BOUND(jend <= (unsigned)kGridRaysSmem);
        if (j < jend) py[u] = sm.ras.sorted[j];
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
        // headings h with yaw + h in [glo, ghi]  <=>  h in [lower, lower + width] on the circle
        float lower = glo - py[u].y;
        if (lower < -pi_f) lower += twopi_f;
        else if (lower >= pi_f) lower -= twopi_f;
        const bool band = py[u].x >= bb.x && py[u].x <= bb.y;
        const int ls = min(max((int)((lower + pi_f) * (kHeadLut / twopi_f)), 0), kHeadLut - 1);
        k[u] = band ? (int)sm.ras.hlut[ls] : H2;
        upper[u] = lower + width;
    }
    // lanes walk their heading windows in lock step (most windows hold 0 or 1 heading)
    for (;;) {
        bool act[RU], any = false;
        float hv[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int kk = min(k[u], H2 - 1);
            act[u] = k[u] < H2 && sm.ras.hsort[kk] <= upper[u];
            hv[u] = sm.ras.hvalf[kk];
            ks[u] = kk >= Hcur ? kk - Hcur : kk;           // the + 2 pi copy is the same heading
            any |= act[u];
        }
        if (!__any_sync(0xffffffffu, any)) break;
        float fy[RU], m[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) {                 // branch-free: the common case of every slot
            float f = py[u].y + hv[u];                  // both in (-pi, pi]: one shift wraps
            f = f > pi_f ? f - twopi_f : (f <= -pi_f ? f + twopi_f : f);
            fy[u] = f;
            const float e0 = fmaf(ent.A0, py[u].x, fmaf(ent.B0, f, ent.C0));
            const float e1 = fmaf(ent.A1, py[u].x, fmaf(ent.B1, f, ent.C1));
            const float e2 = fmaf(ent.A2, py[u].x, fmaf(ent.B2, f, ent.C2));
            m[u] = fminf(e0, fminf(e1, e2));
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const unsigned j = jb + u * 32 + lane;
            const bool seam = fabsf(fy[u]) > 3.14159f;
            bool in = act[u] && m[u] > kEdgeEps && !seam;
            // rare: the rounding band of an edge, or a query within 3e-6 rad of the seam ->
            // redo the query in fp64 exactly as the reference composes it
            if (act[u] && (seam || (m[u] >= -kEdgeEps && m[u] <= kEdgeEps))) {
                const int hh = sm.ras.hperm[ks[u]];
                const int ray = __ldg(rp.grid_sorted_ray + jbase + j);
                double yg = rp.eye_py[2 * (size_t)ray + 1] + sm.head_val[hh];
                if (yg > kPi) yg -= kTwoPi;
                else if (yg <= -kPi) yg += kTwoPi;
                const float f = (float)yg;
                const float e0 = fmaf(ent.A0, py[u].x, fmaf(ent.B0, f, ent.C0));
                const float e1 = fmaf(ent.A1, py[u].x, fmaf(ent.B1, f, ent.C1));
                const float e2 = fmaf(ent.A2, py[u].x, fmaf(ent.B2, f, ent.C2));
                const float mm = fminf(e0, fminf(e1, e2));
                in = mm > kEdgeEps;
                if (!in && mm >= -kEdgeEps) in = inside_exact(exacts[e], rp.eye_py[2 * (size_t)ray], yg);
            }
            // rows of best[] are SORTED heading slots
            if (in) {
                post_copy_u16(best_s, (unsigned)ks[u] * R + j, sm.tile, dist, ent.tri, e);
            }
            if (act[u]) ++k[u];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Column walk (evenly spaced headings over the full circle: every scan of the reference's simulations).
// With headings h0 + k*delta, delta = 2 pi / H, the azimuth of (ray j, heading k) is
//     yaw_j + h0 + k*delta  =  phase_j + (b_j + k)*delta - pi   (mod 2 pi),   phase_j in [0, delta),
// so the H queries of a ray fall one into each of the H azimuth columns [c*delta - pi, (c+1)*delta - pi), and column
// c holds every ray exactly once -- at heading k = (c - b_j) mod H.  A copy is therefore tested against
// (rays of its pitch run) x (columns its azimuth range touches): no window set-up per ray, no walk over headings,
// all lanes busy.  E_k(pitch, phase + off) = (A pitch + B phase + C) + B off: the bracket is formed once per ray, a
// column costs one FMA per edge.  build_phases makes the tables once per CTA while the headings stay the same.
// fp32 budget of the edge value: coefficients 3 pi u, pitch pi/2 u, phase + column offset (delta + pi) u, three FMA
// roundings (delta + 2 pi) u + (2.5 pi + delta) u + 3.5 pi u  ->  2.4e-6, plus the 7.5e-7 between h0 + k*delta
// and the true heading admitted by the uniformity check: 3.1e-6 < kEdgeEps.
// ------------------------------------------------------------------------------------------
__device__ void build_phases(const RenderParams& rp, SmemLayout& sm, const Team& tm, int Hcur, int r_n, unsigned jbase) {
    const double h0 = sm.head_val[sm.ras.hperm[0]];
    const bool same = sm.scan.ph_H == Hcur && sm.scan.ph_lo == (int)jbase && sm.scan.ph_h0 == h0;
    tm.sync();
    if (same) return;
    const double delta = kTwoPi / (double)Hcur, inv = (double)Hcur / kTwoPi;
    for (int j = threadIdx.x; j < r_n; j += tm.n) {
        const int ray = __ldg(rp.grid_sorted_ray + jbase + j);
        const double x = rp.eye_py[2 * (size_t)ray + 1] + h0 + kPi;      // (-pi, 3 pi]
        double b = floor(x * inv), phi = x - b * delta;
        if (phi < 0.0) phi += delta, b -= 1.0;
        else if (phi >= delta) phi -= delta, b += 1.0;
        int bm = (int)b % Hcur;
        if (bm < 0) bm += Hcur;
        sm.scan.phase[j] = (float)phi;
        sm.scan.colb[j] = (unsigned char)bm;
    }
    for (int i = threadIdx.x; i < Hcur + 2; i += tm.n) sm.scan.coloff[i] = (float)((double)(i - 1) * delta - kPi);
    if (threadIdx.x == 0) sm.scan.ph_H = Hcur, sm.scan.ph_lo = (int)jbase, sm.scan.ph_h0 = h0;
}

// 32 * RU consecutive pitch-sorted rays from jb against copy e in columns c_lo .. c_hi (-1 and H are the aliases of
// H - 1 and 0 beyond the seam: only the queries within 3e-4 rad of +-pi are looked at there, in fp64)
template <int RU>
__device__ __forceinline__ void column_chunk(const RenderParams& rp, SmemLayout& sm, const Exact* exacts, const Entry& ent,
                                             int e, unsigned jb, unsigned jend, int c_lo, int c_hi, int Hcur, int R,
                                             unsigned jbase) {
    const int lane = threadIdx.x & 31;
    const unsigned int best_s = (unsigned int)__cvta_generic_to_shared(sm.ras.best);
    float g0[RU], g1[RU], g2[RU], ph[RU];
    const float ninf = __int_as_float(0xff800000);
#pragma unroll
    for (int u = 0; u < RU; ++u) {
        const unsigned j = jb + u * 32 + lane;
        ISY_BOUND(jend <= (unsigned)kGridRaysSmem);
        const bool ok = j < jend;
        const float pitch = ok ? sm.ras.sorted[j].x : 0.f;
        ph[u] = ok ? sm.scan.phase[j] : 0.f;
        g0[u] = ok ? fmaf(ent.A0, pitch, fmaf(ent.B0, ph[u], ent.C0)) : ninf;
        g1[u] = fmaf(ent.A1, pitch, fmaf(ent.B1, ph[u], ent.C1));
        g2[u] = fmaf(ent.A2, pitch, fmaf(ent.B2, ph[u], ent.C2));
    }
    for (int c = c_lo; c <= c_hi; ++c) {
        const float off = sm.scan.coloff[c + 1];
        const bool edge = c <= 0 || c >= Hcur - 1;          // columns that touch the seam
        float m[RU];
        bool cand[RU], any = false;
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            m[u] = fminf(fmaf(ent.B0, off, g0[u]), fminf(fmaf(ent.B1, off, g1[u]), fmaf(ent.B2, off, g2[u])));
            cand[u] = m[u] >= -kEdgeEps;
        }
        if (edge) {
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                const float af = fabsf(ph[u] + off);
                cand[u] = jb + u * 32 + lane < jend && af <= 3.1419f && (af > 3.14159f || cand[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) any |= cand[u];
        if (!__any_sync(0xffffffffu, any)) continue;
        const int cw = c < 0 ? c + Hcur : (c >= Hcur ? c - Hcur : c);
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            if (!cand[u]) continue;
            const unsigned j = jb + u * 32 + lane;
            int ks = cw - (int)sm.scan.colb[j];               // sorted slot of the heading that puts ray j into column c
            if (ks < 0) ks += Hcur;
            const bool seam = edge && fabsf(ph[u] + off) > 3.14159f;
            bool in = m[u] > kEdgeEps && !seam;
            if (!in) {
                // the rounding band of an edge, or a query within 3e-6 rad of the seam: redo the query in fp64
                // exactly as the reference composes it
                const int hh = sm.ras.hperm[ks];
                const int ray = __ldg(rp.grid_sorted_ray + jbase + j);
                double yg = rp.eye_py[2 * (size_t)ray + 1] + sm.head_val[hh];
                if (yg > kPi) yg -= kTwoPi;
                else if (yg <= -kPi) yg += kTwoPi;
                const float f = (float)yg, pitch = sm.ras.sorted[j].x;
                const float e0 = fmaf(ent.A0, pitch, fmaf(ent.B0, f, ent.C0));
                const float e1 = fmaf(ent.A1, pitch, fmaf(ent.B1, f, ent.C1));
                const float e2 = fmaf(ent.A2, pitch, fmaf(ent.B2, f, ent.C2));
                const float mm = fminf(e0, fminf(e1, e2));
                in = mm > kEdgeEps;
                if (!in && mm >= -kEdgeEps) in = inside_exact(exacts[e], rp.eye_py[2 * (size_t)ray], yg);
            }
            if (in) post_copy_u16(best_s, (unsigned)ks * R + j, sm.tile, __hiloint2double(ent.dist_hi, ent.dist_lo), ent.tri, e);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Tile walk (the same evenly spaced headings): the queries of a position are cut into tiles of kTileRays consecutive
// pitch-sorted rays x one azimuth column.  The copies are first filed into the tiles their pitch run and azimuth
// range touch (count, scan, fill: lists of Entry byte offsets kept where bbox / order were); then a warp takes a tile,
// its lanes kTileU queries each, and tests the tile's copies one after the other -- every load is a broadcast, the
// nearest containing copy of a query stays in registers, and the winner buffer is written once per hit, without
// atomics.  A position whose lists would not fit falls back to the column walk.
// The fp32 query azimuth is fl(phase + column offset): three roundings (< 2.5e-7) + 7.5e-7 between h0 + k*delta and
// the true heading: inside the 1.9e-6 that kEdgeEps admits for it (isy_common.cuh).
// ------------------------------------------------------------------------------------------
// one lane of the warp takes n items from a shared-memory counter (elect.sync keeps ptxas from wrapping the atomic
// into its warp-aggregation sequence); returns the first to every lane
__device__ __forceinline__ int warp_fetch(int* counter, int n) {
    int v = 0;
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(counter);
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t@p atom.shared.add.s32 %0, [%1], %2;\n\t}"
                 : "+r"(v) : "r"(a), "r"(n) : "memory");
    return __shfl_sync(0xffffffffu, v, 0);
}

template <bool kFill>
__device__ __forceinline__ void bin_add(SmemLayout& sm, unsigned short* tlist, int tile, int e) {
    ISY_BOUND(tile >= 0 && tile < kMaxTiles);
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(&sm.scan.tile_cnt[tile]);
    if (kFill) {
        int at;
        asm volatile("atom.shared.add.s32 %0, [%1], 1;" : "=r"(at) : "r"(a) : "memory");
        ISY_BOUND(at >= 0 && at < kMaxE * 24);
#ifdef ISY_DEBUG_BOUNDS
        if (at < 0 || at >= kMaxE * 24) return;
#endif
        tlist[at] = (unsigned short)(e * (int)sizeof(Entry));
    } else {
        asm volatile("red.shared.add.s32 [%0], 1;" ::"r"(a) : "memory");    // counting: nothing comes back
    }
}

// One pass over the copies: kFill = false works out the columns of every copy (from its box, which the lists may
// overwrite later) and counts the copies of every tile; kFill = true files them.  Lane l of warp w takes copies
// l * (warps) + w, ...: neighbouring slots hold neighbouring triangles -- the large copies of a tussock next to the
// eye come in runs -- and this way every warp gets its share of them.  The (copy, tile) pairs of a warp's 32 copies
// are then lined up by a prefix sum and dealt one per lane and trip, whatever the sizes of the copies.
template <bool kFill>
__device__ void bin_pass(SmemLayout& sm, const Team& tm, int E, int Hcur, int q_h, unsigned short* tlist) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = tm.n >> 5;
    const float inv_delta = sm.ras.uni[2];
    for (int base = 0; base < E; base += tm.n) {   // warp-uniform trip count
        const int e = base + lane * nw + warp;
        int q0 = 0, nq = 0, c0 = 0, nc = 0;
        if (e < E) {
            unsigned int cr;
            if (kFill) {
                cr = sm.scan.crange[e];
            } else {
                const float4 bb = sm.bbox[e];
                cr = copy_columns(bb.z, bb.w, inv_delta, Hcur);
                sm.scan.crange[e] = (unsigned short)cr;
            }
            const unsigned int jr = sm.jrange[e];
            const unsigned jbeg = jr & 0xffffu, jend = jr >> 16;
            if (jbeg < jend)
                q0 = (int)(jbeg >> kTileShift), nq = (int)((jend - 1) >> kTileShift) - q0 + 1, c0 = (int)(cr & 0xffu) - 1, nc = (int)(cr >> 8);
        }
        const unsigned nt = (unsigned)(nq * nc);
        const float rnc = 1.0f / (float)max(nc, 1);
        unsigned inc = nt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += v;
        }
        const unsigned total = __shfl_sync(0xffffffffu, inc, 31), exc = inc - nt;
        for (unsigned tb = 0; tb < total; tb += 32) {          // warp-uniform trip count
            const unsigned t = tb + lane;
            int k = 0;                                         // the lane whose copy holds pair t: smallest k with inc_k > t
#pragma unroll
            for (int b2 = 16; b2 > 0; b2 >>= 1) {
                const unsigned v = __shfl_sync(0xffffffffu, inc, k + b2 - 1);
                if (v <= t) k += b2;
            }
            k = min(k, 31);
            const unsigned ek = __shfl_sync(0xffffffffu, exc, k);
            const int se = __shfl_sync(0xffffffffu, e, k), sq0 = __shfl_sync(0xffffffffu, q0, k);
            const int sc0 = __shfl_sync(0xffffffffu, c0, k), snc = __shfl_sync(0xffffffffu, nc, k);
            const float srnc = __shfl_sync(0xffffffffu, rnc, k);
            if (t < total) {
                const int i = (int)(t - ek);
                const int dq = (int)(((float)i + 0.5f) * srnc);        // i / snc, exact for these sizes
                int c = sc0 + (i - dq * snc);
                c = c < 0 ? c + Hcur : (c >= Hcur ? c - Hcur : c);
                bin_add<kFill>(sm, tlist, chunk_place(sq0 + dq, q_h) * Hcur + c, se);
            }
        }
    }
}

// where a position's tile lists go: over bbox + order when nothing needs the boxes any more (one group of headings,
// no bands), else into what the E copies leave free of the Entry tile
__device__ __forceinline__ unsigned short* tile_lists(SmemLayout& sm, int E, bool boxes_free, int& cap) {
    const int tail = (kMaxE - E) * (int)(sizeof(Entry) / sizeof(unsigned short));
    if (boxes_free && tail < kTileListCap) {
        cap = kTileListCap;
        return reinterpret_cast<unsigned short*>(sm.bbox);
    }
    cap = tail;
    return reinterpret_cast<unsigned short*>(sm.tile + E);
}

__device__ bool bin_copies(SmemLayout& sm, const Team& tm, int E, int Hcur, int Rrow, unsigned int* s_warp, unsigned short* tlist,
                           int cap, long long* tdbg = nullptr) {
    const long long td0 = clock64();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NQ = (Rrow + kTileRays - 1) >> kTileShift, NT = NQ * Hcur, q_h = min(NQ - 1, NQ / 2);
    ISY_BOUND(NT <= kMaxTiles && NT <= 2 * tm.n);
    for (int i = tid; i <= NT; i += tm.n) sm.scan.tile_cnt[i] = 0;
    tm.sync();
    bin_pass<false>(sm, tm, E, Hcur, q_h, tlist);
    const long long td1 = clock64();
    tm.sync();
    const long long td2 = clock64();
    // exclusive scan of the counts (two tiles per thread)
    const int i0 = 2 * tid;
    const unsigned a = i0 < NT ? sm.scan.tile_cnt[i0] : 0u, b = i0 + 1 < NT ? sm.scan.tile_cnt[i0 + 1] : 0u;
    unsigned incl = a + b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    tm.sync();
    unsigned w = lane < (tm.n >> 5) ? s_warp[lane] : 0u, before = lane < warp ? w : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        w += __shfl_xor_sync(0xffffffffu, w, d);
        before += __shfl_xor_sync(0xffffffffu, before, d);
    }
    const unsigned total = w;
    if (total > (unsigned)cap) return false;
    const unsigned excl = before + incl - (a + b);
    if (i0 < NT) sm.scan.tile_cnt[i0] = excl;
    if (i0 + 1 < NT) sm.scan.tile_cnt[i0 + 1] = excl + a;
    tm.sync();
    const long long td3 = clock64();
    bin_pass<true>(sm, tm, E, Hcur, q_h, tlist);      // tile_cnt[t] ends as the END of tile t's list
    if (tdbg) tdbg[0] += td1 - td0, tdbg[1] += td2 - td1, tdbg[2] += clock64() - td3;   // (profiling: count, wait, fill of thread 0)
    return true;
}

// the winner a query holds while its tile is walked
struct TileBest {
    unsigned int off, lo;   // Entry byte offset of the nearest containing copy (0xffffffff: none), low word of its distance
    int hi, tri;            // high word of the distance, triangle index
};

// copy `eo` may contain the query (edge value m within the rounding band or above, or a query at the seam): settle
// it -- in fp64 where the fp32 value cannot -- and keep the copy if it is nearer than what the query holds
__device__ __forceinline__ void tile_candidate(const RenderParams& rp, SmemLayout& sm, const Exact* exacts, const float4& a,
                                               const float4& b, const float4& d, unsigned eo, float pitch, float m, bool seam,
                                               int ks, unsigned j, unsigned jbase, TileBest& w) {
    bool in = m > kEdgeEps && !seam;
    if (!in) {
        // the rounding band of an edge, or a query within 3e-6 rad of the seam: redo the query in fp64 exactly as
        // the reference composes it
        const int hh = sm.ras.hperm[ks];
        const int ray = __ldg(rp.grid_sorted_ray + jbase + j);
        double yg = rp.eye_py[2 * (size_t)ray + 1] + sm.head_val[hh];
        if (yg > kPi) yg -= kTwoPi;
        else if (yg <= -kPi) yg += kTwoPi;
        const float fe = (float)yg;
        const float mm = fminf(fmaf(a.x, pitch, fmaf(a.y, fe, a.z)),
                               fminf(fmaf(a.w, pitch, fmaf(b.x, fe, b.y)), fmaf(b.z, pitch, fmaf(b.w, fe, d.x))));
        in = mm > kEdgeEps;
        if (!in && mm >= -kEdgeEps) in = inside_exact(exacts[eo / (unsigned)sizeof(Entry)], rp.eye_py[2 * (size_t)ray], yg);
    }
    if (in) {
        // nearer? (distance, then triangle index, then slot: as nearer())
        const int dhi = __float_as_int(d.w), tri = __float_as_int(d.y);
        const unsigned dlo = (unsigned)__float_as_int(d.z);
        bool nr = dhi < w.hi;                  // the upper word decides almost always
        if (dhi == w.hi) nr = dlo < w.lo || (dlo == w.lo && (tri < w.tri || (tri == w.tri && eo < w.off)));
        if (nr) w.hi = dhi, w.lo = dlo, w.tri = tri, w.off = eo;
    }
}

// one copy against the kTileU queries of every lane: edge values, and whether any may be inside
struct TileCopy {
    float4 a, b, d;     // the Entry: (A0 B0 C0 A1) (B1 C1 A2 B2) (C2 tri dist_lo dist_hi)
    unsigned eo;        // its byte offset in the tile
    float m[kTileU];
};
__device__ __forceinline__ void tile_load(TileCopy& t, const char* tile_b, unsigned eo) {
    t.eo = eo;
    t.a = *reinterpret_cast<const float4*>(tile_b + eo);
    t.b = *reinterpret_cast<const float4*>(tile_b + eo + 16);
    t.d = *reinterpret_cast<const float4*>(tile_b + eo + 32);
}
__device__ __forceinline__ float tile_eval(TileCopy& t, const float (&pitch)[kTileU], const float (&f)[kTileU]) {
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int u = 0; u < kTileU; ++u) {
        t.m[u] = fminf(fmaf(t.a.x, pitch[u], fmaf(t.a.y, f[u], t.a.z)),
                       fminf(fmaf(t.a.w, pitch[u], fmaf(t.b.x, f[u], t.b.y)), fmaf(t.b.z, pitch[u], fmaf(t.b.w, f[u], t.d.x))));
        mx = fmaxf(mx, t.m[u]);
    }
    return mx;      // (NaN edge values -- lanes beyond the eye -- never raise it)
}

// the copies [i0, i1) of a tile's list against the queries of the tile; kSeam: some query of the tile lies within
// 3e-6 rad of the seam and meets every copy in fp64
template <bool kSeam>
__device__ __forceinline__ void tile_list(const RenderParams& rp, SmemLayout& sm, const Exact* exacts, const unsigned short* lp,
                                          const unsigned short* le, const float (&pitch)[kTileU], const float (&f)[kTileU],
                                          const bool (&seam)[kTileU], const int (&ks)[kTileU], unsigned j0, unsigned jbase,
                                          TileBest (&w)[kTileU]) {
    const int lane = threadIdx.x & 31;
    const char* const tile_b = reinterpret_cast<const char*>(sm.tile);
    for (; lp < le; ++lp) {
        TileCopy c0;
        tile_load(c0, tile_b, lp[0]);
        const float mx = tile_eval(c0, pitch, f);
        if (!__any_sync(0xffffffffu, kSeam || mx >= -kEdgeEps)) continue;
#pragma unroll
        for (int u = 0; u < kTileU; ++u) {
            const bool sl = kSeam && seam[u];
            if (c0.m[u] >= -kEdgeEps || sl)
                tile_candidate(rp, sm, exacts, c0.a, c0.b, c0.d, c0.eo, pitch[u], c0.m[u], sl, ks[u], j0 + u * 32 + lane, jbase, w[u]);
        }
        __syncwarp();
    }
}

#ifndef ISY_TILE_TAKE
#define ISY_TILE_TAKE 4
#endif
__device__ void raster_tiles(const RenderParams& rp, SmemLayout& sm, const Exact* exacts, const unsigned short* tlist, int Hcur,
                             int* s_next, int Rrow, unsigned jbase, int stop_after = 0x7fffffff) {
    const int lane = threadIdx.x & 31;
    const int NQ = (Rrow + kTileRays - 1) >> kTileShift, NT = NQ * Hcur;
    const int q_h = min(NQ - 1, NQ / 2);
    // the places of the upper half of the eye (all the work, dealt heaviest first) are taken one at a time, the rest --
    // tiles below the horizon, mostly empty -- ISY_TILE_TAKE at a time.  (4 / 6 / 8 / 12 / 16 single-fetch rows of the
    // grid eye's 16: 19.80 / 19.57 / 19.55 / 19.62 / 19.68 ms; none: 24.9 ms)
    const int heavy = ((NQ + 1) / 2) * Hcur;
    const float rH = 1.0f / (float)Hcur;
    const float qnan = __int_as_float(0x7fc00000);
    int take = 1;
    for (;;) {
        const int t0 = warp_fetch(s_next, take);
        if (t0 >= NT) break;
        const int t1 = min(t0 + take, NT);
        if (t0 >= heavy) take = ISY_TILE_TAKE;
        for (int t = t0; t < t1; ++t) {
            const unsigned i1 = sm.scan.tile_cnt[t], i0 = t ? sm.scan.tile_cnt[t - 1] : 0u;
            if (i0 == i1) continue;
#ifdef ISY_DEBUG_BOUNDS
            if (i1 < i0 || i1 > (unsigned)(kMaxE * 24)) { ISY_BOUND(false); continue; }   // (keep the debug build alive)
#endif
            // place t of the queue -> tile (q, c)
            const int tq = (int)(((float)t + 0.5f) * rH), c = t - tq * Hcur, q = chunk_place(tq, q_h);   // (its own inverse)
            ISY_BOUND(q >= 0 && q < NQ && c >= 0 && c < Hcur && i1 <= (unsigned)(kMaxE * 24));
            const float off = sm.scan.coloff[c + 1];
            const bool edge = c == 0 || c == Hcur - 1;
            const unsigned j0 = (unsigned)q * kTileRays;
            float pitch[kTileU], f[kTileU];
            bool seam[kTileU], any_seam = false;
            int ks[kTileU];
            TileBest w[kTileU];
#pragma unroll
            for (int u = 0; u < kTileU; ++u) {
                const unsigned j = j0 + u * 32 + lane;
                const bool ok = j < (unsigned)Rrow;
                pitch[u] = ok ? sm.ras.sorted[j].x : qnan;      // (a NaN pitch fails every test)
                f[u] = (ok ? sm.scan.phase[j] : 0.f) + off;
                seam[u] = ok && edge && fabsf(f[u]) > 3.14159f;
                any_seam |= seam[u];
                ks[u] = c - (ok ? (int)sm.scan.colb[j] : 0);      // sorted slot of the heading that puts ray j into column c
                if (ks[u] < 0) ks[u] += Hcur;
                w[u].off = 0xffffffffu, w[u].lo = 0u, w[u].hi = 0x7fffffff, w[u].tri = 0;
            }
            if (edge && __any_sync(0xffffffffu, any_seam)) tile_list<true>(rp, sm, exacts, tlist + i0, tlist + i1, pitch, f, seam, ks, j0, jbase, w);
            else tile_list<false>(rp, sm, exacts, tlist + i0, tlist + i1, pitch, f, seam, ks, j0, jbase, w);
#pragma unroll
            for (int u = 0; u < kTileU; ++u)
                if (w[u].off != 0xffffffffu) {
                    const unsigned j = j0 + u * 32 + lane;
                    ISY_BOUND((unsigned)ks[u] * Rrow + j < (unsigned)kBestCap);
                    sm.ras.best[(unsigned)ks[u] * Rrow + j] = (unsigned short)(w[u].off / (unsigned)sizeof(Entry));
                }
        }
        if (t1 > stop_after) break;
    }
}

// The pitch-run walk of one group of headings over one band of rays (a small eye is one band): the band's sorted
// ray table and the winner buffer (u16 slots, Rrow per heading) live in shared memory.
// The hot path is fp32: query azimuth = fl(yaw_f + heading_f) wrapped in fp32; kEdgeEps covers that
// rounding.  Rays that land within 3e-6 rad of the +-pi seam, and rays inside the rounding band of
// an edge, redo the query in fp64 exactly as the reference composes it.
template <bool kUniform>
__device__ void raster_copies(const RenderParams& rp, SmemLayout& sm, const Exact* exacts, int E, int Hcur, int* s_next,
                              int Rrow, unsigned jbase, int stop_after = 0x7fffffff) {
    // stop_after: the warp leaves once it has taken a copy at or beyond this place in the queue (the copy warps of
    // rows mode step out in mid-rasterisation to write the template rows, then come back)
    const int lane = threadIdx.x & 31;
    const float pi_f = (float)kPi;
    const int ntall = sm.order_n[0];
    bool last = false;
    for (;;) {
        if (last) break;
        // copies differ a lot in size (a near blade spans hundreds of rays): warps pull them from a counter
        int e = 0;
        if (lane == 0) e = smem_atomic_add(s_next, 1);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e >= E) break;
        last = e >= stop_after;
        ISY_BOUND(e < kMaxE && ntall <= E);
        e = sm.order[e < ntall ? e : kMaxE - 1 - (e - ntall)];   // tall copies first: no long straggler at the end
        ISY_BOUND(e < E);
        // run of pitch-sorted rays that covers the copy's pitch range (a few extra rays at both ends),
        // worked out when the copy was projected (small eyes) or when the band was loaded
        const unsigned int jr = sm.jrange[e];
        const unsigned jbeg = jr & 0xffffu, jend = jr >> 16;
        if (jbeg >= jend) continue;                // no ray of this band in the copy's pitch range
        const float4 bb = sm.bbox[e];
        // azimuth window of the copy clipped to the query domain (-pi, pi]
        const float glo = fmaxf(bb.z, -pi_f - kBinMargin) - kBinMargin, ghi = fminf(bb.w, pi_f + kBinMargin) + kBinMargin;
        if (glo > ghi) continue;
        const float width = ghi - glo;
        const Entry ent = sm.tile[e];
        unsigned jb = jbeg;
        if constexpr (kUniform) {   // column walk
            const float inv_delta = sm.ras.uni[2];
            const int c_lo = (int)floorf((glo + pi_f) * inv_delta), c_hi = (int)floorf((ghi + pi_f) * inv_delta);
            ISY_BOUND(c_lo >= -1 && c_hi <= Hcur);
            for (; jb + 32 * (kRU - 1) < jend; jb += 32 * kRU) {
                __syncwarp();
                column_chunk<kRU>(rp, sm, exacts, ent, e, jb, jend, c_lo, c_hi, Hcur, Rrow, jbase);
            }
            for (; jb < jend; jb += 32) {
                __syncwarp();
                column_chunk<1>(rp, sm, exacts, ent, e, jb, jend, c_lo, c_hi, Hcur, Rrow, jbase);
            }
        } else {                     // pitch-run walk: kRU rays per lane while that many warps' worth are left, then one
            const double dist = __hiloint2double(ent.dist_hi, ent.dist_lo);
            for (; jb + 32 * (kRU - 1) < jend; jb += 32 * kRU) {   // warp-uniform trip counts
                __syncwarp();                                       // keep the lanes converged across iterations
                walk_chunk<kRU>(rp, sm, exacts, ent, bb, dist, e, jb, jend, glo, width, Hcur, Rrow, jbase);
            }
            for (; jb < jend; jb += 32) {
                __syncwarp();
                walk_chunk<1>(rp, sm, exacts, ent, bb, dist, e, jb, jend, glo, width, Hcur, Rrow, jbase);
            }
        }
    }
}

// pitch runs of all copies in the band whose look-up table is resident (large eyes: once per band)
__device__ __forceinline__ void band_runs(SmemLayout& sm, int E) {
    for (int e = threadIdx.x; e < E; e += kThreads) {
        const float4 bb = sm.bbox[e];
        const int s0 = min(max((int)floorf((bb.x + (float)kHalfPi) * (kPitchLut / (float)kPi)), 0), kPitchLut - 1);
        const int s1 = min(max((int)floorf((bb.y + (float)kHalfPi) * (kPitchLut / (float)kPi)) + 2, 0), kPitchLut);
        sm.jrange[e] = sm.ras.plut[s0] | (sm.ras.plut[s1] << 16);
    }
}

// ------------------------------------------------------------------------------------------
// Few headings per position (single views, route batches): walking every ray of a copy's pitch band
// would test ~98 % of the pairs for nothing, so each (copy, heading) pair instead visits the few
// cells of the eye's static (pitch, yaw) ray grid under the copy's bounding box, shifted by the
// heading.  Lanes hold different pairs; the work per position is a few thousand warp instructions.
// ------------------------------------------------------------------------------------------
// tests rays [i0, i1) of the grid's ray list (every `step`-th from i0 + first) against one projected copy at one heading
template <bool kSmem>
__device__ __forceinline__ void raster_list(const RenderParams& rp, SmemLayout& sm, unsigned int* gbest,
                                            const Exact* exacts, unsigned i0, unsigned i1, unsigned first, unsigned step,
                                            const Entry& ent, double dist, int e, int hh, int hs, float hf) {
    const int R = rp.R;
    const float2* gsorted = reinterpret_cast<const float2*>(rp.grid_sorted);
    const float pi_f = (float)kPi, twopi_f = (float)kTwoPi;
    const unsigned int best_s = (unsigned int)__cvta_generic_to_shared(sm.ras.best);
    for (unsigned i = i0 + first; i < i1; i += step) {
        ISY_BOUND(i < (unsigned)R && (!kSmem || i < (unsigned)kGridRaysSmem));
        const unsigned j = kSmem ? (unsigned)sm.ras.cell_list[i] : __ldg(rp.cells_list + i);
        ISY_BOUND(j < (unsigned)R);
        const float2 py = kSmem ? sm.ras.sorted[j] : __ldg(gsorted + j);
        float fy = py.y + hf;                                  // both in (-pi, pi]: one shift wraps
        fy = fy > pi_f ? fy - twopi_f : (fy <= -pi_f ? fy + twopi_f : fy);
        const bool seam = fabsf(fy) > 3.14159f;
        float m = fminf(fmaf(ent.A0, py.x, fmaf(ent.B0, fy, ent.C0)),
                        fminf(fmaf(ent.A1, py.x, fmaf(ent.B1, fy, ent.C1)), fmaf(ent.A2, py.x, fmaf(ent.B2, fy, ent.C2))));
        if (!seam && m < -kEdgeEps) continue;
        bool in = m > kEdgeEps && !seam;
        if (!in) {   // redo the query in fp64 exactly as the reference composes it
            const int ray = __ldg(rp.grid_sorted_ray + j);
            double yg = rp.eye_py[2 * (size_t)ray + 1] + sm.head_val[hh];
            if (yg > kPi) yg -= kTwoPi;
            else if (yg <= -kPi) yg += kTwoPi;
            const float f = (float)yg;
            m = fminf(fmaf(ent.A0, py.x, fmaf(ent.B0, f, ent.C0)),
                      fminf(fmaf(ent.A1, py.x, fmaf(ent.B1, f, ent.C1)), fmaf(ent.A2, py.x, fmaf(ent.B2, f, ent.C2))));
            in = m > kEdgeEps;
            if (!in && m >= -kEdgeEps) in = inside_exact(exacts[e], rp.eye_py[2 * (size_t)ray], yg);
        }
        if (in) {
            if (kSmem) post_copy_u16(best_s, (unsigned)hs * R + j, sm.tile, dist, ent.tri, e);
            else post_copy_u32(&gbest[(size_t)hs * R + j], sm.tile, dist, ent.tri, e);
        }
    }
}

// the cells of one grid row under a pair's window are ncols consecutive columns from c0 (modulo the circle):
// one or two contiguous ranges of the grid's ray list
template <bool kSmem>
__device__ __forceinline__ void raster_row(const RenderParams& rp, SmemLayout& sm, unsigned int* gbest,
                                           const Exact* exacts, int row, int c0, int ncols, unsigned first, unsigned step,
                                           const Entry& ent, double dist, int e, int hh, int hs, float hf) {
    const int Gy = rp.cells_cols;
    const int n1 = min(ncols, Gy - c0);            // columns before the wrap
    const int cellA = row * Gy + c0;
    const unsigned a0 = kSmem ? sm.ras.cell_start[cellA] : __ldg(rp.cells_start + cellA);
    const unsigned a1 = kSmem ? sm.ras.cell_start[cellA + n1] : __ldg(rp.cells_start + cellA + n1);
    raster_list<kSmem>(rp, sm, gbest, exacts, a0, a1, first, step, ent, dist, e, hh, hs, hf);
    if (n1 < ncols) {                              // the window wraps: columns 0 .. ncols - n1 - 1 of the same row
        const int cellB = row * Gy;
        const unsigned b0 = kSmem ? sm.ras.cell_start[cellB] : __ldg(rp.cells_start + cellB);
        const unsigned b1 = kSmem ? sm.ras.cell_start[cellB + ncols - n1] : __ldg(rp.cells_start + cellB + ncols - n1);
        raster_list<kSmem>(rp, sm, gbest, exacts, b0, b1, first, step, ent, dist, e, hh, hs, hf);
    }
}

// The same for a whole warp, all rows of the window at once: lanes first fetch the list ranges of 32 (row, half)
// units, a prefix sum lines their rays up, and every lane then takes every 32nd ray of the concatenation -- the
// lanes stay busy whether a row holds one ray (small eyes) or hundreds (dense eyes).  Units are taken from
// ub0 in steps of ustep (one warp: 0, 32; all warps of the CTA on one pair: 32 * warp, 32 * kWarps).
template <bool kSmem>
__device__ __forceinline__ void raster_rows_flat(const RenderParams& rp, SmemLayout& sm, unsigned int* gbest,
                                                 const Exact* exacts, int r0, int nrows, int c0, int ncols, int ub0, int ustep,
                                                 const Entry& ent, double dist, int e, int hh, int hs, float hf, int seg = 0,
                                                 int nseg = 1) {
    // seg / nseg: this caller takes the seg-th of nseg equal parts of every unit's list range (a pair shared by the
    // whole CTA: every warp gets a slice of every row, whatever the number of rows)
    const int lane = threadIdx.x & 31, Gy = rp.cells_cols;
    const int n1 = min(ncols, Gy - c0);            // columns before the wrap
    const bool wraps = n1 < ncols;
    const int nunits = wraps ? 2 * nrows : nrows;
    for (int ub = ub0; ub < nunits; ub += ustep) {
        const int u = ub + lane;
        unsigned a0 = 0, a1 = 0;
        if (u < nunits) {
            const int row = r0 + (wraps ? (u >> 1) : u);
            const bool second = wraps && (u & 1);
            const int cell = row * Gy + (second ? 0 : c0), n = second ? ncols - n1 : n1;
            a0 = kSmem ? sm.ras.cell_start[cell] : __ldg(rp.cells_start + cell);
            a1 = kSmem ? sm.ras.cell_start[cell + n] : __ldg(rp.cells_start + cell + n);
            if (nseg > 1) {
                const unsigned full = a1 - a0;
                a1 = a0 + (unsigned)(((unsigned long long)full * (unsigned)(seg + 1)) / (unsigned)nseg);
                a0 = a0 + (unsigned)(((unsigned long long)full * (unsigned)seg) / (unsigned)nseg);
            }
        }
        const unsigned len = a1 - a0;
        unsigned inc = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += v;
        }
        const unsigned total = __shfl_sync(0xffffffffu, inc, 31), exc = inc - len;
        for (unsigned tb = 0; tb < total; tb += 32) {          // warp-uniform trip count
            const unsigned t = tb + lane;
            int k = 0;                                         // the unit that holds ray t: smallest k with inc_k > t
#pragma unroll
            for (int b = 16; b > 0; b >>= 1) {
                const unsigned v = __shfl_sync(0xffffffffu, inc, k + b - 1);
                if (v <= t) k += b;
            }
            k = min(k, 31);
            const unsigned ek = __shfl_sync(0xffffffffu, exc, k), ak = __shfl_sync(0xffffffffu, a0, k);
            if (t < total) {
                const unsigned i = ak + (t - ek);
                raster_list<kSmem>(rp, sm, gbest, exacts, i, i + 1, 0, 1, ent, dist, e, hh, hs, hf);
            }
        }
    }
}

// window of pair (copy e, heading hh) in the eye's static ray grid: rows [r0, r0 + nrows), columns ncols from c0
__device__ __forceinline__ void pair_window(const RenderParams& rp, const SmemLayout& sm, int e, int hh, int& r0, int& nrows,
                                            int& c0, int& ncols) {
    const int Gp = rp.cells_rows, Gy = rp.cells_cols;
    const float pi_f = (float)kPi;
    const float inv_dp = (float)Gp / pi_f, inv_dy = (float)Gy / (float)kTwoPi;
    r0 = 0, nrows = 0, c0 = 0, ncols = 0;
    const float4 bb = sm.bbox[e];
    const float glo = fmaxf(bb.z, -pi_f - kBinMargin) - kBinMargin, ghi = fminf(bb.w, pi_f + kBinMargin) + kBinMargin;
    if (glo > ghi) return;
    const float hf = (float)sm.head_val[hh];
    // eye-frame azimuth window [glo - h, ghi - h], as a run of grid columns modulo the circle
    const int c_first = (int)floorf((glo - hf + pi_f) * inv_dy);
    ncols = min((int)floorf((ghi - hf + pi_f) * inv_dy) - c_first + 1, Gy);
    c0 = c_first % Gy;
    if (c0 < 0) c0 += Gy;
    r0 = max((int)floorf((bb.x + (float)kHalfPi) * inv_dp), 0);
    nrows = max(min((int)floorf((bb.y + (float)kHalfPi) * inv_dp), Gp - 1) - r0 + 1, 0);
}

// Work sharing: pairs covering a few cells take one lane each; larger ones are walked by their warp, row by row
// (the rays of a row's window are contiguous in the grid's list); the few giants (a blade right in front of the
// eye, or a degenerate copy that may contain any ray) are set aside and walked by the whole CTA afterwards, so
// that no warp is left alone with them.
template <bool kSmem>
__device__ void raster_cells(const RenderParams& rp, SmemLayout& sm, unsigned int* gbest, const Exact* exacts, int E,
                             int Hcur, int* s_next) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int items = E * Hcur;
    constexpr int kBigItem = 12;      // cells: above this the warp shares the pair
    constexpr int kGiantItem = 1024;  // cells: above this the CTA shares the pair
    // Large eyes (tables in global memory, hundreds of rays under an ordinary copy): nearly every pair is "large", and a
    // warp that walked the 32 it classified one after the other would leave the others idle.  They are queued instead
    // (sm.jrange is free outside the band walk) and pulled one at a time by whichever warp is free.
    constexpr bool kQueue = !kSmem;
    unsigned short* giants = sm.giants;
    int* n_giant = &sm.n_giants;
    int* n_queued = &sm.order_n[1];
    unsigned int* queue = sm.jrange;
    if (threadIdx.x == 0) *n_giant = 0, *n_queued = 0;
    __syncthreads();
    for (;;) {
        int base = 0;
        if (lane == 0) base = smem_atomic_add(s_next, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= items) break;
        const int item = base + lane;
        int e = 0, hh = 0, r0 = 0, nrows = 0, c0 = 0, ncols = 0;
        if (item < items) {
            e = item / Hcur, hh = item - e * Hcur;
            pair_window(rp, sm, e, hh, r0, nrows, c0, ncols);
        }
        int ncell = nrows * ncols;
        if (ncell > kGiantItem && item < 65536) {      // set aside (if the list is full the warp keeps it)
            const int slot = smem_atomic_add(n_giant, 1);
            if (slot < kMaxGiants) giants[slot] = (unsigned short)item, ncell = 0;
        }
        if (kQueue && ncell > kBigItem) {
            const int slot = smem_atomic_add(n_queued, 1);
            if (slot < kMaxE) queue[slot] = (unsigned int)item, ncell = 0;
        }
        // small pairs: one lane each
        if (ncell > 0 && ncell <= kBigItem) {
            const Entry ent = sm.tile[e];
            const double dist = __hiloint2double(ent.dist_hi, ent.dist_lo);
            const float hf = (float)sm.head_val[hh];
            const int hs = sm.ras.hinv[hh];
            for (int row = r0; row < r0 + nrows; ++row)
                raster_row<kSmem>(rp, sm, gbest, exacts, row, c0, ncols, 0, 1, ent, dist, e, hh, hs, hf);
        }
        // large pairs: the warp shares their rows
        unsigned int big = __ballot_sync(0xffffffffu, ncell > kBigItem);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const int be = __shfl_sync(0xffffffffu, e, src), bh = __shfl_sync(0xffffffffu, hh, src);
            const int br0 = __shfl_sync(0xffffffffu, r0, src), bnr = __shfl_sync(0xffffffffu, nrows, src);
            const int bc0 = __shfl_sync(0xffffffffu, c0, src), bnc = __shfl_sync(0xffffffffu, ncols, src);
            const Entry ent = sm.tile[be];
            const double dist = __hiloint2double(ent.dist_hi, ent.dist_lo);
            const float hf = (float)sm.head_val[bh];
            const int hs = sm.ras.hinv[bh];
            raster_rows_flat<kSmem>(rp, sm, gbest, exacts, br0, bnr, bc0, bnc, 0, 32, ent, dist, be, bh, hs, hf);
            __syncwarp();
        }
        __syncwarp();
    }
    __syncthreads();
    if (kQueue) {   // the queued pairs, one at a time per warp
        if (threadIdx.x == 0) *s_next = 0;
        __syncthreads();
        const int nq = min(*n_queued, kMaxE);
        for (;;) {
            int i = 0;
            if (lane == 0) i = smem_atomic_add(s_next, 1);
            i = __shfl_sync(0xffffffffu, i, 0);
            if (i >= nq) break;
            const int item = (int)queue[i];
            const int e = item / Hcur, hh = item - e * Hcur;
            int r0, nrows, c0, ncols;
            pair_window(rp, sm, e, hh, r0, nrows, c0, ncols);
            const Entry ent = sm.tile[e];
            const double dist = __hiloint2double(ent.dist_hi, ent.dist_lo);
            raster_rows_flat<kSmem>(rp, sm, gbest, exacts, r0, nrows, c0, ncols, 0, 32, ent, dist, e, hh, sm.ras.hinv[hh],
                                    (float)sm.head_val[hh]);
            __syncwarp();
        }
    }
    // giants: every warp takes its slice of every row
    const int ng = min(*n_giant, kMaxGiants);
    for (int g = 0; g < ng; ++g) {
        const int item = giants[g];
        const int e = item / Hcur, hh = item - e * Hcur;
        int r0, nrows, c0, ncols;
        pair_window(rp, sm, e, hh, r0, nrows, c0, ncols);
        const Entry ent = sm.tile[e];
        const double dist = __hiloint2double(ent.dist_hi, ent.dist_lo);
        const float hf = (float)sm.head_val[hh];
        const int hs = sm.ras.hinv[hh];
        raster_rows_flat<kSmem>(rp, sm, gbest, exacts, r0, nrows, c0, ncols, 0, 32, ent, dist, e, hh, hs, hf, warp, kWarps);
    }
}

}  // namespace isy

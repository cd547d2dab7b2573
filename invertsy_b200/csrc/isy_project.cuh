// Phase A: cull (world cell grid) and angular projection of one position.
#pragma once
#include "isy_layout.cuh"

namespace isy {

// ------------------------------------------------------------------------------------------
// phase A: cull + project one position. Entries go to the shared-memory tile (first kMaxE) and
// beyond that to the CTA's slab; the exact fp64 records always go to the slab. Returns the number of copies E.
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

// cell index of a coordinate in the world's xy grid -- the host files vertices with the same float ops
__host__ __device__ __forceinline__ int world_cell(float x, float x0, float inv_cell, int n) {
    const float q = (x - x0) * inv_cell;
    int c = (int)floorf(q);
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

__device__ __forceinline__ int float_key(float f) {   // order-preserving int key for atomicMin/Max
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// kRuns: the eye's pitch look-up table is resident in shared memory (small eyes, yaw-only views): the copy's run
// of pitch-sorted rays is worked out here, once per copy by one thread, instead of by every warp that rasterises it
template <bool kRuns>
__device__ __forceinline__ void emit_copy(const double th[3], const double ph[3], double dist, int t, int slot,
                                          Entry* entries, Exact* exacts, SmemLayout& sm, int* s_band) {
    Entry e;
    finish_copy(th, ph, dist, t, e, exacts[slot]);
    if (slot >= kMaxE) entries[slot] = e;   // the slab only takes what the shared-memory tile cannot hold
    ISY_BOUND(slot >= 0);
    if (slot < kMaxE) {
        sm.tile[slot] = e;
        float tlo = (float)fmin(th[0], fmin(th[1], th[2])) - kBinMargin;
        float thi = (float)fmax(th[0], fmax(th[1], th[2])) + kBinMargin;
        float plo = (float)fmin(ph[0], fmin(ph[1], ph[2])) - kBinMargin;
        float phi = (float)fmax(ph[0], fmax(ph[1], ph[2])) + kBinMargin;
        // The three reference crosses of a proper triangle have signs (+,-,+) or (-,+,-) and the same-side
        // region is the triangle itself.  If any of them is 0 (vertices collinear or coincident in angle
        // space) that test passes for EVERY direction (world.py:526), and if rounding breaks the pattern the
        // region is an unbounded wedge: such copies may contain rays anywhere, so their box is the whole domain.
        const Exact& x = exacts[slot];
        const bool proper = (x.o[0] > 0 && x.o[1] < 0 && x.o[2] > 0) || (x.o[0] < 0 && x.o[1] > 0 && x.o[2] < 0);
        if (!proper) tlo = -4.f, thi = 4.f, plo = -8.f, phi = 8.f;
        sm.bbox[slot] = make_float4(tlo, thi, plo, phi);
        if (kRuns) {
            // run of pitch-sorted rays that covers [tlo, thi] (a few extra rays at both ends)
            const int s0 = min(max((int)floorf((tlo + (float)kHalfPi) * (kPitchLut / (float)kPi)), 0), kPitchLut - 1);
            const int s1 = min(max((int)floorf((thi + (float)kHalfPi) * (kPitchLut / (float)kPi)) + 2, 0), kPitchLut);
            sm.jrange[slot] = sm.ras.plut[s0] | (sm.ras.plut[s1] << 16);
        }
        // longest-first order for the rasteriser's dynamic queue: a copy's cost grows with its pitch run
        if (thi - tlo > kTallCopy) sm.order[atomicAdd(&sm.order_n[0], 1)] = (unsigned short)slot;
        else sm.order[kMaxE - 1 - atomicAdd(&sm.order_n[1], 1)] = (unsigned short)slot;
        atomicMin(&s_band[0], float_key(tlo));
        atomicMax(&s_band[1], float_key(thi));
    }
}

// scratch of one cull in shared memory
struct CullScratch {
    int count;                         // visible triangles so far
    unsigned int rowbeg[16], rowoff[17];
};


// A1: the visible triangles of position p -> vis[] (triangle ids) and stage[] (their vertices, 9 planes of stage_cap
// floats; shared or global memory), by a team of threads (gt = this thread's index in it; the team starts at
// a warp boundary).  Returns their number.  The CTA runs it for the position it is about to render; in rows mode the
// copy warps also run it for the NEXT position while the others rasterise (vis / stage in the workspace), so that
// the latency of its three dependent trips to L2 (cell rows -> cell lists -> vertices) is off the critical path.
__device__ int cull_position(const RenderParams& rp, const Team& tm, int p, int gt, int* vis, float* stage, int stage_cap, CullScratch& sc) {
    const int kNT = tm.n;
    const int lane = gt & 31;
    const double px = rp.pos[3 * p + 0], py = rp.pos[3 * p + 1], pz = rp.pos[3 * p + 2];
    const double h = rp.horizon;
    const double h2_max = rp.horizon_sq_max;     // sqrt_rn(s) <= horizon  <=>  s <= h2_max (set by the host)
    if (gt == 0) sc.count = 0;
    tm.sync();
    // ---- A1: cull (world.py:97-98) ----
    // Candidates come from the world's xy cell grid (isy_world_create): a visible triangle has a vertex
    // within the horizon, so it is listed in a cell that meets the disc around the eye.  A triangle is
    // listed once per distinct vertex cell (tagged with that vertex); it is taken from the list entry
    // whose cell holds its NEAREST vertex, i.e. exactly once.  The test itself is the reference's, fp64.
    const float gx0 = rp.wg_x0, gy0 = rp.wg_y0, ginv = rp.wg_inv_cell;
    const int nx = rp.wg_nx, ny = rp.wg_ny;
    const double slack = 1e-2;   // cells; covers the float rounding of the vertex filing (< 1e-3 cells for 4096 cells a side)
    const int cx0 = (int)fmax(fmin(floor((px - h - gx0) * ginv - slack), (double)(nx - 1)), 0.0);
    const int cx1 = (int)fmax(fmin(floor((px + h - gx0) * ginv + slack), (double)(nx - 1)), 0.0);
    const int cy0 = (int)fmax(fmin(floor((py - h - gy0) * ginv - slack), (double)(ny - 1)), 0.0);
    const int cy1 = (int)fmax(fmin(floor((py + h - gy0) * ginv + slack), (double)(ny - 1)), 0.0);
    // the candidate cells form one contiguous list range per grid row: flatten the (<= 16) rows.
    // One lane per row: the row keeps only the columns that meet the disc of radius h around the eye
    // (a visible triangle's nearest vertex is within h in xy), then a warp scan gives the offsets.
    const int nrows = min(cy1 - cy0 + 1, 16);   // cell = horizon / 4 or larger: at most 10 rows
    if (gt < 32) {
        unsigned int beg = 0, cnt = 0;
        if (gt < nrows) {
            const int row = cy0 + gt;
            int rx0 = cx0, rx1 = cx1;
            if (row > 0 && row < ny - 1) {   // (border rows also hold whatever lies beyond the grid: keep the full span)
                const double cell = 1.0 / (double)ginv;
                const double ylo = (double)gy0 + row * cell, yhi = ylo + cell;
                const double dy = fmax(fmax(ylo - py, py - yhi) - slack * cell, 0.0);
                const double half = sqrt(fmax(h * h - dy * dy, 0.0)) + 1e-9;
                rx0 = (int)fmax(fmin(floor((px - half - gx0) * ginv - slack), (double)(nx - 1)), 0.0);
                rx1 = (int)fmax(fmin(floor((px + half - gx0) * ginv + slack), (double)(nx - 1)), 0.0);
            }
            beg = __ldg(rp.wg_cell_start + (size_t)row * nx + rx0);
            cnt = __ldg(rp.wg_cell_start + (size_t)row * nx + rx1 + 1) - beg;
        }
        unsigned int incl = cnt;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        if (gt < nrows) sc.rowbeg[gt] = beg, sc.rowoff[gt] = incl - cnt;
        if (gt == nrows - 1) sc.rowoff[nrows] = incl;
    }
    tm.sync();
    const unsigned int ncand = sc.rowoff[nrows];
    // two candidates per thread and trip: their list reads and vertex gathers (two dependent trips to L2
    // each) overlap
    for (unsigned int base = 0; base < ncand; base += 2 * kNT) {   // block-uniform trip count
        unsigned int packed[2];
        bool have[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const unsigned int c = base + u * kNT + gt;
            have[u] = c < ncand;
            packed[u] = 0;
            if (have[u]) {
                int rr = 0;
                while (rr + 1 < nrows && c >= sc.rowoff[rr + 1]) ++rr;
                packed[u] = __ldg(rp.wg_cell_tris + sc.rowbeg[rr] + (c - sc.rowoff[rr]));
            }
        }
        double s3[2][3];
        float vx[2][3], vy[2][3], vzz[2][3];   // raw vertices: the cell test below files xy like the host did
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int t = (int)(packed[u] & 0x3fffffffu);
            // candidates are gathered by index: three 16-byte loads from the triangle-major copy of the world
            const float4 a = __ldg(rp.tri_aos + 3 * (size_t)t), b = __ldg(rp.tri_aos + 3 * (size_t)t + 1),
                         c = __ldg(rp.tri_aos + 3 * (size_t)t + 2);
            vx[u][0] = a.x, vy[u][0] = a.y, vzz[u][0] = a.z;
            vx[u][1] = a.w, vy[u][1] = b.x, vzz[u][1] = b.y;
            vx[u][2] = b.z, vy[u][2] = b.w, vzz[u][2] = c.x;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                s3[u][k] = norm2_rn((double)vx[u][k] - px, (double)vy[u][k] - py, (double)vzz[u][k] - pz);   // world.py:94-96
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int t = (int)(packed[u] & 0x3fffffffu), tag = (int)(packed[u] >> 30);
            const double s0 = s3[u][0], s1 = s3[u][1], s2 = s3[u][2];
            const double smin = fmin(s0, fmin(s1, s2));
            bool visible = false;
            if (have[u] && smin <= h2_max) {                                      // world.py:97
                const int kn = (s0 <= s1 && s0 <= s2) ? 0 : (s1 <= s2 ? 1 : 2);   // nearest vertex
                // same float arithmetic as the host used to file the vertices into cells
                const float xn = kn == 0 ? vx[u][0] : (kn == 1 ? vx[u][1] : vx[u][2]);
                const float yn = kn == 0 ? vy[u][0] : (kn == 1 ? vy[u][1] : vy[u][2]);
                const float xt = tag == 0 ? vx[u][0] : (tag == 1 ? vx[u][1] : vx[u][2]);
                const float yt = tag == 0 ? vy[u][0] : (tag == 1 ? vy[u][1] : vy[u][2]);
                const int cnx = world_cell(xn, gx0, ginv, nx), cny = world_cell(yn, gy0, ginv, ny);
                const int ctx = world_cell(xt, gx0, ginv, nx), cty = world_cell(yt, gy0, ginv, ny);
                visible = cnx == ctx && cny == cty;
            }
            unsigned m = __ballot_sync(0xffffffffu, visible);
            if (m) {
                int basei = 0;
                if (lane == 0) basei = atomicAdd(&sc.count, __popc(m));
                basei = __shfl_sync(0xffffffffu, basei, 0);
                if (visible) {
                    int slot = basei + __popc(m & ((1u << lane) - 1));
                    if (slot < rp.slab_cap) vis[slot] = t;
                    if (slot < stage_cap) {   // the projection takes the vertices from here instead of gathering them again
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            stage[(3 * k + 0) * stage_cap + slot] = vx[u][k];
                            stage[(3 * k + 1) * stage_cap + slot] = vy[u][k];
                            stage[(3 * k + 2) * stage_cap + slot] = vzz[u][k];
                        }
                    }
                }
            }
        }
    }
    tm.sync();
    tm.sync();
    return sc.count;
}

constexpr int kStage = 1792;   // visible triangles whose vertices are staged in shared memory (63 KB)
static_assert(9 * kStage * sizeof(float) <= sizeof(unsigned short) * kBestCap, "the stage must fit the winner buffer");
static_assert(9 * kStage * sizeof(float) <= sizeof(SmemBins), "... and the bin lists (built after the projection)");

template <bool kRuns>
__device__ int cull_and_project(const RenderParams& rp, const Team& tm, int p, int* vis, Entry* entries, Exact* exacts,
                                SmemLayout& sm, CullScratch* sc, const CullScratch* ahead, const float* ahead_stage,
                                int ahead_cap, int* s_nent, int* s_band, long long* t_mid = nullptr) {
    // ahead != null: position p was culled ahead (vis[] filled, vertices in ahead_stage, ahead->count of them)
    const int tid = threadIdx.x;
    if (rp.T == 0 || rp.horizon_sq_max < 0) {   // no world (sky-only calls) or a negative horizon (world.py:98): nothing to cull
        if (tid == 0) s_band[0] = float_key(CUDART_INF_F), s_band[1] = float_key(-CUDART_INF_F);
        return 0;
    }
    const double px = rp.pos[3 * p + 0], py = rp.pos[3 * p + 1], pz = rp.pos[3 * p + 2];
    // vertices of the visible triangles, cull -> projection: 9 planes of kStage floats over the winner buffer / bin
    // lists, which nothing uses before this position is rasterised
    const float* stage = reinterpret_cast<float*>(&sm.ras);
    int stage_cap = kStage;

    if (tid == 0) {
        *s_nent = 0;
        sm.order_n[0] = 0, sm.order_n[1] = 0;
        s_band[0] = float_key(CUDART_INF_F);
        s_band[1] = float_key(-CUDART_INF_F);
    }
    tm.sync();

    // ---- A1: cull (world.py:97-98), unless the copy warps did it while the previous position was rasterised ----
    int V;
    if (ahead) {
        V = ahead->count;
        stage = ahead_stage, stage_cap = ahead_cap;
    } else {
        V = cull_position(rp, tm, p, tid, vis, reinterpret_cast<float*>(&sm.ras), kStage, *sc);
    }
#ifdef ISY_PHASE_DETAIL
    *t_mid = clock64();   // (profiling build only) end of the cull
#endif
    if (V > rp.slab_cap) {
        if (tid == 0) atomicExch(rp.status, (int)ISY_ERR_OVERFLOW);
        V = rp.slab_cap;
    }

    // ---- A2: angular projection of the visible triangles (world.py:109-120) ----
    for (int j = tid; j < V; j += tm.n) {
        int t = vis[j];
        double v[3][3];
        if (j < stage_cap) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                v[k][0] = (double)stage[(3 * k + 0) * stage_cap + j] - px;   // world.py:94
                v[k][1] = (double)stage[(3 * k + 1) * stage_cap + j] - py;
                v[k][2] = (double)stage[(3 * k + 2) * stage_cap + j] - pz;
            }
        } else {
            load_vertices(rp.tri_planes, rp.T, t, px, py, pz, v);
        }
        float col[3] = {0.f, 0.f, 0.f};
        if (rp.row_bg) col[0] = __ldg(rp.tri_rgb + 3 * (size_t)t), col[1] = __ldg(rp.tri_rgb + 3 * (size_t)t + 1),
                       col[2] = __ldg(rp.tri_rgb + 3 * (size_t)t + 2);   // rows mode: final_pass_patch reads it from shared memory
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
        if (rp.row_bg) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (slot < kMaxE) sm.rgb[3 * slot + c] = col[c];
                if (seam && slot + 1 < kMaxE) sm.rgb[3 * (slot + 1) + c] = col[c];
            }
            if (rp.out.resp) {   // what a ray that hits this triangle answers (one channel in rows mode)
                float sr = col[0], sg = col[1], sb = col[2];
                sensor_finish(rp.sensor, sr, sg, sb);
                const float pn = sensor_pn(sr, sg, sb);
                if (slot < kMaxE) sm.resp1[slot] = pn;
                if (seam && slot + 1 < kMaxE) sm.resp1[slot + 1] = pn;
            }
        }
        if (slot + (seam ? 2 : 1) <= rp.slab_cap) {
            if (seam) {
                double p1[3], p2[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    p1[k] = ph[k] < 0 ? ph[k] + kTwoPi : ph[k];                           // :118
                    p2[k] = ph[k] > 0 ? ph[k] - kTwoPi : ph[k];                           // :119
                }
                emit_copy<kRuns>(th, p1, dist, t, slot, entries, exacts, sm, s_band);
                emit_copy<kRuns>(th, p2, dist, t, slot + 1, entries, exacts, sm, s_band);
            } else {
                emit_copy<kRuns>(th, ph, dist, t, slot, entries, exacts, sm, s_band);
            }
        }
    }
    tm.sync();
    int E = *s_nent;
    if (E > rp.slab_cap) {
        if (tid == 0) atomicExch(rp.status, (int)ISY_ERR_OVERFLOW);
        E = rp.slab_cap;
    }
    return E;
}

}  // namespace isy

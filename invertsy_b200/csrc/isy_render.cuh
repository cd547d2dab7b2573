// The persistent render kernel.
#pragma once
#include "isy_shade.cuh"

// The copy warps can also cull the NEXT position while the others rasterise (cull_position on a named barrier).  It
// shortens the projection phase (22 k -> 16 k cycles per position) but keeps four warps out of the issue-bound
// rasteriser for longer (105 k -> 112 k): 22.1 -> 22.9 ms per step on the benchmark grid.  Off.
#ifndef ISY_CULL_AHEAD
#define ISY_CULL_AHEAD 0
#endif
#ifndef ISY_LATE_COPIES
#define ISY_LATE_COPIES 200
#endif
#ifndef ISY_COPY_WARPS
#define ISY_COPY_WARPS 4
#endif
// Rows mode, scans: the copy warps start on the position's rows the moment the position starts and leave the cull,
// the projection and the filing of the copies to the team of the other warps; they join the rasterisation when their
// rows are out.  (0: they take part in everything and copy while the others rasterise.)
#ifndef ISY_EARLY_ROWS
#define ISY_EARLY_ROWS 1
#endif

namespace isy {

// the template rows are written when this many copies are left to take: about what the other warps rasterise in the
// 25-30 k cycles the rows need on the store path (short positions start at once, long ones later than a fixed share)
constexpr int kLateCopies = ISY_LATE_COPIES;
#ifndef ISY_LATE_TILES
#define ISY_LATE_TILES 64
#endif
constexpr int kLateTiles = ISY_LATE_TILES;   // tile walk: ... when this many 256ths of the tile queue have been taken
constexpr int kCopyWarps = ISY_COPY_WARPS;   // warps that stream the template rows of a position while the rest rasterise

enum { kModeBrute = 0, kModeBins = 1, kModeRaster = 2 };

// kEye picks what is compiled in, so that each shape of call gets a kernel allocated and scheduled for it alone:
//   kEyeAny   any orientation (quaternion views), transient eyes without ray tables: per-position bins, or brute force
//   kEyeSmall yaw-only views of an eye whose tables fit shared memory: ray grid / pitch-run rasterisers
//   kEyeLarge yaw-only views of a larger eye: the same, band by band, or the ray grid with tables in HBM
enum { kEyeAny = 0, kEyeSmall = 1, kEyeLarge = 2 };

template <bool kF64, int kEye>
__global__ void __launch_bounds__(kThreads, 1) render_kernel(const RenderParams rp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
    __shared__ int s_count, s_nent, s_band[2];
    __shared__ unsigned int s_item, s_item_next, s_warp[kWarps];
    __shared__ CullScratch s_cull, s_cull_ahead;   // this position's cull; the next position's, run ahead by the copy warps
    __shared__ unsigned int s_ahead_item;          // item whose cull is in s_cull_ahead / the workspace stage

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* vis = rp.slab_vis + (size_t)blockIdx.x * rp.slab_cap;
    Entry* entries = rp.slab_entry + (size_t)blockIdx.x * rp.slab_cap;
    Exact* exacts = rp.slab_exact + (size_t)blockIdx.x * rp.slab_cap;
    unsigned int* gbest = rp.slab_best ? rp.slab_best + (size_t)blockIdx.x * rp.R : nullptr;
    const bool want_sky = rp.sky.kind != ISY_SKY_NONE;
    const bool perez = rp.sky.kind == ISY_SKY_PEREZ;
    const bool yaw_only = kEye != kEyeAny || !rp.vquat;   // (kEyeAny also serves yaw-only eyes without ray tables)
    const bool raster_ok = kEye != kEyeAny && !(rp.flags & ISY_FLAG_BRUTE_FORCE);
    const bool smem_grid = kEye == kEyeSmall && raster_ok;                // its tables stay in shared memory
    const bool banded = kEye == kEyeLarge && raster_ok && rp.Hg > kSmallH;   // scan plan: rendered in bands

    // row templates + patches (isy_shade.cuh): the host set the launch up for it; a Perez sky also needs every
    // position to look along the same headings, which heading_rows_kernel has just checked
    const bool rows_mode = kEye == kEyeSmall && !kF64 && rp.row_bg && raster_ok && (!perez || *rp.sky_tab_off == 0);
    if (perez) {
        for (int i = tid; i < 2 * (kGTab + 1); i += kThreads) sm.gtab[i] = rp.sky_gtab[i];
        if (tid == 0) sm.sun0 = rp.sun[0];
    }
    float* const ahead_stage = (ISY_CULL_AHEAD && rp.slab_stage) ? rp.slab_stage + (size_t)blockIdx.x * 9 * rp.stage_cap : nullptr;
    __shared__ volatile int s_join[3];             // early rows: (item, group) whose lists are ready + 1, how to walk them, E
    __shared__ int s_E;
    __shared__ volatile int s_ahead_go;
    if (tid == 0) {
        s_ahead_item = 0xffffffffu, s_join[0] = 0;
        if (kEye != kEyeAny && rp.Hg > kSmallH) sm.scan.ph_H = -1;    // (the launch allocated SmemScan, isy_capi.cu)
    }
    const bool early_rows = ISY_EARLY_ROWS && rows_mode && rp.H / rp.ngroups > kSmallH;
    const Team tm = early_rows ? Team{kThreads - 32 * kCopyWarps, 1} : Team{kThreads, 0};
    const bool copier = early_rows && tid >= tm.n;   // (never true outside rows mode)
    if (kEye == kEyeSmall && rp.row_bg)   // rows mode (the tail of the layout is allocated)
        for (int i = tid; i < rp.R; i += kThreads) sm.sorted_ray16[i] = (unsigned short)rp.grid_sorted_ray[i];
    if (smem_grid) {   // the eye's static tables: nothing else uses this memory in a yaw-only launch
        for (int i = tid; i <= kPitchLut; i += kThreads) sm.ras.plut[i] = (unsigned short)rp.grid_plut[i];
        for (int i = tid; i < rp.R; i += kThreads) {
            sm.ras.sorted[i] = reinterpret_cast<const float2*>(rp.grid_sorted)[i];
            sm.ras.cell_list[i] = (unsigned short)rp.cells_list[i];
        }
        for (int i = tid; i <= rp.cells_rows * rp.cells_cols; i += kThreads) sm.ras.cell_start[i] = (unsigned short)rp.cells_start[i];
    }
    long long t_phase[4] = {0, 0, 0, 0};   // project, accel (bins / raster), final ray pass, positions
#ifdef ISY_RASTER_DETAIL
    // profiling build: when the last rasterising warp / the last copy warp is done, and the whole phase (replace phases 0..2)
    __shared__ unsigned long long s_tend[3];
    if (tid == 0) s_tend[0] = 0, s_tend[1] = 0, s_tend[2] = 0;
    long long t_rd[3] = {0, 0, 0};
#endif
#ifdef ISY_SETUP_DETAIL
    long long t_sd[3] = {0, 0, 0};         // profiling build: clear + heading sort, phases, (phases +) filing of the copies
#endif
#ifdef ISY_PHASE_DETAIL
    long long t_detail[3] = {0, 0, 0};     // profiling build: cull, projection, raster set-up (replace phases 0..2)
#endif
    // work items come from a global counter; the next one is requested while the current one is rendered,
    // so the round trip of the atomic is never waited for
    unsigned int pulled = 0;
    if (tid == 0) pulled = atomicAdd(rp.queue, 1u);
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = pulled;
        __syncthreads();
        const unsigned item = s_item;
        if (item >= (unsigned)rp.items) break;
        if (tid == 0) pulled = atomicAdd(rp.queue, 1u);
        const int per_pos = rp.splits * rp.ray_splits;
        const int p = item / per_pos, split = (item % per_pos) / rp.ray_splits, rsplit = item % rp.ray_splits;
        // was this item culled while the previous one was rasterised?
        const CullScratch* const ahead = (ahead_stage && s_ahead_item == item) ? &s_cull_ahead : nullptr;
        long long t0 = clock64();

#ifdef ISY_PHASE_DETAIL
        long long t_sub[1] = {t0};   // end of the cull
        int E = copier ? 0 : smem_grid ? cull_and_project<true>(rp, tm, p, vis, entries, exacts, sm, &s_cull, ahead, ahead_stage, rp.stage_cap, &s_nent, s_band, t_sub)
                                : cull_and_project<false>(rp, tm, p, vis, entries, exacts, sm, &s_cull, ahead, ahead_stage, rp.stage_cap, &s_nent, s_band, t_sub);
        const long long g_detail_t = t_sub[0];
#else
        int E = copier ? 0 : smem_grid ? cull_and_project<true>(rp, tm, p, vis, entries, exacts, sm, &s_cull, ahead, ahead_stage, rp.stage_cap, &s_nent, s_band)
                                : cull_and_project<false>(rp, tm, p, vis, entries, exacts, sm, &s_cull, ahead, ahead_stage, rp.stage_cap, &s_nent, s_band);
#endif
        if (tid == 0) s_item_next = pulled, s_E = E;   // (the item after this one: requested when this one started)
        long long t1 = clock64();
#ifdef ISY_PHASE_DETAIL
        t_detail[0] += g_detail_t - t0, t_detail[1] += t1 - g_detail_t;
#endif
        int mode = kModeBrute;
        BinGrid grid;
        if (copier) {
            mode = kModeRaster;     // (corrected after the group's barrier if the tile overflowed)
        } else if (!(rp.flags & ISY_FLAG_BRUTE_FORCE) && E <= kMaxE) {
            if (kEye != kEyeAny) {
                mode = kModeRaster;
            } else if (build_bins(sm, grid, E, s_band, s_warp)) {
                mode = kModeBins;
            }
        }
        long long t2 = clock64();

        // headings are handled in groups of rp.Hg (the rasteriser's winner buffer holds one group)
        for (int grp = split; grp < rp.ngroups; grp += rp.splits) {
            // group grp = headings grp, grp + ngroups, ...: an evenly spaced scan stays evenly spaced
            const int Hcur = (rp.H - grp + rp.ngroups - 1) / rp.ngroups;
            long long t3 = clock64();
#ifdef ISY_PHASE_DETAIL
            long long t_walk = t3;
#endif
            if (early_rows && grp == split) {   // nothing of an earlier group to wait for: the copy warps go straight to their rows
                if (!copier) tm.sync();
            } else {
                __syncthreads();
            }
            if (yaw_only && tid < Hcur) {
                const double hw = wrap_yaw(rp.yaw ? rp.yaw[(size_t)p * rp.H + grp + (size_t)tid * rp.ngroups] : 0.0);
                sm.head_val[tid] = hw;
                sincos(hw, &sm.head_sin[tid], &sm.head_cos[tid]);
            }
            // A large eye under a scan plan is rendered band by band (kBandRays consecutive rays, sorted by pitch on
            // their own): each band goes through the shared-memory winner buffer like a small eye.  Everything
            // else is a single pass over all rays.
            bool is_raster = kEye != kEyeAny && mode == kModeRaster;
            const bool walk_bands = banded && is_raster;
            const int nbands = walk_bands ? (rp.R + kBandRays - 1) / kBandRays : 1;
            bool uniform = false;
            for (int band = 0; band < nbands; ++band) {
            const int r_lo = walk_bands ? band * kBandRays : 0;
            const int r_n = walk_bands ? min(kBandRays, rp.R - r_lo) : rp.R;
            const bool best_smem = kEye == kEyeSmall || walk_bands;
            long long t3b = clock64();
            if (is_raster) {
                if (band > 0) __syncthreads();     // the previous band's final pass is done with the buffers
                if (copier) {
                    // early rows: everything of this group's views that does not depend on where the eye stands
                    if (rp.rows_vec4) copy_rows<float4, true, true>(rp, p, grp, Hcur, kWarps - kCopyWarps, kCopyWarps);
                    else copy_rows<float, true, true>(rp, p, grp, Hcur, kWarps - kCopyWarps, kCopyWarps);
#ifdef ISY_RASTER_DETAIL
                    if (lane == 0) atomicMax(&s_tend[1], (unsigned long long)(clock64() - t0));
#endif
#if ISY_CULL_AHEAD
                    // ... then the cull of the NEXT position (the team is past this one's: its lists are filed), so that
                    // the three dependent trips to L2 of a cull are off the team's critical path
                    if (ahead_stage && grp == split && rp.T > 0) {
                        // (the copy warps decide together: the cull has barriers of its own)
                        const Team cw = Team{kCopyWarps * 32, 2};
                        cw.sync();
                        if (tid == tm.n) s_ahead_go = s_join[0] == (int)(item * (unsigned)rp.ngroups + (unsigned)grp) + 1;
                        cw.sync();
                        __threadfence_block();
                        const unsigned int nxt = *reinterpret_cast<volatile unsigned int*>(&s_item_next);
                        if (s_ahead_go && nxt < (unsigned)rp.items) {
                            const int gt = tid - tm.n;
                            cull_position(rp, Team{kCopyWarps * 32, 2}, (int)(nxt / (rp.splits * rp.ray_splits)), gt, vis, ahead_stage,
                                          rp.stage_cap, s_cull_ahead);
                            if (gt == 0) s_ahead_item = nxt;
                        }
                    }
#endif
                    // ... then the rasterisation, if the team has the copies filed by now
                    if (s_join[0] == (int)(item * (unsigned)rp.ngroups + (unsigned)grp) + 1) {
                        __threadfence_block();
                        const int how = s_join[1];
                        E = s_join[2];
                        int tcap;
                        if (how == 1) raster_tiles(rp, sm, exacts, tile_lists(sm, E, rp.ngroups == rp.splits, tcap), Hcur, &s_count, r_n, (unsigned)r_lo);
                        else if (how == 2) raster_copies<true>(rp, sm, exacts, E, Hcur, &s_count, r_n, (unsigned)r_lo);
                        else raster_copies<false>(rp, sm, exacts, E, Hcur, &s_count, r_n, (unsigned)r_lo);
                    }
                } else {
                if (best_smem) {
                    unsigned int* b32 = reinterpret_cast<unsigned int*>(sm.ras.best);
                    for (int i = tid; i < (Hcur * r_n + 1) / 2; i += tm.n) b32[i] = 0xffffffffu;
                } else {
                    for (int i = tid; i < Hcur * rp.R; i += tm.n) gbest[i] = kNoCopy;
                }
                if (walk_bands) {                  // this band's sorted rays and pitch look-up table
                    const float2* gs = reinterpret_cast<const float2*>(rp.grid_sorted) + r_lo;
                    for (int i = tid; i < r_n; i += tm.n) sm.ras.sorted[i] = gs[i];
                    const unsigned int* gp = rp.grid_plut + (size_t)band * (kPitchLut + 1);
                    for (int i = tid; i <= kPitchLut; i += tm.n) sm.ras.plut[i] = (unsigned short)gp[i];
                }
                tm.sync();
                if (tid == 0) s_count = 0;     // (the cull's counter is free again: next copy to rasterise)
                if (band == 0) uniform = sort_headings(sm, tm, Hcur);
                if (walk_bands) band_runs(sm, E);
                const bool scan = kEye == kEyeSmall ? Hcur > kSmallH : walk_bands;
#ifdef ISY_SETUP_DETAIL
                const long long t_sd1 = clock64();
#endif
                bool tiles = false;
                int tcap = 0;
                const unsigned short* const tlist = tile_lists(sm, E, !walk_bands && rp.ngroups == rp.splits, tcap);
                if (scan && uniform) {
                    build_phases(rp, sm, tm, Hcur, r_n, (unsigned)r_lo);   // (cached while the headings stay the same)
                    tm.sync();
#if defined(ISY_SETUP_DETAIL) && !defined(ISY_BIN_DETAIL)
                    t_sd[1] += clock64() - t_sd1;
#endif
                    tiles = !(rp.flags & ISY_FLAG_NO_TILES) && bin_copies(sm, tm, E, Hcur, r_n, s_warp, const_cast<unsigned short*>(tlist), tcap
#ifdef ISY_BIN_DETAIL
                                                                            , t_sd
#endif
                                                                            );
                }
                tm.sync();
#if defined(ISY_SETUP_DETAIL) && !defined(ISY_BIN_DETAIL)
                t_sd[0] += t_sd1 - t3b, t_sd[2] += clock64() - t_sd1;
#endif
                if (early_rows && tid == 0) {   // the copy warps may join
                    s_join[1] = tiles ? 1 : (uniform ? 2 : 3), s_join[2] = E;
                    __threadfence_block();
                    s_join[0] = (int)(item * (unsigned)rp.ngroups + (unsigned)grp) + 1;
                }
#ifdef ISY_PHASE_DETAIL
                t_walk = clock64();
#endif
                // rows mode: the copy warps write the sky rows now and the template rows of hit / depth / rgb when
                // kLateCopies copies are left (scans) -- the later, the likelier the patches find them in L2
                const bool copy_warp = rows_mode && !early_rows && warp >= kWarps - kCopyWarps;
#ifdef ISY_RASTER_DETAIL
                const long long t_rd0 = early_rows ? t0 : clock64();
#endif
                if (copy_warp) {
                    if (rp.rows_vec4) copy_rows<float4, true, false>(rp, p, grp, Hcur, kWarps - kCopyWarps, kCopyWarps);
                    else copy_rows<float, true, false>(rp, p, grp, Hcur, kWarps - kCopyWarps, kCopyWarps);
#if ISY_CULL_AHEAD
                    // ... and cull the NEXT position before they join the rasterisation
                    if (ahead_stage && grp == split && rp.T > 0) {
                        const unsigned int nxt = s_item_next;
                        if (nxt < (unsigned)rp.items) {
                            const int gt = tid - (kWarps - kCopyWarps) * 32;
                            cull_position(rp, Team{kCopyWarps * 32, 2}, (int)(nxt / (rp.splits * rp.ray_splits)), gt, vis, ahead_stage,
                                          rp.stage_cap, s_cull_ahead);
                            if (gt == 0) s_ahead_item = nxt;
                        }
                    }
#endif
                    if (scan) {
                        const int half = max(E - kLateCopies, 0);
                        if (tiles) raster_tiles(rp, sm, exacts, tlist, Hcur, &s_count, r_n, (unsigned)r_lo, (((r_n + kTileRays - 1) >> kTileShift) * Hcur * kLateTiles) >> 8);
                        else if (uniform) raster_copies<true>(rp, sm, exacts, E, Hcur, &s_count, r_n, (unsigned)r_lo, half);
                        else raster_copies<false>(rp, sm, exacts, E, Hcur, &s_count, r_n, (unsigned)r_lo, half);
                    }
                    if (rp.rows_vec4) copy_rows<float4, false, true>(rp, p, grp, Hcur, kWarps - kCopyWarps, kCopyWarps);
                    else copy_rows<float, false, true>(rp, p, grp, Hcur, kWarps - kCopyWarps, kCopyWarps);
                }
#ifdef ISY_RASTER_DETAIL
                if (copy_warp && lane == 0) atomicMax(&s_tend[1], (unsigned long long)(clock64() - t_rd0));
#endif
                if (scan) {
                    if (tiles) raster_tiles(rp, sm, exacts, tlist, Hcur, &s_count, r_n, (unsigned)r_lo);
                    else if (uniform) raster_copies<true>(rp, sm, exacts, E, Hcur, &s_count, r_n, (unsigned)r_lo);
                    else raster_copies<false>(rp, sm, exacts, E, Hcur, &s_count, r_n, (unsigned)r_lo);
#ifdef ISY_RASTER_DETAIL
                    if (lane == 0 && !copy_warp) {
                        atomicMax(&s_tend[0], (unsigned long long)(clock64() - t_rd0));
                        atomicAdd(&s_tend[2], (unsigned long long)(clock64() - t_rd0));
                    }
#endif
                } else if (kEye == kEyeSmall) {
                    raster_cells<true>(rp, sm, gbest, exacts, E, Hcur, &s_count);
                } else {   // large eye, one heading at a time
                    raster_cells<false>(rp, sm, gbest, exacts, E, Hcur, &s_count);
                }
                }   // (the team)
            }
            __syncthreads();
#ifdef ISY_RASTER_DETAIL
            // max end of the rasterising warps, end of the copy warps' rows, mean end of the rasterising warps
            if (tid == 0) {
                t_rd[0] += s_tend[0], t_rd[1] += s_tend[1], t_rd[2] += s_tend[2] / (kWarps - (rows_mode ? kCopyWarps : 0));
                s_tend[0] = 0, s_tend[1] = 0, s_tend[2] = 0;
            }
#endif
            if (copier) {   // what the team found
                E = s_E;
                mode = E <= kMaxE ? kModeRaster : kModeBrute;
                is_raster = kEye != kEyeAny && mode == kModeRaster;
            }
            long long t4 = clock64();
            t_phase[1] += t4 - t3b;
#ifdef ISY_PHASE_DETAIL
            t_detail[2] += t_walk - t3b;
#endif

            if (is_raster && !kF64 && !rp.init_c && (rp.S == 1 || kEye == kEyeLarge)) {
                const bool all_out = rp.out.rgb && rp.out.hit && rp.out.depth && rp.out.Y && rp.out.dop && rp.out.aop;
                // the common shapes get their own instantiations; everything else takes the general ones
                if (kEye == kEyeLarge) {
                    if (walk_bands && rp.S > 1) {   // a band of a cone-sampled eye
                        if (Hcur == 1) final_pass_fast<true, false, false, 1, true, true>(rp, sm, gbest, p, grp, Hcur, r_lo, r_n);
                        else if (rp.sun_per_view) final_pass_fast<true, false, false, kShadeU, true, true>(rp, sm, gbest, p, grp, Hcur, r_lo, r_n);
                        else final_pass_fast<true, true, false, kShadeU, true, true>(rp, sm, gbest, p, grp, Hcur, r_lo, r_n);
                    } else if (walk_bands) {        // a band of a large eye
                        if (Hcur == 1) final_pass_fast<true, false, false, 1, true>(rp, sm, gbest, p, grp, Hcur, r_lo, r_n);
                        else if (rp.sun_per_view) final_pass_fast<true, false, false, kShadeU, true>(rp, sm, gbest, p, grp, Hcur, r_lo, r_n);
                        else final_pass_fast<true, true, false, kShadeU, true>(rp, sm, gbest, p, grp, Hcur, r_lo, r_n);
                    } else if (rp.S > 1) {   // one heading at a time, winner buffer in HBM
                        final_pass_fast<false, false, false, 1, false, true>(rp, sm, gbest, p, grp, Hcur, 0, 0);
                    } else {
                        final_pass_fast<false, false, false, 1, false>(rp, sm, gbest, p, grp, Hcur, 0, 0);
                    }
                } else if (rows_mode) {   // the template rows are on their way: only the rays that hit something are left
                    final_pass_patch(rp, sm, p, grp, Hcur);
                } else if (Hcur > 1 && !rp.sun_per_view && all_out) {   // scans of a small eye, everything asked for
                    final_pass_fast<true, true, true, kShadeU, false>(rp, sm, gbest, p, grp, Hcur, 0, 0);
                } else if (Hcur == 1) {   // single views, sun sweeps: one heading per trip
                    final_pass_fast<true, false, false, 1, false>(rp, sm, gbest, p, grp, Hcur, 0, 0);
                } else {
                    if (!rp.sun_per_view) final_pass_fast<true, true, false, kShadeU, false>(rp, sm, gbest, p, grp, Hcur, 0, 0);
                    else final_pass_fast<true, false, false, kShadeU, false>(rp, sm, gbest, p, grp, Hcur, 0, 0);
                }
                t_phase[2] += clock64() - t4;
                continue;
            }
            // ---- generic final pass: warp units = (block of 32 consecutive rays) x (chunk of headings);
            //      the eye data of a ray are loaded once per unit and reused for its headings
            // (a single view of very many explicit rays -- world(pos, ori) with 1e5 directions -- is shared by ray_splits
            // CTAs: each takes a contiguous share of the 32-ray blocks, whole ommatidia)
            const int nb_all = (r_n + 31) >> 5;
            const int rb0 = (int)(((long long)nb_all * rsplit) / rp.ray_splits), rb1 = (int)(((long long)nb_all * (rsplit + 1)) / rp.ray_splits);
            const int nb = rb1 - rb0;
            int nch = min(Hcur, max(1, (2 * kWarps + max(nb, 1) - 1) / max(nb, 1)));
            const int hper = (Hcur + nch - 1) / nch;
            nch = (Hcur + hper - 1) / hper;
            for (int u = warp; u < nb * nch; u += kWarps) {
                const int rb = rb0 + u / nch, hc = u - (u / nch) * nch;
                const int r = r_lo + rb * 32 + lane;
                const bool valid = r < r_lo + r_n;
                const int rr = valid ? r : r_lo + r_n - 1;
                double e_pitch = 0, e_yaw = 0, f_theta = 0;
                double4 t4v = make_double4(0, 1, 0, 1);
                int pos = 0;
                if (yaw_only) {
                    e_pitch = rp.eye_py[2 * (size_t)rr + 0];
                    e_yaw = rp.eye_py[2 * (size_t)rr + 1];
                    if (perez) {
                        t4v = *reinterpret_cast<const double4*>(rp.eye_trig + 4 * (size_t)rr);
                        f_theta = rp.eye_ftheta[rr];
                    }
                    if (is_raster) pos = __ldg(rp.grid_pos_of + rr) - (walk_bands ? r_lo : 0);
                }
                const int hend = min(Hcur, (hc + 1) * hper);
                for (int hh = hc * hper; hh < hend; ++hh) {
                    const size_t b = (size_t)p * rp.H + grp + (size_t)hh * rp.ngroups;
                    Ray ray;
                    double dx = 0, dy = 0, dz = 0;
                    if (!yaw_only) {
                        const double* qv = rp.vquat + 4 * b;
                        const double* qo = rp.eye_quat + 4 * (size_t)rr;
                        double x1 = qv[0], y1 = qv[1], z1 = qv[2], w1 = qv[3];
                        double x2 = qo[0], y2 = qo[1], z2 = qo[2], w2 = qo[3];
                        double w = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2;
                        double x = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2;
                        double y = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2;
                        double z = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2;
                        quat_dir(x, y, z, w, dx, dy, dz);
                        dir_pitch_yaw(dx, dy, dz, ray.pitch, ray.yaw);
                        if (perez) f_theta = perez_f(dz, ray.pitch + kHalfPi < kHalfPi, rp.sky.A, rp.sky.B);
                    } else {
                        ray.pitch = e_pitch;
                        double yg = e_yaw + sm.head_val[hh];   // both in (-pi, pi]: one shift wraps
                        if (yg > kPi) yg -= kTwoPi;
                        else if (yg <= -kPi) yg += kTwoPi;
                        ray.yaw = yg;
                        if (perez) {
                            const double hs = sm.head_sin[hh], hcs = sm.head_cos[hh];
                            // sin/cos(yaw_e + heading) by the addition formulas
                            double sy = t4v.z * hcs + t4v.w * hs, cy = t4v.w * hcs - t4v.z * hs;
                            dx = t4v.y * cy, dy = t4v.y * sy, dz = -t4v.x;
                        }
                    }
                    ray.fp = (float)ray.pitch;
                    ray.fy = (float)ray.yaw;
                    ray.best = CUDART_INF;
                    ray.hit = -1;

                    if (is_raster) {
                        const int hs = sm.ras.hinv[hh];
                        const unsigned int rk = best_smem ? (unsigned)sm.ras.best[hs * r_n + pos]
                                                          : gbest[(size_t)hs * rp.R + pos];
                        if (rk != kNoCopy) {
                            const Entry& w = sm.tile[rk];
                            ray.hit = w.tri;
                            ray.best = __hiloint2double(w.dist_hi, w.dist_lo);
                        }
                    } else if (kEye == kEyeAny && mode == kModeBins) {
                        if (ray.fp >= grid.th_lo && ray.fp <= grid.th_hi) {
                            int row = min((int)((ray.fp - grid.th_lo) * grid.inv_dth), grid.rows - 1);
                            int col = min(max((int)((ray.fy + (float)kPi) * grid.inv_dph), 0), grid.cols - 1);
                            int bin = row * grid.cols + col;
                            unsigned i0 = sm.bins.bin_off[bin], i1 = sm.bins.bin_off[bin + 1];
                            for (unsigned i = i0; i < i1; ++i) {
                                int e = sm.bins.list[i];
                                test_entry(ray, &sm.tile[e], &exacts[e]);
                            }
                        }
                    } else if (E <= kMaxE) {
                        for (int j = 0; j < E; ++j) test_entry(ray, &sm.tile[j], &exacts[j]);
                    } else {   // more copies than the tile holds: stream them from the slab
                        for (int j = 0; j < E; ++j) test_entry(ray, j < kMaxE ? &sm.tile[j] : &entries[j], &exacts[j]);
                    }

                    RayOut o;
                    o.Y = 0, o.P = 0, o.A = CUDART_NAN, o.sA = CUDART_NAN, o.cA = CUDART_NAN;
                    if (perez) {
                        const SunConst sc = rp.sun[rp.sun_per_view ? b : 0];
                        sky_eval<kF64>(rp.sky, sc, dx, dy, dz, ray.pitch + kHalfPi, f_theta, o.Y, o.P, o.A, false, sm.gtab);
                        if (rp.S > 1) {
                            // A = atan2(dy - sy, dx - sx) + pi/2 (sky.py:323)  =>  (sin A, cos A) = (X, -Y) / |(X, Y)|
                            const double X = dx - sc.sx, Yd = dy - sc.sy, h2 = X * X + Yd * Yd;
                            const double ih = h2 > 0.0 ? rsqrt(h2) : 0.0;
                            o.sA = h2 > 0.0 ? X * ih : 1.0;        // atan2(0, 0) = 0 -> A = pi/2
                            o.cA = -Yd * ih;
                        }
                    } else if (want_sky) {
                        o.Y = rp.sky.luminance;   // UniformSky: y = luminance, p = 0, a = nan (sky.py:77-79)
                    }
                    // world colour (world.py:79-90, :123)
                    o.r = CUDART_NAN, o.g = CUDART_NAN, o.b = CUDART_NAN;
                    const bool from_sky = (rp.flags & ISY_FLAG_INIT_FROM_SKY) && want_sky;
                    if (rp.init_c || from_sky) {
                        double lum = rp.init_c ? rp.init_c[b * (size_t)rp.R + rr] : o.Y;
                        o.r = fmin(fmax(19.0 * lum + 13.0, 0.0), 255.0);      // luminance2skycolour, world.py:479
                        o.g = fmin(fmax(19.0 * lum + 135.0, 0.0), 255.0);
                        o.b = fmin(fmax(19.0 * lum + 201.0, 0.0), 255.0);
                    }
                    if (ray.pitch >= 0) o.r = 229.0 / 255.0, o.g = 183.0 / 255.0, o.b = 90.0 / 255.0;   // :90
                    if (ray.hit >= 0) {
                        const float* col = rp.tri_rgb + 3 * (size_t)ray.hit;
                        o.r = __ldg(col + 0), o.g = __ldg(col + 1), o.b = __ldg(col + 2);
                    }
                    o.sr = 0.f, o.sg = 0.f, o.sb = 0.f;
                    if (rp.out.resp) {
                        float pol = 1.f;
                        if (perez && rp.sensor.pol_op > 0.f) {
                            const SunConst& sc = rp.sun[rp.sun_per_view ? b : 0];
                            const double hxy = sqrt(dx * dx + dy * dy);
                            pol = sensor_pol_term(rp.sensor.pol_op, (float)o.P, (float)(dx - sc.sx), (float)(dy - sc.sy),
                                                  hxy > 0 ? (float)(dx / hxy) : 1.f, hxy > 0 ? (float)(dy / hxy) : 0.f);
                        }
                        sensor_colour(rp.sensor, ray.hit >= 0, (float)o.r, (float)o.g, (float)o.b, ray.pitch >= 0, want_sky, o.Y, pol,
                                      o.sr, o.sg, o.sb);
                    }
                    write_view<kF64>(rp, b, rr, valid, ray.hit, ray.best, o);
                }
            }
            t_phase[2] += clock64() - t4;
            }   // bands
        }
        t_phase[0] += t1 - t0, t_phase[1] += t2 - t1, t_phase[3] += 1;
    }
    if (tid == 0) {
#ifdef ISY_PHASE_DETAIL
        for (int k = 0; k < 3; ++k) t_phase[k] = t_detail[k];
#endif
#ifdef ISY_RASTER_DETAIL
        for (int k = 0; k < 3; ++k) t_phase[k] = t_rd[k];
#endif
#ifdef ISY_SETUP_DETAIL
        for (int k = 0; k < 3; ++k) t_phase[k] = t_sd[k];
#endif
        for (int k = 0; k < 4; ++k) atomicAdd(&rp.phase_cycles[k], (unsigned long long)t_phase[k]);
    }
}

}  // namespace isy

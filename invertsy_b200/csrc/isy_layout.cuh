// Tunables and the shared-memory layout of the render kernel.
#pragma once
#include <cstddef>
#include "isy_common.cuh"

namespace isy {

constexpr int kThreads = ISY_THREADS;  // CTA size, one persistent CTA per SM
constexpr int kWarps = kThreads / 32;
constexpr int kMaxE = 1152;           // projected copies resident in shared memory (54 KB)
constexpr int kBinRows = 32;          // finest angular grid over the vegetation band
constexpr int kBinCols = 128;         // ... and over azimuth [-pi, pi]
constexpr int kBins = kBinRows * kBinCols;
constexpr int kListCap = 24576;       // candidate list entries (u16, 48 KB)
constexpr int kMaxH = 128;            // headings whose sin/cos are cached per position
constexpr float kBinMargin = 1e-4f;   // rad; >> every fp32 rounding in the bin arithmetic
#ifndef ISY_TALL
#define ISY_TALL 0.35f
#endif
constexpr float kTallCopy = ISY_TALL;    // rad of pitch: copies taller than this are rasterised first (they take longest)
constexpr float kBinReject = -1e-5f;  // a bin is dropped only if an edge function is below this on all of it

constexpr int kBestCap = 36864;       // (heading, ray) slots of the rasteriser's winner buffer (u16 copy slots, 72 KB)
constexpr int kBandRays = 1024;       // larger eyes are rendered in bands of this many consecutive rays (a multiple of 32)
constexpr int kGridRaysSmem = 1280;   // eyes up to this many rays keep their pitch-sorted ray table in shared memory
constexpr int kPitchLut = 1024;       // slots of the pitch -> first sorted ray look-up table
constexpr int kHeadLut = 1024;        // slots of the azimuth -> first sorted heading look-up table
constexpr int kCellsSmem = 2048;      // cells of the eye's (pitch, yaw) ray grid kept in shared memory
constexpr int kMaxGiants = 254;       // (copy, heading) pairs the cell walk can set aside per group
constexpr int kSmallH = 4;            // up to this many headings per group the rasteriser walks the ray grid
constexpr int kMaxTiles = kBestCap / 32 + kMaxH;   // (ceil(R / 32) x H) tiles of a scan whose winner buffer holds H x R slots

// The threads [0, n) of the CTA, synchronised on hardware barrier `id` (0: the whole CTA, __syncthreads).  In rows mode
// the last kCopyWarps warps stream the position's rows from the moment it starts, and everything up to the
// rasterisation is done by the team of the others.
struct Team {
    int n, id;
    __device__ __forceinline__ void sync() const {
        if (id == 0) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
    }
    __device__ __forceinline__ bool sync_and(bool pred) const {
        if (id == 0) return __syncthreads_and(pred) != 0;
        int r;
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %3, 0;\n\tbar.red.and.pred q, %1, %2, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
                     : "=r"(r) : "r"(id), "r"(n), "r"((int)pred) : "memory");
        return r != 0;
    }
};

struct SmemBins {
    unsigned int bin_off[kBins + 4];  // exclusive offsets into list (counts during the count pass)
    unsigned int bin_cur[kBins];      // fill cursors
    unsigned short list[kListCap];
};

struct SmemRaster {
    unsigned short best[kBestCap];    // nearest containing copy (slot in tile) per (sorted heading, sorted ray)
    float2 sorted[kGridRaysSmem];     // the eye's rays in pitch order: (pitch, yaw) as floats (copied once per CTA)
    unsigned short plut[kPitchLut + 4]; // first sorted ray at or above each pitch slot
    float hsort[2 * kMaxH];           // wrapped headings of the group, ascending, then the same + 2 pi
    float hvalf[2 * kMaxH];           // the wrapped heading itself (no + 2 pi) of each sorted slot
    unsigned char hperm[2 * kMaxH];   // heading index of each sorted slot
    unsigned char hinv[kMaxH];        // sorted slot of each heading
    unsigned short hlut[kHeadLut];    // first sorted slot that can reach each azimuth slot
    float uni[4];                     // {h0, delta, 1/delta, flag}: sorted headings are h0 + k*delta over the full circle
    unsigned short cell_start[kCellsSmem + 4];    // the eye's static (pitch, yaw) ray grid: first list entry of each cell
    unsigned short cell_list[kGridRaysSmem];      // ... and the pitch-sorted positions of the rays in each cell
};

// Scans with evenly spaced headings (column walk / tile walk, isy_raster.cuh).  Launches that cannot scan -- single views,
// routes, general orientations -- do not allocate it: what they leave of the 228 KB is their L1.
struct SmemScan {
    // column walk (evenly spaced headings h0 + k*delta): the query azimuth of (ray j, sorted heading k) is
    // phase[j] + c*delta - pi in column c = (colb[j] + k) mod H -- every column holds every ray exactly once
    float phase[kGridRaysSmem];       // yaw_j + h0 + pi folded into [0, delta)
    unsigned char colb[kGridRaysSmem];// floor((yaw_j + h0 + pi) / delta) mod H
    float coloff[kMaxH + 2];          // c*delta - pi for c = -1 .. H
    double ph_h0;                     // what phase[] / colb[] / coloff[] were built for: first sorted heading,
    int ph_H, ph_lo;                  // ... number of headings, first sorted ray of the band
    // tile walk: a tile = 32 consecutive pitch-sorted rays x one column; tile_cnt[t] = end of tile t's copy list
    unsigned int tile_cnt[kMaxTiles + 1];
    unsigned short crange[kMaxE];     // columns of each copy: (first + 1) | (number << 8)
};

struct SmemLayout {
    Entry tile[kMaxE];
    float4 bbox[kMaxE];               // (theta_lo, theta_hi, phi_lo, phi_hi) of each copy, margin included
    unsigned short order[kMaxE];      // rasterisation order: tall copies from the front, the rest from the back
                                      // (the tile walk keeps its copy lists in bbox + order once the ranges are taken)
    double head_sin[kMaxH], head_cos[kMaxH], head_val[kMaxH];   // per heading of the current group
    double2 gtab[2 * (kGTab + 1)];    // exp(D acos c) table of this call's sky (copied once per CTA)
    SunConst sun0;                    // sun constants when all views share one sun (copied once per CTA)
    int order_n[2];                   // ... and how many of each
    unsigned short giants[kMaxGiants]; // cell walk: (copy, heading) pairs set aside for the whole CTA
    int n_giants;
    unsigned int jrange[kMaxE];       // pitch run of each copy in the eye's sorted rays: first | (end << 16) (smem-grid eyes)
    union {
        SmemBins bins;
        SmemRaster ras;
    };
    SmemScan scan;                    // (launches whose groups hold more than kSmallH headings)
    // rows mode only (isy_shade.cuh): what final_pass_patch looks up per hit, so that it issues no global load.
    // Launches that are not in rows mode do not allocate this tail (kSmemBase bytes of dynamic shared memory).
    float rgb[kMaxE * 3];                         // colour of each copy's triangle
    unsigned short sorted_ray16[kGridRaysSmem];   // ray index of each pitch-sorted position
    float resp1[kMaxE];                           // PN response of a ray that hits each copy's triangle (out.resp, 1 channel)
};
constexpr size_t kSmemBase = offsetof(SmemLayout, scan);   // no scans, no rows mode
constexpr size_t kSmemScan = offsetof(SmemLayout, rgb);    // scans, per-view final passes
static_assert(kSmemBase <= 192 * 1024, "leave some of the 228 KB for L1");
static_assert(sizeof(SmemLayout) <= 224 * 1024, "rows mode: 227 KB per CTA is the limit");

constexpr int kTileListCap = (int)((sizeof(float4) + sizeof(unsigned short)) * kMaxE / sizeof(unsigned short));   // 10 368 entries
static_assert(offsetof(SmemLayout, order) == offsetof(SmemLayout, bbox) + sizeof(float4) * kMaxE, "the tile lists span bbox + order");

// ---- tile walk (isy_raster.cuh): tiles of kTileRays consecutive pitch-sorted rays x one azimuth column
#ifndef ISY_BIN_SOLO
#define ISY_BIN_SOLO 8
#endif
#ifndef ISY_TILE_U
#define ISY_TILE_U 2
#endif
constexpr int kBinSolo = ISY_BIN_SOLO;       // a copy that touches up to this many tiles is filed by its own lane
constexpr int kTileU = ISY_TILE_U;           // queries per lane in a tile
constexpr int kTileRays = 32 * kTileU, kTileShift = kTileU == 1 ? 5 : (kTileU == 2 ? 6 : 7);
static_assert(kTileU == 1 || kTileU == 2 || kTileU == 4, "tiles of 32, 64 or 128 rays");

// place of chunk q in the walk's queue: the chunks at and above the horizon (the middle of a pitch-sorted eye) hold
// nearly all the work and go first, from the horizon upwards; the chunks below it follow
__device__ __forceinline__ int chunk_place(int q, int q_h) { return q <= q_h ? q_h - q : q; }

// columns of a copy with azimuth box [plo, phi] (margin included): (first + 1) | (number << 8); -1 and H alias H - 1 and 0
__device__ __forceinline__ unsigned int copy_columns(float plo, float phi, float inv_delta, int Hcur) {
    const float pi_f = (float)kPi;
    const float glo = fmaxf(plo, -pi_f - kBinMargin) - kBinMargin, ghi = fminf(phi, pi_f + kBinMargin) + kBinMargin;
    if (!(glo <= ghi)) return 0u;
    const int c_lo = (int)floorf((glo + pi_f) * inv_delta), c_hi = (int)floorf((ghi + pi_f) * inv_delta);
    return (unsigned)(c_lo + 1) | ((unsigned)min(c_hi - c_lo + 1, Hcur) << 8);
}

struct RayOut {
    double r, g, b;
    double Y, P, A;
    double sA, cA;   // sin / cos of A for the cone average (S > 1), formed without the angle itself
    float sr, sg, sb; // the sample's colour on the 0..1 scale for the photoreceptor transform (out.resp)
};

template <typename T>
__device__ __forceinline__ void store_val(void* base, size_t idx, double v) {
    __stcs(reinterpret_cast<T*>(base) + idx, (T)v);   // streaming: written once, never read back by the kernel
}

}  // namespace isy

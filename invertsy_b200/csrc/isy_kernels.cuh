// Kernels of the sm_100a view renderer (persistent, one CTA per position work item).
//
//   phase A1  cull      : min-vertex distance of every world triangle to the eye, fp64,
//                         reference order (world.py:94-98)                      -> visible list
//   phase A2  project   : vertex (theta, phi) by atan2, seam rule, reference crosses,
//                         normalised fp32 edge functions (world.py:109-120, 482-526)
//                                                                -> entry tile in shared memory
//   phase A3  bin       : conservative rasterisation of the projected copies into a
//                         (theta, phi) grid over the vegetation band (count / scan / fill)
//   phase B   rays      : every ray of every orientation at this position looks up its bin and
//                         tests only that bin's copies: fp32 filter + exact fp64 re-check near
//                         edges; argmin of distance over containing copies == the painter's
//                         far->near overwrite (world.py:106-123); ground / background rule
//                         (world.py:79-90); analytic sky per ray (sky.py:304-334); cone
//                         average over the S samples of an ommatidium; ONE write of the view.
//
// ISY_FLAG_BRUTE_FORCE (and any position whose copies do not fit the shared-memory budget)
// skips A3 and streams every copy past every ray instead; results are identical.
#pragma once
#include "isy_common.cuh"

#ifndef ISY_THREADS
#define ISY_THREADS 512
#endif

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
constexpr float kBinReject = -1e-5f;  // a bin is dropped only if an edge function is below this on all of it

constexpr int kBestCap = 36864;       // (heading, ray) slots of the rasteriser's winner buffer (u16 copy slots, 72 KB)
constexpr int kGridRaysSmem = 1280;   // eyes up to this many rays keep their pitch-sorted ray table in shared memory
constexpr int kPitchLut = 1024;       // slots of the pitch -> first sorted ray look-up table
constexpr int kHeadLut = 1024;        // slots of the azimuth -> first sorted heading look-up table
constexpr int kCellsSmem = 2048;      // cells of the eye's (pitch, yaw) ray grid kept in shared memory
constexpr int kSmallH = 4;            // up to this many headings per group the rasteriser walks the ray grid

struct SmemBins {
    unsigned int bin_off[kBins + 4];  // exclusive offsets into list (counts during the count pass)
    unsigned int bin_cur[kBins];      // fill cursors
    unsigned short list[kListCap];
};

struct SmemRaster {
    unsigned short best[kBestCap];    // nearest containing copy (slot in tile) per (sorted heading, sorted ray)
    float2 sorted[kGridRaysSmem];     // the eye's rays in pitch order: (pitch, yaw) as floats (copied once per CTA)
    unsigned int plut[kPitchLut + 4]; // first sorted ray at or above each pitch slot
    float hsort[2 * kMaxH];           // wrapped headings of the group, ascending, then the same + 2 pi
    float hvalf[2 * kMaxH];           // the wrapped heading itself (no + 2 pi) of each sorted slot
    unsigned char hperm[2 * kMaxH];   // heading index of each sorted slot
    unsigned char hinv[kMaxH];        // sorted slot of each heading
    unsigned short hlut[kHeadLut];    // first sorted slot that can reach each azimuth slot
    float uni[4];                     // {h0, delta, 1/delta, flag}: sorted headings are h0 + k*delta over the full circle
    unsigned int cell_start[kCellsSmem + 4];      // the eye's static (pitch, yaw) ray grid: first list entry of each cell
    unsigned short cell_list[kGridRaysSmem];      // ... and the pitch-sorted positions of the rays in each cell
};

struct SmemLayout {
    Entry tile[kMaxE];
    float4 bbox[kMaxE];               // (theta_lo, theta_hi, phi_lo, phi_hi) of each copy, margin included
    double head_sin[kMaxH], head_cos[kMaxH], head_val[kMaxH];   // per heading of the current group
    double2 gtab[2 * (kGTab + 1)];    // exp(D acos c) table of this call's sky (copied once per CTA)
    SunConst sun0;                    // sun constants when all views share one sun (copied once per CTA)
    union {
        SmemBins bins;
        SmemRaster ras;
    };
};
static_assert(sizeof(SmemLayout) <= 192 * 1024, "leave some of the 228 KB for L1");

struct RayOut {
    double r, g, b;
    double Y, P, A;
};

template <typename T>
__device__ __forceinline__ void store_val(void* base, size_t idx, double v) {
    reinterpret_cast<T*>(base)[idx] = (T)v;
}

// eye-frame (pitch, yaw) and their sines / cosines from the ommatidia quaternions, once per eye
__global__ void eye_setup_kernel(const double* __restrict__ quat, double* __restrict__ py, double* __restrict__ trig,
                                 int R) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    double dx, dy, dz, p, y;
    quat_dir(quat[4 * r + 0], quat[4 * r + 1], quat[4 * r + 2], quat[4 * r + 3], dx, dy, dz);
    dir_pitch_yaw(dx, dy, dz, p, y);
    py[2 * r + 0] = p;
    py[2 * r + 1] = y;
    double s, c;
    sincos(p, &s, &c);
    trig[4 * r + 0] = s, trig[4 * r + 1] = c;
    sincos(y, &s, &c);
    trig[4 * r + 2] = s, trig[4 * r + 3] = c;
}

// gradation term f(theta) of the Perez function per eye ray (depends on pitch only for yaw-only views)
__global__ void eye_sky_kernel(const double* __restrict__ py, const double* __restrict__ trig, int R,
                               isy_sky_params sp, double* __restrict__ f_theta) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    double theta = py[2 * r] + kHalfPi;
    f_theta[r] = perez_f(-trig[4 * r + 0], theta < kHalfPi, sp.A, sp.B);   // cos(theta) = -sin(pitch)
}

// Hermite table of G(c) = exp(D acos c) in u = sqrt(1 - |c|), see g_table_eval
__global__ void sky_table_kernel(double D, double2* __restrict__ tab) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * (kGTab + 1)) return;
    const int half = i / (kGTab + 1), k = i - half * (kGTab + 1);
    const double h = 1.0 / kGTab, u = k * h;
    const double g = 2.0 * asin(u * 0.70710678118654752440);       // acos(1 - u^2)
    const double dg = 2.0 / sqrt(2.0 - u * u);                     // d gamma / du
    double2 v;
    if (half == 0) {
        v.x = exp(D * g);
        v.y = v.x * D * dg * h;
    } else {                                                       // c < 0: gamma = pi - g
        v.x = exp(D * (kPi - g));
        v.y = -v.x * D * dg * h;
    }
    tab[i] = v;
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
    c.inv_i00 = 1.0 / (c.i00 + kEps);
    c.pad_ = 0.f;
    double chi = (4.0 / 9.0 - sp.tau_L / 120.0) * (kPi - 2 * (kHalfPi - ths));
    c.Yz = (4.0453 * sp.tau_L - 4.9710) * tan(chi) - 0.2155 * sp.tau_L + 2.4192;
    c.Mp = exp(-(sp.tau_L - sp.c1) / (sp.c2 + kEps));
    c.inv_i00_f = (float)c.inv_i00, c.K_f = (float)c.K, c.Mp2_f = (float)(2.0 / kPi * c.Mp);
    out[b] = c;
}

// ------------------------------------------------------------------------------------------
// phase A: cull + project one position. Entries go to the CTA's slab (always) and to the
// shared-memory tile (first kMaxE). Returns the number of copies E.
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

__device__ __forceinline__ void emit_copy(const double th[3], const double ph[3], double dist, int t, int slot,
                                          Entry* entries, Exact* exacts, SmemLayout& sm, int* s_band) {
    Entry e;
    finish_copy(th, ph, dist, t, e, exacts[slot]);
    entries[slot] = e;
    if (slot < kMaxE) {
        sm.tile[slot] = e;
        float tlo = (float)fmin(th[0], fmin(th[1], th[2])) - kBinMargin;
        float thi = (float)fmax(th[0], fmax(th[1], th[2])) + kBinMargin;
        float plo = (float)fmin(ph[0], fmin(ph[1], ph[2])) - kBinMargin;
        float phi = (float)fmax(ph[0], fmax(ph[1], ph[2])) + kBinMargin;
        sm.bbox[slot] = make_float4(tlo, thi, plo, phi);
        atomicMin(&s_band[0], float_key(tlo));
        atomicMax(&s_band[1], float_key(thi));
    }
}

__device__ int cull_and_project(const RenderParams& rp, int p, int* vis, Entry* entries, Exact* exacts,
                                SmemLayout& sm, int* s_count, int* s_nent, int* s_band) {
    const int tid = threadIdx.x, lane = tid & 31;
    const double px = rp.pos[3 * p + 0], py = rp.pos[3 * p + 1], pz = rp.pos[3 * p + 2];
    const double h = rp.horizon;
    const double h2_hi = h * h * (1.0 + 1e-12);  // above this the rounded norm cannot be <= horizon

    if (tid == 0) {
        *s_count = 0;
        *s_nent = 0;
        s_band[0] = float_key(CUDART_INF_F);
        s_band[1] = float_key(-CUDART_INF_F);
    }
    __syncthreads();

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
    // the candidate cells form one contiguous list range per grid row: flatten the (<= 16) rows
    __shared__ unsigned int s_rowbeg[16], s_rowoff[17];
    const int nrows = min(cy1 - cy0 + 1, 16);   // cell = horizon / 4 or larger: at most 10 rows
    if (tid == 0) {
        unsigned int acc = 0;
        for (int rr = 0; rr < nrows; ++rr) {
            const unsigned int beg = __ldg(rp.wg_cell_start + (size_t)(cy0 + rr) * nx + cx0);
            const unsigned int end = __ldg(rp.wg_cell_start + (size_t)(cy0 + rr) * nx + cx1 + 1);
            s_rowbeg[rr] = beg;
            s_rowoff[rr] = acc;
            acc += end - beg;
        }
        s_rowoff[nrows] = acc;
    }
    __syncthreads();
    const unsigned int ncand = s_rowoff[nrows];
    for (unsigned int base = 0; base < ncand; base += kThreads) {   // block-uniform trip count
        const unsigned int c = base + tid;
        bool visible = false;
        int t = 0;
        if (c < ncand) {
            int rr = 0;
            while (rr + 1 < nrows && c >= s_rowoff[rr + 1]) ++rr;
            const unsigned int packed = __ldg(rp.wg_cell_tris + s_rowbeg[rr] + (c - s_rowoff[rr]));
            t = (int)(packed & 0x3fffffffu);
            const int tag = (int)(packed >> 30);
            double v[3][3];
            load_vertices(rp.tri_planes, rp.T, t, px, py, pz, v);
            const double s0 = norm2_rn(v[0][0], v[0][1], v[0][2]);
            const double s1 = norm2_rn(v[1][0], v[1][1], v[1][2]);
            const double s2 = norm2_rn(v[2][0], v[2][1], v[2][2]);
            const double smin = fmin(s0, fmin(s1, s2));
            if (smin <= h2_hi && __dsqrt_rn(smin) <= h) {
                const int kn = (s0 <= s1 && s0 <= s2) ? 0 : (s1 <= s2 ? 1 : 2);   // nearest vertex
                // same float arithmetic as the host used to file the vertices into cells
                const float* pl = rp.tri_planes;
                const float xn = __ldg(pl + (size_t)(3 * kn) * rp.T + t), yn = __ldg(pl + (size_t)(3 * kn + 1) * rp.T + t);
                const float xt = __ldg(pl + (size_t)(3 * tag) * rp.T + t), yt = __ldg(pl + (size_t)(3 * tag + 1) * rp.T + t);
                const int cnx = world_cell(xn, gx0, ginv, nx), cny = world_cell(yn, gy0, ginv, ny);
                const int ctx = world_cell(xt, gx0, ginv, nx), cty = world_cell(yt, gy0, ginv, ny);
                visible = cnx == ctx && cny == cty;
            }
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
                emit_copy(th, p1, dist, t, slot, entries, exacts, sm, s_band);
                emit_copy(th, p2, dist, t, slot + 1, entries, exacts, sm, s_band);
            } else {
                emit_copy(th, ph, dist, t, slot, entries, exacts, sm, s_band);
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
constexpr int kRU = 4;                // rays per lane and iteration in the rasteriser

// is copy e (distance d, triangle tri) nearer than copy c? (ties: lower triangle index, then lower slot)
__device__ __forceinline__ bool nearer(const Entry* tile, double d, int tri, unsigned int e, unsigned int c) {
    const Entry& o = tile[c];
    const double dc = __hiloint2double(o.dist_hi, o.dist_lo);
    return d < dc || (d == dc && (tri < o.tri || (tri == o.tri && e < c)));
}

// best[idx] = nearer(e, best[idx]) ? e : best[idx], atomically (CAS on the 32-bit word holding the u16)
__device__ __forceinline__ void post_copy_u16(unsigned short* best, unsigned int idx, const Entry* tile, double d, int tri,
                                              unsigned int e) {
    unsigned int* word = reinterpret_cast<unsigned int*>(best) + (idx >> 1);
    const bool hi = idx & 1u;
    unsigned int old = *reinterpret_cast<volatile unsigned int*>(word);
    for (;;) {
        const unsigned int cur = hi ? (old >> 16) : (old & 0xffffu);
        if (cur != kNoCopy && !nearer(tile, d, tri, e, cur)) return;
        const unsigned int nw = hi ? ((old & 0xffffu) | (e << 16)) : ((old & 0xffff0000u) | e);
        const unsigned int prev = atomicCAS(word, old, nw);
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

// sorts the group's wrapped headings (head_val[0..Hcur)) and builds the azimuth look-up table
__device__ void sort_headings(SmemLayout& sm, int Hcur) {
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
    __syncthreads();
    if (tid == 0) {
        // evenly spaced headings over the full circle (every scan of the reference's simulations, and a
        // single heading) let the window walk use arithmetic instead of the look-up table
        const float h0 = sm.ras.hsort[0];
        const float delta = (float)kTwoPi / (float)Hcur;
        bool ok = true;
        for (int k = 1; k < Hcur; ++k) ok = ok && fabsf(sm.ras.hsort[k] - (h0 + k * delta)) < 5e-7f;
        sm.ras.uni[0] = h0, sm.ras.uni[1] = delta, sm.ras.uni[2] = 1.0f / delta, sm.ras.uni[3] = ok ? 1.f : 0.f;
    }
    for (int s = tid; s < kHeadLut; s += kThreads) {
        // first sorted slot whose heading is not below this azimuth slot (minus a guard band):
        // windows start their walk here; whatever they pick up below `lower` just fails the edge test
        const float start = -(float)kPi + s * ((float)kTwoPi / kHeadLut) - 1e-4f;
        int k = 0;
        while (k < 2 * Hcur && sm.ras.hsort[k] < start) ++k;
        sm.ras.hlut[s] = (unsigned short)k;
    }
}

// kSmem: the eye's sorted ray table and the winner buffer live in shared memory (u16 slots);
// otherwise both are in global memory (u32 slots), for eyes too large for that.
// The hot path is fp32: query azimuth = fl(yaw_f + heading_f) wrapped in fp32; kEdgeEps covers that
// rounding.  Rays that land within 3e-6 rad of the +-pi seam, and rays inside the rounding band of
// an edge, redo the query in fp64 exactly as the reference composes it.
template <bool kSmem, bool kUniform>
__device__ void raster_copies(const RenderParams& rp, SmemLayout& sm, unsigned int* gbest, const Exact* exacts, int E,
                              int Hcur, int* s_next) {
    const int R = rp.R, lane = threadIdx.x & 31;
    const float2* gsorted = reinterpret_cast<const float2*>(rp.grid_sorted);
    const int H2 = 2 * Hcur;
    const float pi_f = (float)kPi, twopi_f = (float)kTwoPi;
    const float uh0 = sm.ras.uni[0], udelta = sm.ras.uni[1], uinv = sm.ras.uni[2];
    for (;;) {
        // copies differ a lot in size (a near blade spans hundreds of rays): warps pull them from a counter
        int e = 0;
        if (lane == 0) e = smem_atomic_add(s_next, 1);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e >= E) break;
        const float4 bb = sm.bbox[e];
        // azimuth window of the copy clipped to the query domain (-pi, pi]
        const float glo = fmaxf(bb.z, -pi_f - kBinMargin) - kBinMargin, ghi = fminf(bb.w, pi_f + kBinMargin) + kBinMargin;
        if (glo > ghi) continue;
        const float width = ghi - glo;
        const Entry ent = sm.tile[e];
        const double dist = __hiloint2double(ent.dist_hi, ent.dist_lo);
        // run of pitch-sorted rays that covers [bb.x, bb.y] (a few extra rays at both ends)
        const int s0 = min(max((int)floorf((bb.x + (float)kHalfPi) * (kPitchLut / (float)kPi)), 0), kPitchLut - 1);
        const int s1 = min(max((int)floorf((bb.y + (float)kHalfPi) * (kPitchLut / (float)kPi)) + 2, 0), kPitchLut);
        const unsigned jbeg = kSmem ? sm.ras.plut[s0] : __ldg(rp.grid_plut + s0);
        const unsigned jend = kSmem ? sm.ras.plut[s1] : __ldg(rp.grid_plut + s1);
        // kRU rays per lane and iteration: their window walks and edge tests are independent
        // instruction streams, which is what hides the shared-memory and branch latencies here
        for (unsigned jb = jbeg; jb < jend; jb += 32 * kRU) {   // warp-uniform trip count
            __syncwarp();                                       // keep the lanes converged across iterations
            float2 py[kRU];
            int k[kRU], ks[kRU];    // window cursor; sorted slot of the heading under the cursor
            float upper[kRU];
#pragma unroll
            for (int u = 0; u < kRU; ++u) {
                const unsigned j = jb + u * 32 + lane;
                py[u] = make_float2(4.f, 0.f);                  // pitch 4 rad: outside every band
                if (j < jend) py[u] = kSmem ? sm.ras.sorted[j] : __ldg(gsorted + j);
            }
#pragma unroll
            for (int u = 0; u < kRU; ++u) {
                // headings h with yaw + h in [glo, ghi]  <=>  h in [lower, lower + width] on the circle
                float lower = glo - py[u].y;
                if (lower < -pi_f) lower += twopi_f;
                else if (lower >= pi_f) lower -= twopi_f;
                const bool band = py[u].x >= bb.x && py[u].x <= bb.y;
                if (kUniform) {
                    // headings are h0 + k*delta (k mod H): the window is the index range [k0, k1];
                    // k[] counts the headings left in the window, ks[] is the sorted slot of the next one
                    const float a = (lower - uh0) * uinv;
                    const int k0 = (int)ceilf(a - 2e-4f), k1 = (int)floorf(fmaf(width, uinv, a) + 2e-4f);
                    k[u] = band ? min(k1 - k0 + 1, Hcur) : 0;            // lower in [-pi, pi): k0 in [-H, H]
                    ks[u] = k0 < 0 ? k0 + Hcur : (k0 >= Hcur ? k0 - Hcur : k0);
                    upper[u] = 0.f;
                } else {
                    const int ls = min(max((int)((lower + pi_f) * (kHeadLut / twopi_f)), 0), kHeadLut - 1);
                    k[u] = band ? (int)sm.ras.hlut[ls] : H2;
                    upper[u] = lower + width;
                }
            }
            // lanes walk their heading windows in lock step (most windows hold 0 or 1 heading)
            for (;;) {
                bool act[kRU], any = false;
                float hv[kRU];
#pragma unroll
                for (int u = 0; u < kRU; ++u) {
                    if (kUniform) {
                        act[u] = k[u] > 0;
                        hv[u] = fmaf((float)ks[u], udelta, uh0);
                    } else {
                        const int kk = min(k[u], H2 - 1);
                        act[u] = k[u] < H2 && sm.ras.hsort[kk] <= upper[u];
                        hv[u] = sm.ras.hvalf[kk];
                        ks[u] = kk >= Hcur ? kk - Hcur : kk;       // the + 2 pi copy is the same heading
                    }
                    any |= act[u];
                }
                if (!__any_sync(0xffffffffu, any)) break;
                float fy[kRU], m[kRU];
#pragma unroll
                for (int u = 0; u < kRU; ++u) {                 // branch-free: the common case of every slot
                    float f = py[u].y + hv[u];                  // both in (-pi, pi]: one shift wraps
                    f = f > pi_f ? f - twopi_f : (f <= -pi_f ? f + twopi_f : f);
                    fy[u] = f;
                    const float e0 = fmaf(ent.A0, py[u].x, fmaf(ent.B0, f, ent.C0));
                    const float e1 = fmaf(ent.A1, py[u].x, fmaf(ent.B1, f, ent.C1));
                    const float e2 = fmaf(ent.A2, py[u].x, fmaf(ent.B2, f, ent.C2));
                    m[u] = fminf(e0, fminf(e1, e2));
                }
#pragma unroll
                for (int u = 0; u < kRU; ++u) {
                    const unsigned j = jb + u * 32 + lane;
                    const bool seam = fabsf(fy[u]) > 3.14159f;
                    bool in = act[u] && m[u] > kEdgeEps && !seam;
                    // rare: the rounding band of an edge, or a query within 3e-6 rad of the seam ->
                    // redo the query in fp64 exactly as the reference composes it
                    if (act[u] && (seam || (m[u] >= -kEdgeEps && m[u] <= kEdgeEps))) {
                        const int hh = sm.ras.hperm[ks[u]];
                        const int ray = __ldg(rp.grid_sorted_ray + j);
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
                        if (kSmem) post_copy_u16(sm.ras.best, (unsigned)ks[u] * R + j, sm.tile, dist, ent.tri, e);
                        else post_copy_u32(&gbest[(size_t)ks[u] * R + j], sm.tile, dist, ent.tri, e);
                    }
                    if (kUniform) {
                        if (act[u]) {
                            --k[u];
                            ks[u] = ks[u] + 1 == Hcur ? 0 : ks[u] + 1;
                        }
                    } else if (act[u]) {
                        ++k[u];
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Few headings per position (single views, route batches): walking every ray of a copy's pitch band
// would test ~98 % of the pairs for nothing, so each (copy, heading) pair instead visits the few
// cells of the eye's static (pitch, yaw) ray grid under the copy's bounding box, shifted by the
// heading.  Lanes hold different pairs; the work per position is a few thousand warp instructions.
// ------------------------------------------------------------------------------------------
// tests the rays stored in one cell of the eye's ray grid against one projected copy at one heading
template <bool kSmem>
__device__ __forceinline__ void raster_cell(const RenderParams& rp, SmemLayout& sm, unsigned int* gbest,
                                            const Exact* exacts, int cell, const Entry& ent, double dist, int e, int hh,
                                            int hs, float hf) {
    const int R = rp.R;
    const float2* gsorted = reinterpret_cast<const float2*>(rp.grid_sorted);
    const float pi_f = (float)kPi, twopi_f = (float)kTwoPi;
    const unsigned i0 = kSmem ? sm.ras.cell_start[cell] : __ldg(rp.cells_start + cell);
    const unsigned i1 = kSmem ? sm.ras.cell_start[cell + 1] : __ldg(rp.cells_start + cell + 1);
    for (unsigned i = i0; i < i1; ++i) {
        const unsigned j = kSmem ? (unsigned)sm.ras.cell_list[i] : __ldg(rp.cells_list + i);
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
            if (kSmem) post_copy_u16(sm.ras.best, (unsigned)hs * R + j, sm.tile, dist, ent.tri, e);
            else post_copy_u32(&gbest[(size_t)hs * R + j], sm.tile, dist, ent.tri, e);
        }
    }
}

template <bool kSmem>
__device__ void raster_cells(const RenderParams& rp, SmemLayout& sm, unsigned int* gbest, const Exact* exacts, int E,
                             int Hcur, int* s_next) {
    const int lane = threadIdx.x & 31, Gp = rp.cells_rows, Gy = rp.cells_cols;
    const float pi_f = (float)kPi, twopi_f = (float)kTwoPi;
    const float inv_dp = (float)Gp / pi_f, inv_dy = (float)Gy / twopi_f;
    const int items = E * Hcur;
    constexpr int kBigItem = 12;    // pairs that cover more cells than this are walked by the whole warp
    for (;;) {
        int base = 0;
        if (lane == 0) base = smem_atomic_add(s_next, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= items) break;
        const int item = base + lane;
        int e = 0, hh = 0, r0 = 0, nrows = 0, c0 = 0, ncols = 0;
        if (item < items) {
            e = item / Hcur, hh = item - e * Hcur;
            const float4 bb = sm.bbox[e];
            const float glo = fmaxf(bb.z, -pi_f - kBinMargin) - kBinMargin, ghi = fminf(bb.w, pi_f + kBinMargin) + kBinMargin;
            if (glo <= ghi) {
                const float hf = (float)sm.head_val[hh];
                // eye-frame azimuth window [glo - h, ghi - h], as a run of grid columns modulo the circle
                const int c_first = (int)floorf((glo - hf + pi_f) * inv_dy);
                ncols = min((int)floorf((ghi - hf + pi_f) * inv_dy) - c_first + 1, Gy);
                c0 = c_first % Gy;
                if (c0 < 0) c0 += Gy;
                r0 = max((int)floorf((bb.x + (float)kHalfPi) * inv_dp), 0);
                nrows = max(min((int)floorf((bb.y + (float)kHalfPi) * inv_dp), Gp - 1) - r0 + 1, 0);
            }
        }
        const int ncell = nrows * ncols;
        // small pairs: one lane each
        if (ncell > 0 && ncell <= kBigItem) {
            const Entry ent = sm.tile[e];
            const double dist = __hiloint2double(ent.dist_hi, ent.dist_lo);
            const float hf = (float)sm.head_val[hh];
            const int hs = sm.ras.hinv[hh];
            for (int q = 0; q < ncell; ++q) {
                const int row = r0 + q / ncols;
                int c = c0 + q % ncols;
                if (c >= Gy) c -= Gy;
                raster_cell<kSmem>(rp, sm, gbest, exacts, row * Gy + c, ent, dist, e, hh, hs, hf);
            }
        }
        // large pairs (a blade right in front of the eye): the warp shares their cells
        unsigned int big = __ballot_sync(0xffffffffu, ncell > kBigItem);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const int be = __shfl_sync(0xffffffffu, e, src), bh = __shfl_sync(0xffffffffu, hh, src);
            const int br0 = __shfl_sync(0xffffffffu, r0, src), bc0 = __shfl_sync(0xffffffffu, c0, src);
            const int bnc = __shfl_sync(0xffffffffu, ncols, src), bn = __shfl_sync(0xffffffffu, ncell, src);
            const Entry ent = sm.tile[be];
            const double dist = __hiloint2double(ent.dist_hi, ent.dist_lo);
            const float hf = (float)sm.head_val[bh];
            const int hs = sm.ras.hinv[bh];
            for (int q = lane; q < bn; q += 32) {
                const int row = br0 + q / bnc;
                int c = bc0 + q % bnc;
                if (c >= Gy) c -= Gy;
                raster_cell<kSmem>(rp, sm, gbest, exacts, row * Gy + c, ent, dist, be, bh, hs, hf);
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// final pass of the rasteriser path, float outputs, one sample per ommatidium: per ray pick the
// winner from best[], shade, evaluate the sky and write the view (the only HBM traffic).
// Warp units = (block of 32 consecutive rays) x (chunk of headings); a ray's eye data are loaded
// once per unit and reused for all its headings; each store instruction covers 32 consecutive
// rays of one view.
// ------------------------------------------------------------------------------------------
template <bool kSmem>
__device__ void final_pass_fast(const RenderParams& rp, SmemLayout& sm, const unsigned int* gbest, int p, int grp,
                                int Hcur) {
    const int R = rp.R, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool perez = rp.sky.kind == ISY_SKY_PEREZ, uniform = rp.sky.kind == ISY_SKY_UNIFORM;
    const bool from_sky = (rp.flags & ISY_FLAG_INIT_FROM_SKY) && (perez || uniform);
    float* const o_rgb = reinterpret_cast<float*>(rp.out.rgb);
    float* const o_depth = reinterpret_cast<float*>(rp.out.depth);
    float* const o_Y = reinterpret_cast<float*>(rp.out.Y);
    float* const o_P = reinterpret_cast<float*>(rp.out.dop);
    float* const o_A = reinterpret_cast<float*>(rp.out.aop);
    const float nanf_ = __int_as_float(0x7fc00000);
    const int nb = (R + 31) >> 5;
    int nch = min(Hcur, max(1, (2 * kWarps + nb - 1) / nb));
    const int hper = (Hcur + nch - 1) / nch;
    nch = (Hcur + hper - 1) / hper;
    for (int u = warp; u < nb * nch; u += kWarps) {
        const int rb = u / nch, hc = u - rb * nch;
        const int r = rb * 32 + lane;
        const bool valid = r < R;
        const int rr = valid ? r : R - 1;
        const double e_pitch = rp.eye_py[2 * (size_t)rr];
        const bool ground = e_pitch >= 0;                               // world.py:90
        double4 t4v = make_double4(0, 1, 0, 1);
        double f_theta = 0;
        if (perez) {
            t4v = *reinterpret_cast<const double4*>(rp.eye_trig + 4 * (size_t)rr);
            f_theta = rp.eye_ftheta[rr];
        }
        const int pos = __ldg(rp.grid_pos_of + rr);
        const int hend = min(Hcur, (hc + 1) * hper);

        // two headings per trip: their sky evaluations are independent instruction streams (the pass is
        // bound by the latency of the fp64 chain, not by issue slots)
        for (int hh0 = hc * hper; hh0 < hend; hh0 += 2) {
            double Y[2] = {0, 0}, P[2] = {0, 0}, A[2] = {CUDART_NAN, CUDART_NAN};
            size_t view[2];
            unsigned int id[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int hh = min(hh0 + q, hend - 1);
                view[q] = (size_t)p * rp.H + grp + (size_t)hh * rp.ngroups;   // group = every ngroups-th heading
                const int hs = sm.ras.hinv[hh];
                id[q] = kSmem ? (unsigned)sm.ras.best[hs * R + pos] : gbest[(size_t)hs * R + pos];
                if (perez) {
                    const double hsn = sm.head_sin[hh], hcs = sm.head_cos[hh];
                    const double sy = t4v.z * hcs + t4v.w * hsn, cy = t4v.w * hcs - t4v.z * hsn;   // yaw_e + heading
                    const SunConst& sc = rp.sun_per_view ? rp.sun[view[q]] : sm.sun0;   // one sun: shared memory
                    sky_eval<false>(rp.sky, sc, t4v.y * cy, t4v.y * sy, -t4v.x, e_pitch + kHalfPi, f_theta, Y[q], P[q], A[q],
                                    false, sm.gtab);
                } else if (uniform) {
                    Y[q] = rp.sky.luminance;
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
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
                if (valid) {
                    const size_t i = view[q] * (size_t)R + r;
                    if (o_rgb) {
                        float* qq = o_rgb + 3 * i;
                        qq[0] = cr, qq[1] = cg, qq[2] = cb;
                    }
                    if (rp.out.hit) rp.out.hit[i] = hit;
                    if (o_depth) o_depth[i] = depth;
                    if (o_Y) o_Y[i] = (float)Y[q];
                    if (o_P) o_P[i] = (float)P[q];
                    if (o_A) o_A[i] = (float)A[q];
                }
            }
        }
    }
}

enum { kModeBrute = 0, kModeBins = 1, kModeRaster = 2 };

template <bool kF64>
__global__ void __launch_bounds__(kThreads, 1) render_kernel(const RenderParams rp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
    __shared__ int s_count, s_nent, s_band[2];
    __shared__ unsigned int s_item, s_warp[kWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* vis = rp.slab_vis + (size_t)blockIdx.x * rp.slab_cap;
    Entry* entries = rp.slab_entry + (size_t)blockIdx.x * rp.slab_cap;
    Exact* exacts = rp.slab_exact + (size_t)blockIdx.x * rp.slab_cap;
    unsigned int* gbest = rp.slab_best ? rp.slab_best + (size_t)blockIdx.x * rp.R : nullptr;
    const bool want_sky = rp.sky.kind != ISY_SKY_NONE;
    const bool perez = rp.sky.kind == ISY_SKY_PEREZ;
    const bool yaw_only = !rp.vquat;
    const bool raster_ok = yaw_only && rp.grid_sorted && !(rp.flags & ISY_FLAG_BRUTE_FORCE);
    const bool smem_grid = raster_ok && !gbest;
    bool grid_loaded = false;

    if (perez) {
        for (int i = tid; i < 2 * (kGTab + 1); i += kThreads) sm.gtab[i] = rp.sky_gtab[i];
        if (tid == 0) sm.sun0 = rp.sun[0];
    }
    long long t_phase[4] = {0, 0, 0, 0};   // project, accel (bins / raster), final ray pass, positions
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(rp.queue, 1u);
        __syncthreads();
        const unsigned item = s_item;
        if (item >= (unsigned)rp.items) break;
        const int p = item / rp.splits, split = item % rp.splits;
        long long t0 = clock64();

        const int E = cull_and_project(rp, p, vis, entries, exacts, sm, &s_count, &s_nent, s_band);
        long long t1 = clock64();
        int mode = kModeBrute;
        BinGrid grid;
        if (!(rp.flags & ISY_FLAG_BRUTE_FORCE) && E <= kMaxE) {
            if (raster_ok) {
                mode = kModeRaster;
                if (smem_grid && !grid_loaded) {   // the bins path shares this memory: (re)load when needed
                    for (int i = tid; i <= kPitchLut; i += kThreads) sm.ras.plut[i] = rp.grid_plut[i];
                    for (int i = tid; i < rp.R; i += kThreads) {
                        sm.ras.sorted[i] = reinterpret_cast<const float2*>(rp.grid_sorted)[i];
                        sm.ras.cell_list[i] = (unsigned short)rp.cells_list[i];
                    }
                    for (int i = tid; i <= rp.cells_rows * rp.cells_cols; i += kThreads) sm.ras.cell_start[i] = rp.cells_start[i];
                    grid_loaded = true;
                }
            } else if (build_bins(sm, grid, E, s_band, s_warp)) {
                mode = kModeBins;
                grid_loaded = false;
            }
        }
        long long t2 = clock64();

        // headings are handled in groups of rp.Hg (the rasteriser's winner buffer holds one group)
        for (int grp = split; grp < rp.ngroups; grp += rp.splits) {
            // group grp = headings grp, grp + ngroups, ...: an evenly spaced scan stays evenly spaced
            const int Hcur = (rp.H - grp + rp.ngroups - 1) / rp.ngroups;
            long long t3 = clock64();
            __syncthreads();
            if (yaw_only && tid < Hcur) {
                const double hw = wrap_yaw(rp.yaw ? rp.yaw[(size_t)p * rp.H + grp + (size_t)tid * rp.ngroups] : 0.0);
                sm.head_val[tid] = hw;
                sincos(hw, &sm.head_sin[tid], &sm.head_cos[tid]);
            }
            if (mode == kModeRaster) {
                if (smem_grid) {
                    unsigned int* b32 = reinterpret_cast<unsigned int*>(sm.ras.best);
                    for (int i = tid; i < (Hcur * rp.R + 1) / 2; i += kThreads) b32[i] = 0xffffffffu;
                } else {
                    for (int i = tid; i < Hcur * rp.R; i += kThreads) gbest[i] = kNoCopy;
                }
                __syncthreads();
                if (tid == 0) s_count = 0;     // (the cull's counter is free again: next copy to rasterise)
                sort_headings(sm, Hcur);
                __syncthreads();
                const bool uniform = sm.ras.uni[3] != 0.f;
                if (Hcur <= kSmallH && smem_grid) {
                    // (large eyes keep the pitch-run walk: their cells hold many rays and the lanes of a run stay full)
                    raster_cells<true>(rp, sm, gbest, exacts, E, Hcur, &s_count);
                } else if (smem_grid) {
                    if (uniform) raster_copies<true, true>(rp, sm, gbest, exacts, E, Hcur, &s_count);
                    else raster_copies<true, false>(rp, sm, gbest, exacts, E, Hcur, &s_count);
                } else {
                    if (uniform) raster_copies<false, true>(rp, sm, gbest, exacts, E, Hcur, &s_count);
                    else raster_copies<false, false>(rp, sm, gbest, exacts, E, Hcur, &s_count);
                }
            }
            __syncthreads();
            long long t4 = clock64();
            t_phase[1] += t4 - t3;

            if (mode == kModeRaster && !kF64 && rp.S == 1 && !rp.init_c) {
                if (smem_grid) final_pass_fast<true>(rp, sm, gbest, p, grp, Hcur);
                else final_pass_fast<false>(rp, sm, gbest, p, grp, Hcur);
                t_phase[2] += clock64() - t4;
                continue;
            }
            // ---- generic final pass: warp units = (block of 32 consecutive rays) x (chunk of headings);
            //      the eye data of a ray are loaded once per unit and reused for its headings
            const int nb = (rp.R + 31) >> 5;
            int nch = min(Hcur, max(1, (2 * kWarps + nb - 1) / nb));
            const int hper = (Hcur + nch - 1) / nch;
            nch = (Hcur + hper - 1) / hper;
            for (int u = warp; u < nb * nch; u += kWarps) {
                const int rb = u / nch, hc = u - rb * nch;
                const int r = rb * 32 + lane;
                const bool valid = r < rp.R;
                const int rr = valid ? r : rp.R - 1;
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
                    if (mode == kModeRaster) pos = __ldg(rp.grid_pos_of + rr);
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

                    if (mode == kModeRaster) {
                        const int hs = sm.ras.hinv[hh];
                        const unsigned int rk = smem_grid ? (unsigned)sm.ras.best[hs * rp.R + pos]
                                                          : gbest[(size_t)hs * rp.R + pos];
                        if (rk != kNoCopy) {
                            const Entry& w = sm.tile[rk];
                            ray.hit = w.tri;
                            ray.best = __hiloint2double(w.dist_hi, w.dist_lo);
                        }
                    } else if (mode == kModeBins) {
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
                        for (int j = 0; j < E; ++j) test_entry(ray, &entries[j], &exacts[j]);
                    }

                    RayOut o;
                    o.Y = 0, o.P = 0, o.A = CUDART_NAN;
                    if (perez) {
                        const SunConst sc = rp.sun[rp.sun_per_view ? b : 0];
                        sky_eval<kF64>(rp.sky, sc, dx, dy, dz, ray.pitch + kHalfPi, f_theta, o.Y, o.P, o.A, false, sm.gtab);
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
                    write_view<kF64>(rp, b, rr, valid, ray.hit, ray.best, o);
                }
            }
            t_phase[2] += clock64() - t4;
        }
        t_phase[0] += t1 - t0, t_phase[1] += t2 - t1, t_phase[3] += 1;
    }
    if (tid == 0) {
        for (int k = 0; k < 4; ++k) atomicAdd(&rp.phase_cycles[k], (unsigned long long)t_phase[k]);
    }
}

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

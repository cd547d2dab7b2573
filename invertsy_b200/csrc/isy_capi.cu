// C ABI of the sm_100a view renderer (include/isy_b200.h). Host-side plumbing only:
// handle management, workspace carving, launches, and the host-buffer variants.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "isy_kernels.cuh"
#include "isy_fam.cuh"
#include "isy_willshaw.cuh"

using namespace isy;

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                    \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return fail(ISY_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

}  // namespace

struct isy_world {
    int device;
    int64_t T;
    double horizon;
    float* d_planes;  // 9 planes of T floats
    float* d_rgb;     // T*3
    // xy cell grid (cell = horizon / 4): triangles filed under the cell of each of their vertices
    unsigned int* d_cell_start;
    unsigned int* d_cell_tris;
    float gx0, gy0, ginv;
    int gnx, gny;
    float* d_aos;     // T x 12 floats, one triangle after the other (x0 y0 z0 x1 | y1 z1 x2 y2 | z2 - - -): the cull
                      // gathers candidates by index, three 16-byte loads each instead of nine 4-byte ones
};

struct isy_eye {
    int device;
    int64_t N;
    int32_t S;
    int64_t R;
    double* d_quat;  // R*4
    double* d_py;    // R*2
    double* d_trig;  // R*4
    float* d_w;      // R
    // rays sorted by pitch for the rasteriser
    unsigned int* d_plut;        // kPitchLut + 1
    void* d_sorted;              // R x float2 (pitch, yaw), pitch order
    int* d_pos_of;               // R: sorted position of ray r
    int* d_sorted_ray;           // R: ray at sorted position j
    // static (pitch, yaw) ray grid for the few-headings rasteriser
    int cells_rows, cells_cols;
    unsigned int* d_cells_start; // rows*cols + 1
    unsigned int* d_cells_list;  // R: pitch-sorted positions, cell by cell
};

extern "C" {

int isy_abi_version(void) { return ISY_ABI_VERSION; }
const char* isy_last_error(void) { return g_err.c_str(); }
int64_t isy_launch_count(void) { return g_launches.load(); }

int64_t isy_debug_bounds_violations(int device, int32_t* first_line) {
#ifdef ISY_DEBUG_BOUNDS
    DeviceGuard g(device);
    unsigned int n = 0, line = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return -2;
    if (cudaMemcpyFromSymbol(&n, g_isy_bounds_violations, sizeof n) != cudaSuccess) return -2;
    cudaMemcpyFromSymbol(&line, g_isy_bounds_first_line, sizeof line);
    if (first_line) *first_line = (int32_t)line;
    return (int64_t)n;
#else
    (void)device;
    if (first_line) *first_line = 0;
    return -1;      // not a bounds-checked build
#endif
}

static int world_fill(isy_world* w, const float* tri_xyz, const float* tri_rgb, int64_t T, double horizon);

int isy_world_create(const float* tri_xyz, const float* tri_rgb, int64_t T, double horizon, int device,
                     isy_world** out) {
    if (!tri_xyz || !tri_rgb || !out || T < 0 || T >= (1 << 30)) return fail(ISY_ERR_INVALID, "isy_world_create: bad argument");
    DeviceGuard g(device);
    if (!g.ok) return fail(ISY_ERR_CUDA, "isy_world_create: cannot select CUDA device %d (no CPU fallback)", device);
    isy_world* w = new isy_world{device, T, horizon, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f, 1.f, 1, 1, nullptr};
    const int rc = world_fill(w, tri_xyz, tri_rgb, T, horizon);
    if (rc != ISY_OK) {          // nothing allocated so far outlives a failed creation
        const std::string msg = g_err;
        isy_world_destroy(w);
        g_err = msg;
        return rc;
    }
    *out = w;
    return ISY_OK;
}

static int world_fill(isy_world* w, const float* tri_xyz, const float* tri_rgb, int64_t T, double horizon) {
    size_t n = (size_t)std::max<int64_t>(T, 1);
    std::vector<float> planes(9 * n);
    for (int64_t t = 0; t < T; ++t)
        for (int k = 0; k < 9; ++k) planes[(size_t)k * T + t] = tri_xyz[t * 9 + k];
    CUDA_TRY(cudaMalloc(&w->d_planes, 9 * n * sizeof(float)));
    CUDA_TRY(cudaMalloc(&w->d_rgb, 3 * n * sizeof(float)));
    CUDA_TRY(cudaMemcpy(w->d_planes, planes.data(), 9 * (size_t)T * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(w->d_rgb, tri_rgb, 3 * (size_t)T * sizeof(float), cudaMemcpyHostToDevice));
    {
        std::vector<float> aos(12 * n, 0.f);
        for (int64_t t = 0; t < T; ++t)
            for (int k = 0; k < 9; ++k) aos[(size_t)t * 12 + k] = tri_xyz[t * 9 + k];
        CUDA_TRY(cudaMalloc(&w->d_aos, 12 * n * sizeof(float)));
        CUDA_TRY(cudaMemcpy(w->d_aos, aos.data(), 12 * n * sizeof(float), cudaMemcpyHostToDevice));
    }
    // ---- xy cell grid over the vertices ----
    {
        float xmin = 0.f, xmax = 1.f, ymin = 0.f, ymax = 1.f;
        bool any = false;
        for (int64_t t = 0; t < T; ++t)
            for (int k = 0; k < 3; ++k) {
                const float x = tri_xyz[t * 9 + 3 * k], y = tri_xyz[t * 9 + 3 * k + 1];
                if (!(x == x) || !(y == y) || std::isinf(x) || std::isinf(y)) continue;
                if (!any) xmin = xmax = x, ymin = ymax = y, any = true;
                xmin = std::min(xmin, x), xmax = std::max(xmax, x), ymin = std::min(ymin, y), ymax = std::max(ymax, y);
            }
        double cell = horizon > 0 ? horizon / 4 : 1.0;
        const double span_x = std::max<double>(xmax - xmin, 1e-3), span_y = std::max<double>(ymax - ymin, 1e-3);
        while ((span_x / cell + 1) * (span_y / cell + 1) > 4.0e6) cell *= 2;       // at most ~4 M cells
        w->gx0 = xmin, w->gy0 = ymin, w->ginv = (float)(1.0 / cell);
        w->gnx = (int)(span_x / cell) + 1, w->gny = (int)(span_y / cell) + 1;
        const size_t ncell = (size_t)w->gnx * w->gny;
        std::vector<unsigned int> start(ncell + 1, 0);
        std::vector<unsigned int> cells(3 * n);     // cell of each vertex, or ~0u for "not filed"
        for (int64_t t = 0; t < T; ++t) {
            bool bad = false;
            for (int k = 0; k < 9; ++k) bad = bad || !(tri_xyz[t * 9 + k] == tri_xyz[t * 9 + k]);
            unsigned int c[3];
            for (int k = 0; k < 3; ++k) {
                const int cx = world_cell(tri_xyz[t * 9 + 3 * k], w->gx0, w->ginv, w->gnx);
                const int cy = world_cell(tri_xyz[t * 9 + 3 * k + 1], w->gy0, w->ginv, w->gny);
                c[k] = (unsigned int)cy * w->gnx + cx;
            }
            for (int k = 0; k < 3; ++k) {
                bool dup = bad;                       // NaN triangles are never visible (world.py:98, :102)
                for (int j = 0; j < k; ++j) dup = dup || c[j] == c[k];
                cells[3 * t + k] = dup ? ~0u : c[k];
                if (!dup) start[c[k] + 1]++;
            }
        }
        for (size_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
        std::vector<unsigned int> list(std::max<size_t>(start[ncell], 1));
        std::vector<unsigned int> cur(start.begin(), start.end() - 1);
        for (int64_t t = 0; t < T; ++t)
            for (int k = 0; k < 3; ++k)
                if (cells[3 * t + k] != ~0u) list[cur[cells[3 * t + k]]++] = (unsigned int)t | ((unsigned int)k << 30);
        CUDA_TRY(cudaMalloc(&w->d_cell_start, (ncell + 1) * sizeof(unsigned int)));
        CUDA_TRY(cudaMalloc(&w->d_cell_tris, list.size() * sizeof(unsigned int)));
        CUDA_TRY(cudaMemcpy(w->d_cell_start, start.data(), (ncell + 1) * sizeof(unsigned int), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(w->d_cell_tris, list.data(), list.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
    }
    return ISY_OK;
}

int isy_world_destroy(isy_world* w) {
    if (!w) return ISY_OK;
    DeviceGuard g(w->device);
    cudaFree(w->d_planes);
    cudaFree(w->d_rgb);
    cudaFree(w->d_aos);
    cudaFree(w->d_cell_start);
    cudaFree(w->d_cell_tris);
    delete w;
    return ISY_OK;
}

int64_t isy_world_triangles(const isy_world* w) { return w ? w->T : -1; }

static int eye_fill(isy_eye* e, const double* omm_quat, const float* weights, int32_t S);

int isy_eye_create(const double* omm_quat, const float* weights, int64_t N, int32_t S, int device, isy_eye** out) {
    if (!omm_quat || !out || N <= 0 || S <= 0) return fail(ISY_ERR_INVALID, "isy_eye_create: bad argument");
    if (S > 32 || (S & (S - 1))) return fail(ISY_ERR_INVALID, "isy_eye_create: samples per ommatidium must be 1,2,4,8,16 or 32 (got %d)", S);
    if (N * S > (1ll << 30)) return fail(ISY_ERR_INVALID, "isy_eye_create: too many rays");
    DeviceGuard g(device);
    if (!g.ok) return fail(ISY_ERR_CUDA, "isy_eye_create: cannot select CUDA device %d (no CPU fallback)", device);
    isy_eye* e = new isy_eye{device, N, S, N * S, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr};
    const int rc = eye_fill(e, omm_quat, weights, S);
    if (rc != ISY_OK) {
        const std::string msg = g_err;
        isy_eye_destroy(e);
        g_err = msg;
        return rc;
    }
    *out = e;
    return ISY_OK;
}

static int eye_fill(isy_eye* e, const double* omm_quat, const float* weights, int32_t S) {
    size_t R = (size_t)e->R;
    CUDA_TRY(cudaMalloc(&e->d_quat, R * 4 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&e->d_py, R * 2 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&e->d_trig, R * 4 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&e->d_w, R * sizeof(float)));
    CUDA_TRY(cudaMemcpy(e->d_quat, omm_quat, R * 4 * sizeof(double), cudaMemcpyHostToDevice));
    std::vector<float> w(R, 1.0f / (float)S);
    if (weights) std::copy(weights, weights + R, w.begin());
    CUDA_TRY(cudaMemcpy(e->d_w, w.data(), R * sizeof(float), cudaMemcpyHostToDevice));
    eye_setup_kernel<<<(unsigned)((R + 255) / 256), 256>>>(e->d_quat, e->d_py, e->d_trig, (int)R);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());

    // ---- rays sorted by pitch + look-up table pitch slot -> first sorted ray (rasteriser, yaw-only views) ----
    {
        std::vector<double> py(2 * R);
        CUDA_TRY(cudaMemcpy(py.data(), e->d_py, 2 * R * sizeof(double), cudaMemcpyDeviceToHost));
        struct SortedRay {
            float pitch, yaw;
            int32_t ray;
        };
        // Small eyes are one band. Large and cone-sampled eyes are cut into bands of kBandRays consecutive rays (whole ommatidia:
        // the cone samples of one stay together), each sorted by pitch on its own with its own look-up table:
        // the render kernel then treats a large eye as a sequence of small ones that fit shared memory.
        const size_t band = (R > (size_t)kGridRaysSmem || S > 1) ? std::min<size_t>((size_t)kBandRays, R) : R;
        const size_t nbands = (R + band - 1) / band;
        std::vector<SortedRay> sorted(R);
        for (size_t r = 0; r < R; ++r) sorted[r] = SortedRay{(float)py[2 * r], (float)py[2 * r + 1], (int32_t)r};
        for (size_t b = 0; b < nbands; ++b)
            std::stable_sort(sorted.begin() + b * band, sorted.begin() + std::min(R, (b + 1) * band),
                             [](const SortedRay& a, const SortedRay& b2) { return a.pitch < b2.pitch; });
        std::vector<int> pos_of(R), ray_of(R);
        std::vector<float> hot(2 * R);
        for (size_t j = 0; j < R; ++j) {
            pos_of[sorted[j].ray] = (int)j;
            ray_of[j] = sorted[j].ray;
            hot[2 * j] = sorted[j].pitch;
            hot[2 * j + 1] = sorted[j].yaw;
        }
        const double pi = 3.141592653589793238462643383279502884;
        std::vector<unsigned int> plut(nbands * (kPitchLut + 1));   // per band, indices local to the band
        for (size_t b = 0; b < nbands; ++b) {
            const size_t j0 = b * band, j1 = std::min(R, j0 + band);
            size_t j = j0;
            for (int sl = 0; sl <= kPitchLut; ++sl) {
                const float edge = (float)(-pi / 2 + sl * (pi / kPitchLut)) - 1e-5f;   // conservative slot start
                while (j < j1 && sorted[j].pitch < edge) ++j;
                plut[b * (kPitchLut + 1) + sl] = (unsigned int)(j - j0);
            }
            plut[b * (kPitchLut + 1) + kPitchLut] = (unsigned int)(j1 - j0);   // the end bound of every run
        }
        CUDA_TRY(cudaMalloc(&e->d_plut, plut.size() * sizeof(unsigned int)));
        CUDA_TRY(cudaMalloc(&e->d_sorted, 2 * R * sizeof(float)));
        CUDA_TRY(cudaMalloc(&e->d_pos_of, R * sizeof(int)));
        CUDA_TRY(cudaMalloc(&e->d_sorted_ray, R * sizeof(int)));
        CUDA_TRY(cudaMemcpy(e->d_sorted_ray, ray_of.data(), R * sizeof(int), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(e->d_plut, plut.data(), plut.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(e->d_sorted, hot.data(), 2 * R * sizeof(float), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(e->d_pos_of, pos_of.data(), R * sizeof(int), cudaMemcpyHostToDevice));

        // (pitch, yaw) grid with about one ray per cell; lists hold pitch-sorted positions
        int gp = 4;
        while (gp < 512 && (size_t)gp * gp < R) gp *= 2;
        if ((size_t)gp * 2 * gp > (size_t)kCellsSmem && R <= (size_t)kGridRaysSmem) gp = 32;   // shared-memory budget
        if (const char* ov = std::getenv("ISY_CELLS_GP")) gp = std::max(4, std::atoi(ov));      // (tuning experiments)
        e->cells_rows = gp, e->cells_cols = 2 * gp;
        const size_t ncell = (size_t)gp * 2 * gp;
        std::vector<unsigned int> cstart(ncell + 1, 0), cell_of(R), clist(R);
        for (size_t r = 0; r < R; ++r) {
            int row = (int)std::floor((py[2 * r] + pi / 2) * e->cells_rows / pi);
            int colc = (int)std::floor((py[2 * r + 1] + pi) * e->cells_cols / (2 * pi));
            row = std::min(std::max(row, 0), e->cells_rows - 1);
            colc = std::min(std::max(colc, 0), e->cells_cols - 1);
            cell_of[r] = (unsigned)row * e->cells_cols + colc;
            cstart[cell_of[r] + 1]++;
        }
        for (size_t c = 0; c < ncell; ++c) cstart[c + 1] += cstart[c];
        std::vector<unsigned int> cur(cstart.begin(), cstart.end() - 1);
        for (size_t r = 0; r < R; ++r) clist[cur[cell_of[r]]++] = (unsigned int)pos_of[r];
        CUDA_TRY(cudaMalloc(&e->d_cells_start, (ncell + 1) * sizeof(unsigned int)));
        CUDA_TRY(cudaMalloc(&e->d_cells_list, R * sizeof(unsigned int)));
        CUDA_TRY(cudaMemcpy(e->d_cells_start, cstart.data(), (ncell + 1) * sizeof(unsigned int), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(e->d_cells_list, clist.data(), R * sizeof(unsigned int), cudaMemcpyHostToDevice));
    }
    return ISY_OK;
}

int isy_eye_destroy(isy_eye* e) {
    if (!e) return ISY_OK;
    DeviceGuard g(e->device);
    cudaFree(e->d_quat);
    cudaFree(e->d_py);
    cudaFree(e->d_trig);
    cudaFree(e->d_w);
    cudaFree(e->d_plut);
    cudaFree(e->d_sorted);
    cudaFree(e->d_pos_of);
    cudaFree(e->d_sorted_ray);
    cudaFree(e->d_cells_start);
    cudaFree(e->d_cells_list);
    delete e;
    return ISY_OK;
}

int64_t isy_eye_rays(const isy_eye* e) { return e ? e->R : -1; }

}  // extern "C"

namespace {

struct Plan {
    int grid;
    int splits;
    int ray_splits;   // what a general-orientation call may use (the slabs are sized for it)
    int Hg, ngroups;
    int slab_cap;
    bool global_best;
    size_t off_queue, off_status, off_sun, off_ftheta, off_gtab, off_vis, off_entry, off_exact, off_best, off_skytab, off_rowbg, off_stage, total;
    int stage_cap;
    bool rows;      // room for the launch-wide row templates, sky table and patch queues (isy_shade.cuh)
};

int device_sms(int device) {
    static std::mutex mu;
    static int cache[64] = {0};
    std::lock_guard<std::mutex> lk(mu);
    if (device >= 0 && device < 64 && cache[device]) return cache[device];
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) sms = 148;
    if (device >= 0 && device < 64) cache[device] = sms;
    return sms;
}

constexpr int64_t kRowsMinP = 16;              // fewer positions / views than this do not repay the set-up launches
constexpr int64_t kRowsMinViews = 4096;        // (a route of 812 single views: 0.27 ms without, 0.29 ms with rows mode)
constexpr size_t kSkyTabMaxBytes = 16u << 20;  // ... and it should stay L2-resident

Plan make_plan(const isy_world* w, const isy_eye* e, int64_t P, int64_t H) {
    Plan pl;
    int max_ctas = device_sms(w->device);  // one persistent CTA per SM (191 KB of shared memory, 207 KB in rows mode; kThreads threads)
    if (const char* cap = getenv("ISY_MAX_CTAS")) {   // profiling aid: fewer CTAs = fewer SMs sharing L2 / HBM
        const int c = atoi(cap);
        if (c > 0) max_ctas = std::min(max_ctas, c);
    }
    // headings per group: what the rasteriser's shared-memory winner buffer holds
    // Large eyes: scans (more than kSmallH headings per group) are rendered band by band through the same
    // shared-memory winner buffer; single views and short batches walk the ray grid with a winner buffer in HBM,
    // one heading at a time.
    pl.global_best = e->R > kGridRaysSmem || e->S > 1;   // (cone-sampled eyes take the large-eye kernel whatever their size)
    const int64_t rows = pl.global_best ? kBandRays : e->R;
    int64_t hg_max = std::min<int64_t>(kBestCap / rows, kMaxH);
    int64_t splits = 1;
    if (P < max_ctas) splits = std::min<int64_t>(H, (max_ctas + std::max<int64_t>(P, 1) - 1) / std::max<int64_t>(P, 1));
    int64_t ngroups = std::max<int64_t>((H + hg_max - 1) / hg_max, splits);
    pl.Hg = (int)((H + ngroups - 1) / ngroups);
    if (pl.global_best && pl.Hg <= kSmallH) pl.Hg = 1;
    pl.ngroups = (int)((H + pl.Hg - 1) / pl.Hg);
    pl.splits = (int)std::max<int64_t>(1, std::min<int64_t>(splits, pl.ngroups));
    int64_t items = P * pl.splits;
    // explicit rays / general orientations, fewer work items than SMs, many rays: up to 16 CTAs share the rays of one
    // item (at least 128 blocks of 32 rays each); every one of them needs its own slab
    pl.ray_splits = 1;
    if (items < max_ctas / 2 && e->R >= 8192)
        pl.ray_splits = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(16, max_ctas / std::max<int64_t>(items, 1)), e->R / 4096));
    items *= pl.ray_splits;
    pl.grid = (int)std::max<int64_t>(1, std::min<int64_t>(items, max_ctas));
    pl.slab_cap = (int)std::min<int64_t>(2 * std::max<int64_t>(w->T, 1), 1 << 17);
    size_t off = 0;
    pl.off_queue = off, off += 256;
    pl.off_status = off, off += 256;
    pl.off_sun = off, off += align_up((size_t)std::max<int64_t>(P * H, 1) * sizeof(SunConst));
    pl.off_ftheta = off, off += align_up((size_t)e->R * sizeof(double));
    pl.off_gtab = off, off += align_up((size_t)2 * (kGTab + 1) * sizeof(double2));
    pl.off_vis = off, off += align_up((size_t)pl.grid * pl.slab_cap * sizeof(int));
    pl.off_entry = off, off += align_up((size_t)pl.grid * pl.slab_cap * sizeof(Entry));
    pl.off_exact = off, off += align_up((size_t)pl.grid * pl.slab_cap * sizeof(Exact));
    pl.off_best = off, off += pl.global_best ? align_up((size_t)pl.grid * e->R * sizeof(unsigned int)) : 0;
    // launch-wide row templates of small eyes: sky table [4][H][R], template rows [6R]
    pl.rows = !pl.global_best && P >= kRowsMinP && P * H >= kRowsMinViews && (size_t)H * e->R * 3 * sizeof(float) <= kSkyTabMaxBytes;
    pl.off_skytab = off, off += pl.rows ? align_up((size_t)H * e->R * 4 * sizeof(float)) : 0;   // Y, DoP, AoP, response
    pl.off_rowbg = off, off += pl.rows ? align_up((size_t)e->R * 6 * sizeof(float)) : 0;        // hit, depth, rgb, response
    pl.stage_cap = std::min(pl.slab_cap, 4096);   // vertices of a position culled ahead (9 planes per CTA)
    pl.off_stage = off, off += (ISY_CULL_AHEAD && pl.rows) ? align_up((size_t)pl.grid * 9 * pl.stage_cap * sizeof(float)) : 0;
    pl.total = off;
    return pl;
}

int launch_render(const isy_world* w, const isy_eye* e, int64_t P, int64_t H, const double* pos, const double* yaw,
                  const double* vquat, const double* sun, const double* init_c, const isy_sky_params* sky,
                  const isy_sensor_params* sensor, uint32_t flags, const isy_outputs* out, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
    if (!w || !e || !pos || !out || P < 0 || H <= 0) return fail(ISY_ERR_INVALID, "isy_render: bad argument");
    if (w->device != e->device) return fail(ISY_ERR_INVALID, "isy_render: world and eye live on different devices");
    if (P * H * e->R >= (1ll << 31) * 1024) return fail(ISY_ERR_INVALID, "isy_render: batch too large");
    if (P == 0) return ISY_OK;
    isy_sky_params sp;
    memset(&sp, 0, sizeof sp);
    if (sky) sp = *sky;
    if (sp.kind == ISY_SKY_PEREZ && !sun) return fail(ISY_ERR_INVALID, "isy_render: Perez sky needs a sun position");
    if ((out->Y || out->dop || out->aop) && sp.kind == ISY_SKY_NONE)
        return fail(ISY_ERR_INVALID, "isy_render: sky outputs requested without a sky");
    if (out->resp && (!sensor || (sensor->channels != 0 && sensor->channels != 1 && sensor->channels != 3)))
        return fail(ISY_ERR_INVALID, "isy_render: photoreceptor responses need isy_sensor_params with 3, 1 or 0 channels");
    Plan pl = make_plan(w, e, P, H);
    if (!ws || ws_bytes < pl.total)
        return fail(ISY_ERR_WORKSPACE, "isy_render: workspace %zu B < required %zu B", ws_bytes, pl.total);

    unsigned char* base = (unsigned char*)ws;
    RenderParams rp;
    memset(&rp, 0, sizeof rp);
    rp.tri_planes = w->d_planes, rp.tri_aos = reinterpret_cast<const float4*>(w->d_aos), rp.tri_rgb = w->d_rgb;
    rp.T = (int)w->T, rp.horizon = w->horizon;
    {   // sqrt is monotone and correctly rounded on both sides: {s : sqrt(s) <= horizon} = [0, horizon_sq_max]
        double s2 = w->horizon * w->horizon;
        if (w->horizon >= 0 && std::isfinite(s2)) {
            while (std::sqrt(s2) <= w->horizon) s2 = std::nextafter(s2, INFINITY);
            while (s2 > 0 && std::sqrt(s2) > w->horizon) s2 = std::nextafter(s2, -INFINITY);
        } else {
            s2 = w->horizon >= 0 ? INFINITY : -1.0;   // negative horizon: nothing is visible
        }
        rp.horizon_sq_max = s2;
    }
    rp.wg_cell_start = w->d_cell_start, rp.wg_cell_tris = w->d_cell_tris;
    rp.wg_x0 = w->gx0, rp.wg_y0 = w->gy0, rp.wg_inv_cell = w->ginv, rp.wg_nx = w->gnx, rp.wg_ny = w->gny;
    rp.eye_py = e->d_py, rp.eye_quat = e->d_quat, rp.eye_w = e->d_w, rp.eye_trig = e->d_trig;
    rp.eye_ftheta = (const double*)(base + pl.off_ftheta);
    rp.sky_gtab = (const double2*)(base + pl.off_gtab);
    rp.N = (int)e->N, rp.S = e->S, rp.R = (int)e->R;
    rp.P = (int)P, rp.H = (int)H;
    rp.pos = pos, rp.yaw = yaw, rp.vquat = vquat, rp.init_c = init_c;
    rp.sun = (const SunConst*)(base + pl.off_sun);
    rp.sun_per_view = (flags & ISY_FLAG_SUN_PER_VIEW) ? 1 : 0;
    rp.sky = sp, rp.flags = flags, rp.out = *out;
    if (out->resp) rp.sensor = *sensor;
    rp.queue = (unsigned int*)(base + pl.off_queue);
    rp.status = (int*)(base + pl.off_status);
    rp.phase_cycles = (unsigned long long*)(base + pl.off_status + 64);
    rp.slab_vis = (int*)(base + pl.off_vis);
    rp.slab_entry = (Entry*)(base + pl.off_entry);
    rp.slab_exact = (Exact*)(base + pl.off_exact);
    rp.slab_best = pl.global_best ? (unsigned int*)(base + pl.off_best) : nullptr;
    const bool any_eye = vquat || !e->d_sorted;     // (the only kernel whose final pass can share a view's rays)
    rp.ray_splits = any_eye ? pl.ray_splits : 1;
    rp.slab_cap = pl.slab_cap, rp.splits = pl.splits, rp.items = (int)(P * pl.splits * rp.ray_splits);
    rp.Hg = pl.Hg, rp.ngroups = pl.ngroups;
    {   // The fast final pass deals (block of 32 rays) x (chunk of headings) units to the warps round-robin; it is
        // throughput-bound, so pick the chunking whose busiest warp has the least to do (ties: the coarsest, which
        // reuses a ray's data most).
        const int64_t rays = pl.global_best ? std::min<int64_t>(e->R, kBandRays) : e->R;
        const int nb = (int)((rays + 31) / 32), warps = kThreads / 32, su = pl.Hg > 1 ? 2 : 1;
        int best_nch = 1;
        int64_t best_load = INT64_MAX;
        for (int nch = 1; nch <= pl.Hg; ++nch) {
            int hper = (pl.Hg + nch - 1) / nch;
            hper = std::min(pl.Hg, (hper + su - 1) / su * su);
            const int n = (pl.Hg + hper - 1) / hper;
            const int64_t load = (int64_t)((nb * n + warps - 1) / warps) * hper;
            if (load < best_load) best_load = load, best_nch = nch;
        }
        rp.shade_nch = best_nch;
    }
    rp.grid_plut = e->d_plut, rp.grid_sorted = e->d_sorted, rp.grid_pos_of = e->d_pos_of;
    rp.grid_sorted_ray = e->d_sorted_ray;
    rp.cells_start = e->d_cells_start, rp.cells_list = e->d_cells_list;
    rp.cells_rows = e->cells_rows, rp.cells_cols = e->cells_cols;

    // work counter, heading flag and phase cycles start from zero in every launch; the status word is sticky: it is
    // cleared when it is read (check_status), or when the workspace is seen for the first time
    ws_reset_kernel<<<1, 32, 0, st>>>(reinterpret_cast<unsigned int*>(base));
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    if (sp.kind == ISY_SKY_PEREZ) {
        int n = rp.sun_per_view ? (int)(P * H) : 1;
        sun_setup_kernel<<<(n + 127) / 128, 128, 0, st>>>(sun, n, sp, (SunConst*)(base + pl.off_sun));
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        sky_table_kernel<<<(2 * (kGTab + 1) + 127) / 128, 128, 0, st>>>(sp.D, (double2*)(base + pl.off_gtab));
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        if (!vquat) {
            eye_sky_kernel<<<(unsigned)((e->R + 255) / 256), 256, 0, st>>>(e->d_py, e->d_trig, (int)e->R, sp,
                                                                         (double*)(base + pl.off_ftheta));
            g_launches++;
            CUDA_TRY(cudaGetLastError());
        }
    }
    // sky-only call: nothing to render but the element-wise sky (isy_setup.cuh, sky_views_kernel)
    if (w->T == 0 && !vquat && e->S == 1 && sp.kind != ISY_SKY_NONE && !(flags & ISY_FLAG_OUT_F64) && !out->rgb && !out->hit &&
        !out->depth && !out->resp && (out->Y || out->dop || out->aop)) {
        const long long B = (long long)P * H;
        const int grid = (int)std::min<long long>(B, 8ll * device_sms(w->device));
        sky_views_kernel<<<grid, 256, 0, st>>>(yaw, B, (int)e->R, e->d_py, e->d_trig, (const double*)(base + pl.off_ftheta), sp,
                                                (const SunConst*)(base + pl.off_sun), rp.sun_per_view,
                                                (const double2*)(base + pl.off_gtab), (float*)out->Y, (float*)out->dop,
                                                (float*)out->aop);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        return ISY_OK;
    }
    const int eye_kind0 = (vquat || !e->d_sorted) ? kEyeAny : (pl.global_best ? kEyeLarge : kEyeSmall);
    if (pl.rows && eye_kind0 == kEyeSmall && !init_c && e->S == 1 && !(sp.kind == ISY_SKY_PEREZ && rp.sun_per_view) &&
        !(flags & (ISY_FLAG_OUT_F64 | ISY_FLAG_INIT_FROM_SKY | ISY_FLAG_BRUTE_FORCE | ISY_FLAG_PER_VIEW_SKY)) &&
        !(out->resp && sensor->channels != 1)) {
        // Rows mode (isy_shade.cuh): template rows are streamed into the views while the copies are rasterised and only
        // the rays that hit something are patched afterwards.  Under a Perez sky that needs every position to look
        // along the same row of headings (grids, heat maps): checked here, on the device, in every such call; the
        // sky is then evaluated once per (heading, ray) for the whole launch.
        int* differs = (int*)(base + pl.off_queue + 128);
        float* tab = (float*)(base + pl.off_skytab);
        float* bg = (float*)(base + pl.off_rowbg);
        const bool want_sky = out->Y || out->dop || out->aop;
        const int want_resp = out->resp ? 1 : 0;
        if (sp.kind == ISY_SKY_PEREZ && (want_sky || want_resp)) {
            if (yaw) {
                const long long n = (long long)P * H;
                const int blocks = (int)std::min<long long>((n + 255) / 256, 4 * device_sms(w->device));
                heading_rows_kernel<<<blocks, 256, 0, st>>>(yaw, n, (int)H, differs);
                g_launches++;
                CUDA_TRY(cudaGetLastError());
            }
            const int n = (int)(H * e->R);
            sky_shared_kernel<<<(n + 127) / 128, 128, 0, st>>>(yaw, (int)H, (int)e->R, e->d_py, e->d_trig,
                                                              (const double*)(base + pl.off_ftheta), sp,
                                                              (const SunConst*)(base + pl.off_sun),
                                                              (const double2*)(base + pl.off_gtab), differs, tab, want_resp,
                                                              rp.sensor);
            g_launches++;
            CUDA_TRY(cudaGetLastError());
            rp.sky_rows_h = (int)H;
        } else if (sp.kind == ISY_SKY_UNIFORM && want_sky) {
            rp.sky_rows_h = 1;
        }
        row_templates_kernel<<<(unsigned)((e->R + 127) / 128), 128, 0, st>>>(e->d_py, (int)e->R, sp.kind == ISY_SKY_UNIFORM && want_sky,
                                                                            sp.luminance, bg, tab, want_resp,
                                                                            sp.kind != ISY_SKY_NONE, rp.sensor);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        rp.sky_tab = tab, rp.sky_tab_off = differs, rp.row_bg = bg, rp.resp_bg = bg + 5 * (size_t)e->R;
        {   // 16 bytes per lane needs rows that start on 16-byte boundaries in every buffer that is written
            const void* ptrs[7] = {out->rgb, out->hit, out->depth, out->Y, out->dop, out->aop, out->resp};
            bool aligned = (e->R & 3) == 0;
            for (const void* q : ptrs) aligned = aligned && ((uintptr_t)q & 15u) == 0;
            rp.rows_vec4 = aligned ? 1 : 0;
        }
        rp.slab_stage = ISY_CULL_AHEAD ? (float*)(base + pl.off_stage) : nullptr, rp.stage_cap = pl.stage_cap;
    }
    const int eye_kind = (vquat || !e->d_sorted) ? kEyeAny : (pl.global_best ? kEyeLarge : kEyeSmall);
    // (rows mode adds the tail of the layout; launches whose heading groups are scans add the scan tables before it)
    const size_t smem = rp.row_bg ? sizeof(SmemLayout) : ((eye_kind != kEyeAny && rp.Hg > kSmallH) ? kSmemScan : kSmemBase);
    const bool f64 = (flags & ISY_FLAG_OUT_F64) != 0;
#define ISY_LAUNCH(F64, EYE)                                                                                              \
    do {                                                                                                                  \
        CUDA_TRY(cudaFuncSetAttribute(render_kernel<F64, EYE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        render_kernel<F64, EYE><<<std::min(pl.grid, std::max(rp.items, 1)), kThreads, smem, st>>>(rp);                                                     \
    } while (0)
    if (f64) {
        if (eye_kind == kEyeAny) ISY_LAUNCH(true, kEyeAny);
        else if (eye_kind == kEyeSmall) ISY_LAUNCH(true, kEyeSmall);
        else ISY_LAUNCH(true, kEyeLarge);
    } else {
        if (eye_kind == kEyeAny) ISY_LAUNCH(false, kEyeAny);
        else if (eye_kind == kEyeSmall) ISY_LAUNCH(false, kEyeSmall);
        else ISY_LAUNCH(false, kEyeLarge);
    }
#undef ISY_LAUNCH
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return ISY_OK;
}

// read back the sticky status word of a workspace (synchronises the stream)
int check_status(void* ws, cudaStream_t st) {
    int status = 0;
    CUDA_TRY(cudaMemcpyAsync(&status, (unsigned char*)ws + 256, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemsetAsync((unsigned char*)ws + 256, 0, sizeof(int), st));     // read = cleared
    CUDA_TRY(cudaStreamSynchronize(st));
    if (status == ISY_ERR_OVERFLOW)
        return fail(ISY_ERR_OVERFLOW, "isy_render: a position sees more triangle copies than the workspace slab holds");
    return status ? fail(status, "isy_render: device-side error %d", status) : ISY_OK;
}

template <typename T>
struct DevBuf {
    T* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
};

}  // namespace

extern "C" {

size_t isy_render_workspace_bytes(const isy_world* w, const isy_eye* e, int64_t P, int64_t H) {
    if (!w || !e || P < 0 || H <= 0) return 0;
    DeviceGuard g(w->device);
    return make_plan(w, e, P, H).total;
}

int isy_render(const isy_world* w, const isy_eye* e, int64_t P, int64_t H, const double* pos_dev,
               const double* yaw_dev, const double* view_quat_dev, const double* sun_dev, const isy_sky_params* sky,
               const isy_sensor_params* sensor, uint32_t flags, const isy_outputs* out_dev, void* workspace_dev,
               size_t workspace_bytes, void* stream) {
    if (!w) return fail(ISY_ERR_INVALID, "isy_render: null world");
    DeviceGuard g(w->device);
    if (!g.ok) return fail(ISY_ERR_CUDA, "isy_render: cannot select CUDA device %d (no CPU fallback)", w->device);
    return launch_render(w, e, P, H, pos_dev, yaw_dev, view_quat_dev, sun_dev, nullptr, sky, sensor, flags, out_dev,
                         workspace_dev, workspace_bytes, (cudaStream_t)stream);
}

int isy_workspace_status(void* workspace_dev, void* stream) { return check_status(workspace_dev, (cudaStream_t)stream); }

int isy_workspace_phase_cycles(void* workspace_dev, void* stream, uint64_t* out4) {
    if (!workspace_dev || !out4) return fail(ISY_ERR_INVALID, "isy_workspace_phase_cycles: bad argument");
    CUDA_TRY(cudaMemcpyAsync(out4, (unsigned char*)workspace_dev + 256 + 64, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return ISY_OK;
}

}  // extern "C"

namespace {

struct CallCtx;
int call_ctx(int device, size_t dev_need, size_t pin_need, CallCtx** out);
std::mutex& call_mutex(int device);
int render_host_small(const isy_world* w, const isy_eye* e, int64_t P, int64_t H, const double* pos_host,
                      const double* yaw_host, const double* view_quat_host, const double* sun_host,
                      const isy_sky_params* sky, const isy_sensor_params* sensor, uint32_t flags,
                      const isy_outputs* out_host, const size_t* per_view, int n_out);

// Persistent staging of the host-buffer path: two compute/copy slots (device arenas that only
// grow), two streams and their events, one per device.  Calls on one device are serialised.
struct HostCtx {
    std::mutex mu;
    bool init = false;
    cudaStream_t s_comp = nullptr, s_copy = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    void* arena[2] = {nullptr, nullptr};
    size_t arena_bytes[2] = {0, 0};
    void* in_arena = nullptr;
    size_t in_bytes = 0;
};
HostCtx g_host_ctx[16];

cudaError_t grow(void** p, size_t* have, size_t need) {
    if (*have >= need) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr, *have = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e == cudaSuccess) *have = need;
    return e;
}

}  // namespace

extern "C" {

int isy_render_host(const isy_world* w, const isy_eye* e, int64_t P, int64_t H, const double* pos_host,
                    const double* yaw_host, const double* view_quat_host, const double* sun_host,
                    const isy_sky_params* sky, const isy_sensor_params* sensor, uint32_t flags,
                    const isy_outputs* out_host) {
    if (out_host && out_host->resp && (!sensor || (sensor->channels != 0 && sensor->channels != 1 && sensor->channels != 3)))
        return fail(ISY_ERR_INVALID, "isy_render_host: photoreceptor responses need isy_sensor_params with 3, 1 or 0 channels");
    if (!w || !e || !pos_host || !out_host || P < 0 || H <= 0) return fail(ISY_ERR_INVALID, "isy_render_host: bad argument");
    if (w->device < 0 || w->device >= 16) return fail(ISY_ERR_INVALID, "isy_render_host: device index out of range");
    DeviceGuard g(w->device);
    if (!g.ok) return fail(ISY_ERR_CUDA, "isy_render_host: cannot select CUDA device %d (no CPU fallback)", w->device);
    if (P == 0) return ISY_OK;
    HostCtx& cx = g_host_ctx[w->device];
    std::lock_guard<std::mutex> lk(cx.mu);
    if (!cx.init) {
        CUDA_TRY(cudaStreamCreateWithFlags(&cx.s_comp, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&cx.s_copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&cx.done[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&cx.copied[i], cudaEventDisableTiming));
        }
        cx.init = true;
    }
    const size_t esz = (flags & ISY_FLAG_OUT_F64) ? 8 : 4;
    const int64_t R = e->R, N = e->N;
    const bool sun_pv = flags & ISY_FLAG_SUN_PER_VIEW;

    // positions are processed in chunks so that the compute of chunk c+1 overlaps the D2H of chunk c
    constexpr int kOut = 7;
    const size_t per_view[kOut] = {out_host->rgb ? 3 * (size_t)N * esz : 0, out_host->hit ? (size_t)R * 4 : 0,
                                   out_host->depth ? (size_t)R * esz : 0, out_host->Y ? (size_t)N * esz : 0,
                                   out_host->dop ? (size_t)N * esz : 0, out_host->aop ? (size_t)N * esz : 0,
                                   out_host->resp ? (size_t)N * std::max(sensor->channels, 1) * esz : 0};
    size_t bytes_per_pos = 0;
    for (size_t b : per_view) bytes_per_pos += (size_t)H * b;
    // a few views (the closed-loop step of a simulation: one view per call): one packed transfer each way
    if ((size_t)P * bytes_per_pos <= (2u << 20) && P * H <= 256)
        return render_host_small(w, e, P, H, pos_host, yaw_host, view_quat_host, sun_host, sky, sensor, flags, out_host, per_view, kOut);
    int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(P, (int64_t)((96ull << 20) / std::max<size_t>(bytes_per_pos, 1))));
    if (chunk < P) chunk = std::max<int64_t>(chunk, std::min<int64_t>(P, 2 * device_sms(w->device)));
    const int nslots = chunk < P ? 2 : 1;

    // inputs: one small H2D each
    const size_t n_pos = (size_t)P * 3, n_yaw = yaw_host ? (size_t)P * H : 0, n_vq = view_quat_host ? (size_t)P * H * 4 : 0;
    const size_t n_sun = sun_host ? (sun_pv ? (size_t)P * H * 2 : 2) : 0;
    CUDA_TRY(grow(&cx.in_arena, &cx.in_bytes, (n_pos + n_yaw + n_vq + n_sun + 8) * sizeof(double)));
    double* d_pos = (double*)cx.in_arena;
    double* d_yaw = n_yaw ? d_pos + n_pos : nullptr;
    double* d_vq = n_vq ? d_pos + n_pos + n_yaw : nullptr;
    double* d_sun = n_sun ? d_pos + n_pos + n_yaw + n_vq : nullptr;
    CUDA_TRY(cudaMemcpyAsync(d_pos, pos_host, n_pos * 8, cudaMemcpyHostToDevice, cx.s_comp));
    if (d_yaw) CUDA_TRY(cudaMemcpyAsync(d_yaw, yaw_host, n_yaw * 8, cudaMemcpyHostToDevice, cx.s_comp));
    if (d_vq) CUDA_TRY(cudaMemcpyAsync(d_vq, view_quat_host, n_vq * 8, cudaMemcpyHostToDevice, cx.s_comp));
    if (d_sun) CUDA_TRY(cudaMemcpyAsync(d_sun, sun_host, n_sun * 8, cudaMemcpyHostToDevice, cx.s_comp));

    const size_t ws_bytes = align_up(isy_render_workspace_bytes(w, e, chunk, H));
    size_t off[kOut + 1];
    size_t total = ws_bytes;
    for (int k = 0; k < kOut; ++k) {
        off[k] = total;
        total += align_up((size_t)chunk * H * per_view[k]);
    }
    off[kOut] = total;
    for (int i = 0; i < nslots; ++i) CUDA_TRY(grow(&cx.arena[i], &cx.arena_bytes[i], total));

    void* const host_ptr[kOut] = {out_host->rgb, out_host->hit, out_host->depth, out_host->Y, out_host->dop, out_host->aop,
                                  out_host->resp};
    int rc = ISY_OK, ci = 0;
    for (int64_t p0 = 0; p0 < P && rc == ISY_OK; p0 += chunk, ++ci) {
        const int sl = ci % nslots;
        unsigned char* base = (unsigned char*)cx.arena[sl];
        const int64_t pc = std::min(chunk, P - p0);
        const size_t v0 = (size_t)p0 * H, V = (size_t)pc * H;
        if (ci >= nslots) CUDA_TRY(cudaStreamWaitEvent(cx.s_comp, cx.copied[sl], 0));   // slot buffers free again
        isy_outputs o;
        o.rgb = per_view[0] ? base + off[0] : nullptr;
        o.hit = per_view[1] ? (int32_t*)(base + off[1]) : nullptr;
        o.depth = per_view[2] ? base + off[2] : nullptr;
        o.Y = per_view[3] ? base + off[3] : nullptr;
        o.dop = per_view[4] ? base + off[4] : nullptr;
        o.aop = per_view[5] ? base + off[5] : nullptr;
        o.resp = per_view[6] ? base + off[6] : nullptr;
        rc = launch_render(w, e, pc, H, d_pos + 3 * p0, d_yaw ? d_yaw + v0 : nullptr, d_vq ? d_vq + 4 * v0 : nullptr,
                           d_sun ? (sun_pv ? d_sun + 2 * v0 : d_sun) : nullptr, nullptr, sky, sensor, flags, &o, base,
                           ws_bytes, cx.s_comp);
        if (rc != ISY_OK) break;
        CUDA_TRY(cudaEventRecord(cx.done[sl], cx.s_comp));
        CUDA_TRY(cudaStreamWaitEvent(cx.s_copy, cx.done[sl], 0));
        for (int k = 0; k < kOut; ++k) {
            if (!per_view[k]) continue;
            CUDA_TRY(cudaMemcpyAsync((unsigned char*)host_ptr[k] + v0 * per_view[k], base + off[k], V * per_view[k],
                                     cudaMemcpyDeviceToHost, cx.s_copy));
        }
        CUDA_TRY(cudaEventRecord(cx.copied[sl], cx.s_copy));
    }
    cudaError_t e1 = cudaStreamSynchronize(cx.s_comp), e2 = cudaStreamSynchronize(cx.s_copy);
    if (rc != ISY_OK) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess)
        return fail(ISY_ERR_CUDA, "isy_render_host: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    for (int i = 0; i < nslots && rc == ISY_OK; ++i) rc = check_status(cx.arena[i], cx.s_comp);
    return rc;
}

}  // extern "C"

namespace {

// Scratch of the two "call" entry points (one view of N explicit rays, the shape of a closed-loop
// simulation step): a device arena and a pinned host mirror that only grow, one stream, per device.
// One packed H2D, the kernels, one packed D2H per call; no allocation after warm-up.
struct CallCtx {
    std::mutex mu;
    bool init = false;
    cudaStream_t st = nullptr;
    void* dev = nullptr;
    size_t dev_bytes = 0;
    void* pin = nullptr;
    size_t pin_bytes = 0;
};
CallCtx g_call_ctx[16];
std::mutex& call_mutex(int device) { return g_call_ctx[std::min(std::max(device, 0), 15)].mu; }

cudaError_t grow_pinned(void** p, size_t* have, size_t need) {
    if (*have >= need) return cudaSuccess;
    if (*p) cudaFreeHost(*p);
    *p = nullptr, *have = 0;
    cudaError_t e = cudaMallocHost(p, need);
    if (e == cudaSuccess) *have = need;
    return e;
}

// Few views with host buffers: inputs packed into the pinned mirror -> one H2D; the kernels; outputs and the head of the
// workspace (its status word) in one contiguous device block -> one D2H; one synchronisation.  No allocation after
// warm-up.  This is what `eye(sky=, scene=)` costs per step of a closed-loop simulation (simulation.py:836-838).
int render_host_small(const isy_world* w, const isy_eye* e, int64_t P, int64_t H, const double* pos_host,
                      const double* yaw_host, const double* view_quat_host, const double* sun_host,
                      const isy_sky_params* sky, const isy_sensor_params* sensor, uint32_t flags,
                      const isy_outputs* out_host, const size_t* per_view, int n_out) {
    std::lock_guard<std::mutex> lk(call_mutex(w->device));
    const size_t B = (size_t)P * H;
    const bool sun_pv = flags & ISY_FLAG_SUN_PER_VIEW;
    const size_t n_pos = (size_t)P * 3, n_yaw = yaw_host ? B : 0, n_vq = view_quat_host ? B * 4 : 0;
    const size_t n_sun = sun_host ? (sun_pv ? B * 2 : 2) : 0;
    const size_t in_bytes = (n_pos + n_yaw + n_vq + n_sun) * sizeof(double);
    size_t off[8], out_bytes = 0;
    for (int k = 0; k < n_out; ++k) off[k] = out_bytes, out_bytes += align_up(B * per_view[k], 16);
    const size_t ws_bytes = align_up(isy_render_workspace_bytes(w, e, P, H));
    const size_t off_out = align_up(in_bytes), off_ws = off_out + align_up(out_bytes);
    CallCtx* cx = nullptr;
    int rc = call_ctx(w->device, off_ws + ws_bytes, std::max(in_bytes, align_up(out_bytes) + 512), &cx);
    if (rc != ISY_OK) return rc;
    unsigned char* dev = (unsigned char*)cx->dev;
    double* pin = (double*)cx->pin;
    memcpy(pin, pos_host, n_pos * 8);
    if (n_yaw) memcpy(pin + n_pos, yaw_host, n_yaw * 8);
    if (n_vq) memcpy(pin + n_pos + n_yaw, view_quat_host, n_vq * 8);
    if (n_sun) memcpy(pin + n_pos + n_yaw + n_vq, sun_host, n_sun * 8);
    CUDA_TRY(cudaMemcpyAsync(dev, pin, in_bytes, cudaMemcpyHostToDevice, cx->st));
    double* d_in = (double*)dev;
    unsigned char* d_out = dev + off_out;
    isy_outputs o;
    void** op[7] = {&o.rgb, (void**)&o.hit, &o.depth, &o.Y, &o.dop, &o.aop, &o.resp};
    void* const hp[7] = {out_host->rgb, out_host->hit, out_host->depth, out_host->Y, out_host->dop, out_host->aop, out_host->resp};
    for (int k = 0; k < n_out; ++k) *op[k] = per_view[k] ? d_out + off[k] : nullptr;
    // the status word of this private workspace starts clean in every call (it is read below, with the outputs)
    CUDA_TRY(cudaMemsetAsync(dev + off_ws + 256, 0, sizeof(int), cx->st));
    rc = launch_render(w, e, P, H, d_in, n_yaw ? d_in + n_pos : nullptr, n_vq ? d_in + n_pos + n_yaw : nullptr,
                       n_sun ? d_in + n_pos + n_yaw + n_vq : nullptr, nullptr, sky, sensor, flags, &o, dev + off_ws, ws_bytes, cx->st);
    if (rc != ISY_OK) return rc;
    // outputs and the first 512 bytes of the workspace are contiguous on the device
    CUDA_TRY(cudaMemcpyAsync(pin, d_out, align_up(out_bytes) + 512, cudaMemcpyDeviceToHost, cx->st));
    CUDA_TRY(cudaStreamSynchronize(cx->st));
    const int status = *(const int*)((const unsigned char*)pin + align_up(out_bytes) + 256);
    if (status == ISY_ERR_OVERFLOW)
        return fail(ISY_ERR_OVERFLOW, "isy_render: a position sees more triangle copies than the workspace slab holds");
    if (status) return fail(status, "isy_render: device-side error %d", status);
    for (int k = 0; k < n_out; ++k)
        if (per_view[k]) memcpy(hp[k], (const unsigned char*)pin + off[k], B * per_view[k]);
    return ISY_OK;
}

int call_ctx(int device, size_t dev_need, size_t pin_need, CallCtx** out) {
    if (device < 0 || device >= 16) return fail(ISY_ERR_INVALID, "device index out of range");
    CallCtx& cx = g_call_ctx[device];
    if (!cx.init) {
        CUDA_TRY(cudaStreamCreateWithFlags(&cx.st, cudaStreamNonBlocking));
        cx.init = true;
    }
    CUDA_TRY(grow(&cx.dev, &cx.dev_bytes, dev_need));
    CUDA_TRY(grow_pinned(&cx.pin, &cx.pin_bytes, pin_need));
    *out = &cx;
    return ISY_OK;
}

}  // namespace

extern "C" {

int isy_world_call_host(const isy_world* w, const double* pos3_host, const double* ray_quat_host, int64_t N,
                        const double* init_c_host, double* rgb_host, int32_t* hit_host, double* depth_host) {
    if (!w || !pos3_host || !ray_quat_host || N < 0) return fail(ISY_ERR_INVALID, "isy_world_call_host: bad argument");
    if (N == 0) return ISY_OK;
    if (N > (1ll << 28)) return fail(ISY_ERR_INVALID, "isy_world_call_host: too many rays");
    DeviceGuard g(w->device);
    if (!g.ok) return fail(ISY_ERR_CUDA, "isy_world_call_host: cannot select CUDA device %d (no CPU fallback)", w->device);
    std::lock_guard<std::mutex> lk(g_call_ctx[std::min(std::max(w->device, 0), 15)].mu);
    const size_t n = (size_t)N;
    // a transient eye of N explicit rays (no pitch-sorted table: the per-position bins path serves it)
    isy_eye eye;
    memset(&eye, 0, sizeof eye);
    eye.device = w->device, eye.N = N, eye.S = 1, eye.R = N;
    const size_t ws_bytes = align_up(isy_render_workspace_bytes(w, &eye, 1, 1));
    // device layout: [in: pos 3 | quat 4n | init n] [scratch: py 2n | trig 4n] [out: rgb 3n | depth n | hit n(i32)] [ws]
    const size_t in_d = 3 + 4 * n + (init_c_host ? n : 0);
    const size_t off_py = align_up(in_d * 8), off_trig = off_py + align_up(2 * n * 8);
    const size_t off_out = off_trig + align_up(4 * n * 8);
    const size_t out_bytes = 3 * n * 8 + n * 8 + n * 4;
    const size_t off_ws = off_out + align_up(out_bytes);
    CallCtx* cx = nullptr;
    int rc = call_ctx(w->device, off_ws + ws_bytes, std::max(in_d * 8, out_bytes), &cx);
    if (rc != ISY_OK) return rc;
    unsigned char* dev = (unsigned char*)cx->dev;
    double* pin = (double*)cx->pin;
    memcpy(pin, pos3_host, 24);
    memcpy(pin + 3, ray_quat_host, 4 * n * 8);
    if (init_c_host) memcpy(pin + 3 + 4 * n, init_c_host, n * 8);
    CUDA_TRY(cudaMemcpyAsync(dev, pin, in_d * 8, cudaMemcpyHostToDevice, cx->st));
    double* d_in = (double*)dev;
    eye.d_quat = d_in + 3;
    eye.d_py = (double*)(dev + off_py);
    eye.d_trig = (double*)(dev + off_trig);
    eye_setup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cx->st>>>(eye.d_quat, eye.d_py, eye.d_trig, (int)n);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    isy_outputs o;
    memset(&o, 0, sizeof o);
    double* d_rgb = (double*)(dev + off_out);
    double* d_depth = d_rgb + 3 * n;
    int32_t* d_hit = (int32_t*)(d_depth + n);
    if (rgb_host) o.rgb = d_rgb;
    if (depth_host) o.depth = d_depth;
    if (hit_host) o.hit = d_hit;
    rc = launch_render(w, &eye, 1, 1, d_in, nullptr, nullptr, nullptr, init_c_host ? d_in + 3 + 4 * n : nullptr, nullptr,
                       nullptr, ISY_FLAG_OUT_F64, &o, dev + off_ws, ws_bytes, cx->st);
    if (rc != ISY_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(pin, d_rgb, out_bytes, cudaMemcpyDeviceToHost, cx->st));
    rc = check_status(dev + off_ws, cx->st);   // synchronises the stream
    if (rc != ISY_OK) return rc;
    if (rgb_host) memcpy(rgb_host, pin, 3 * n * 8);
    if (depth_host) memcpy(depth_host, pin + 3 * n, n * 8);
    if (hit_host) memcpy(hit_host, pin + 4 * n, n * 4);
    return ISY_OK;
}

int isy_sky_call_host(const isy_sky_params* sky, double theta_s, double phi_s, const double* ray_quat_host, int64_t N,
                      const uint8_t* eta_host, int device, double* Y_host, double* dop_host, double* aop_host,
                      double* theta_host, double* phi_host) {
    if (!sky || !ray_quat_host || !Y_host || !dop_host || !aop_host || N < 0)
        return fail(ISY_ERR_INVALID, "isy_sky_call_host: bad argument");
    if (sky->kind != ISY_SKY_UNIFORM && sky->kind != ISY_SKY_PEREZ)
        return fail(ISY_ERR_INVALID, "isy_sky_call_host: sky kind must be UNIFORM or PEREZ");
    if (N == 0) return ISY_OK;
    DeviceGuard g(device);
    if (!g.ok) return fail(ISY_ERR_CUDA, "isy_sky_call_host: cannot select CUDA device %d (no CPU fallback)", device);
    std::lock_guard<std::mutex> lk(g_call_ctx[std::min(std::max(device, 0), 15)].mu);
    const size_t n = (size_t)N;
    // device layout: [in: sun 2 | quat 4n | eta n bytes] [SunConst] [out: Y P A theta phi, n each]
    const size_t in_bytes = (2 + 4 * n) * 8 + (eta_host ? n : 0);
    const size_t off_sc = align_up(in_bytes), off_out = off_sc + align_up(sizeof(SunConst));
    const size_t out_bytes = 5 * n * 8;
    CallCtx* cx = nullptr;
    int rc = call_ctx(device, off_out + out_bytes, std::max(in_bytes, out_bytes), &cx);
    if (rc != ISY_OK) return rc;
    unsigned char* dev = (unsigned char*)cx->dev;
    double* pin = (double*)cx->pin;
    pin[0] = theta_s, pin[1] = phi_s;
    memcpy(pin + 2, ray_quat_host, 4 * n * 8);
    if (eta_host) memcpy(pin + 2 + 4 * n, eta_host, n);
    CUDA_TRY(cudaMemcpyAsync(dev, pin, in_bytes, cudaMemcpyHostToDevice, cx->st));
    double* d_in = (double*)dev;
    SunConst* d_sc = (SunConst*)(dev + off_sc);
    if (sky->kind == ISY_SKY_PEREZ) {
        sun_setup_kernel<<<1, 32, 0, cx->st>>>(d_in, 1, *sky, d_sc);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    double *dY = (double*)(dev + off_out), *dP = dY + n, *dA = dP + n, *dT = dA + n, *dF = dT + n;
    sky_dirs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cx->st>>>(
        d_in + 2, N, *sky, d_sc, eta_host ? (const uint8_t*)(d_in + 2 + 4 * n) : nullptr, dY, dP, dA, dT, dF);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(pin, dY, out_bytes, cudaMemcpyDeviceToHost, cx->st));
    CUDA_TRY(cudaStreamSynchronize(cx->st));
    memcpy(Y_host, pin, n * 8);
    memcpy(dop_host, pin + n, n * 8);
    memcpy(aop_host, pin + 2 * n, n * 8);
    if (theta_host) memcpy(theta_host, pin + 3 * n, n * 8);
    if (phi_host) memcpy(phi_host, pin + 4 * n, n * 8);
    return ISY_OK;
}

int isy_spectrum_influence_host(const double* v_host, int64_t n, const double* irgbu_host, int64_t rows, int device,
                                double* sum_host, double* channels_host) {
    if (!v_host || !irgbu_host || (!sum_host && !channels_host) || n < 0 || rows <= 0)
        return fail(ISY_ERR_INVALID, "isy_spectrum_influence_host: bad argument");
    if (n % rows) return fail(ISY_ERR_INVALID, "isy_spectrum_influence_host: %lld values cannot be tiled by %lld irgbu rows",
                              (long long)n, (long long)rows);
    if (n == 0) return ISY_OK;
    DeviceGuard g(device);
    if (!g.ok) return fail(ISY_ERR_CUDA, "isy_spectrum_influence_host: cannot select CUDA device %d (no CPU fallback)", device);
    DevBuf<double> d_v, d_w, d_o, d_c;
    CUDA_TRY(d_v.alloc((size_t)n));
    if (sum_host) CUDA_TRY(d_o.alloc((size_t)n));
    if (channels_host) CUDA_TRY(d_c.alloc((size_t)n * 5));
    CUDA_TRY(d_w.alloc((size_t)rows * 5));
    CUDA_TRY(cudaMemcpy(d_v.p, v_host, (size_t)n * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_w.p, irgbu_host, (size_t)rows * 40, cudaMemcpyHostToDevice));
    spectrum_kernel<double><<<1, 1024>>>(d_v.p, n, d_w.p, rows, d_o.p, d_c.p);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    if (sum_host) CUDA_TRY(cudaMemcpy(sum_host, d_o.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (channels_host) CUDA_TRY(cudaMemcpy(channels_host, d_c.p, (size_t)n * 40, cudaMemcpyDeviceToHost));
    return ISY_OK;
}

}  // extern "C"


// ================================================================================================
// familiarity (isy_fam.cuh)
// ================================================================================================
struct isy_memory {
    int device;
    int64_t M, K;
    int64_t ldm;        // row stride of the device copies in floats (K rounded up to 4: TMA strides are multiples of 16 B)
    int64_t U;          // unique stored vectors (bit-identical copies are kept once)
    int64_t Mpad;       // rows of the device copies (U rounded up to the 256-row TMA box, zero rows)
    float* d_stored;    // [Mpad][ldm] the vectors as given (exact re-check)
    float* d_centred;   // [Mpad][ldm] vectors minus their mean (operand of the TF32 filter)
    float* d_cn;        // [Mpad] |m|^2
    int32_t* d_orig;    // [Mpad] index (as given to isy_memory_create) of each unique vector: the lowest of its copies
    float mc_max;       // max over the stored vectors of |m - mean|_2 (error bound of the TF32 filter)
    CUtensorMap map_b;  // d_centred, box 32 x 256, 128-byte swizzle
};

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// [rows][K] floats with row stride ld (floats), box = 32 floats x box_rows, 128-byte swizzle, zero fill
int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(ISY_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)fam::kBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ISY_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return ISY_OK;
}

int memory_fill(isy_memory* m, const float* v) {
    const int64_t M = m->M, K = m->K, ld = m->ldm;
    // Bit-identical stored vectors (an eye that saturates sees the same thing from neighbouring route steps) are kept
    // once: the filter multiplies the unique ones only, and exact ties between copies cannot crowd the four candidate
    // slots.  orig[u] = lowest index among the copies of unique vector u, which is what an arg-min over all M returns.
    std::vector<int32_t> orig;
    {
        std::vector<int64_t> order((size_t)M);
        for (int64_t i = 0; i < M; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
            return memcmp(v + a * K, v + b * K, (size_t)K * sizeof(float)) < 0;
        });
        std::vector<char> dup((size_t)M, 0);
        for (int64_t j = 1; j < M; ++j)      // stable sort: the first of a run of equals has the lowest index
            if (memcmp(v + order[j - 1] * K, v + order[j] * K, (size_t)K * sizeof(float)) == 0) dup[order[j]] = 1;
        for (int64_t i = 0; i < M; ++i)
            if (!dup[i]) orig.push_back((int32_t)i);
    }
    const int64_t U = (int64_t)orig.size();
    m->U = U;
    m->Mpad = (U + fam::kNB - 1) / fam::kNB * fam::kNB;
    const int64_t Mp = m->Mpad;
    std::vector<double> mean((size_t)K, 0.0);
    for (int64_t u = 0; u < U; ++u)
        for (int64_t k = 0; k < K; ++k) mean[k] += v[(int64_t)orig[u] * K + k];
    for (int64_t k = 0; k < K; ++k) mean[k] /= (double)U;
    std::vector<float> stored((size_t)Mp * ld, 0.f), centred((size_t)Mp * ld, 0.f), cn((size_t)Mp, 0.f);
    for (int64_t u = 0; u < U; ++u) {
        double s = 0.0, sc = 0.0;
        for (int64_t k = 0; k < K; ++k) {
            const double x = v[(int64_t)orig[u] * K + k];
            stored[(size_t)u * ld + k] = (float)x;
            centred[(size_t)u * ld + k] = (float)(x - mean[k]);
            s += x * x;
            sc += (double)centred[(size_t)u * ld + k] * centred[(size_t)u * ld + k];
        }
        cn[u] = (float)s;
        m->mc_max = std::max(m->mc_max, (float)(std::sqrt(sc) * (1.0 + 1e-6)));
    }
    orig.resize((size_t)Mp, -1);
    CUDA_TRY(cudaMalloc(&m->d_stored, stored.size() * sizeof(float)));
    CUDA_TRY(cudaMalloc(&m->d_centred, centred.size() * sizeof(float)));
    CUDA_TRY(cudaMalloc(&m->d_cn, cn.size() * sizeof(float)));
    CUDA_TRY(cudaMalloc(&m->d_orig, orig.size() * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpy(m->d_stored, stored.data(), stored.size() * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(m->d_centred, centred.data(), centred.size() * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(m->d_cn, cn.data(), cn.size() * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(m->d_orig, orig.data(), orig.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    return make_map(&m->map_b, m->d_centred, Mp, K, ld, fam::kNB);
}

}  // namespace

extern "C" {

int isy_memory_create(const float* vectors_host, int64_t M, int64_t K, int device, isy_memory** out) {
    if (!vectors_host || !out || M <= 0 || K <= 0 || M >= (1ll << 30) || K >= (1ll << 24))
        return fail(ISY_ERR_INVALID, "isy_memory_create: bad argument");
    DeviceGuard g(device);
    if (!g.ok) return fail(ISY_ERR_CUDA, "isy_memory_create: cannot select CUDA device %d (no CPU fallback)", device);
    isy_memory* m = new isy_memory;
    memset(m, 0, sizeof *m);
    m->device = device, m->M = M, m->K = K, m->ldm = (K + 3) / 4 * 4;
    const int rc = memory_fill(m, vectors_host);
    if (rc != ISY_OK) {
        const std::string msg = g_err;
        isy_memory_destroy(m);
        g_err = msg;
        return rc;
    }
    *out = m;
    return ISY_OK;
}

int isy_memory_destroy(isy_memory* m) {
    if (!m) return ISY_OK;
    DeviceGuard g(m->device);
    cudaFree(m->d_stored);
    cudaFree(m->d_centred);
    cudaFree(m->d_cn);
    cudaFree(m->d_orig);
    delete m;
    return ISY_OK;
}

int64_t isy_memory_size(const isy_memory* m) { return m ? m->M : -1; }

size_t isy_familiarity_workspace_bytes(const isy_memory* m, int64_t Q) {
    if (!m || Q < 0) return 0;
    // candidates (int4) and their filter scores (float4) per query, then the counter of unresolved queries
    return align_up((size_t)std::max<int64_t>(Q, 1) * (sizeof(int4) + sizeof(float4))) + 256;
}

int isy_familiarity(const isy_memory* m, const float* queries_dev, int64_t Q, int64_t row_stride, float* familiarity_dev,
                    int32_t* nearest_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (m && Q == 0) return ISY_OK;      // empty input (the reference's loops simply do not run)
    if (!m || !queries_dev || Q < 0 || (!familiarity_dev && !nearest_dev))
        return fail(ISY_ERR_INVALID, "isy_familiarity: bad argument");
    if (Q >= (1ll << 31) - fam::kBM) return fail(ISY_ERR_INVALID, "isy_familiarity: too many queries for one call");
    if (row_stride < m->K || (row_stride & 3) || ((uintptr_t)queries_dev & 15))
        return fail(ISY_ERR_INVALID, "isy_familiarity: queries must be 16-byte aligned with a row stride that is a multiple of 4 floats and at least K");
    if (!workspace_dev || workspace_bytes < isy_familiarity_workspace_bytes(m, Q))
        return fail(ISY_ERR_WORKSPACE, "isy_familiarity: workspace %zu B < required %zu B", workspace_bytes,
                    isy_familiarity_workspace_bytes(m, Q));
    DeviceGuard g(m->device);
    if (!g.ok) return fail(ISY_ERR_CUDA, "isy_familiarity: cannot select CUDA device %d (no CPU fallback)", m->device);
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap map_a;
    int rc = make_map(&map_a, queries_dev, Q, m->K, row_stride, fam::kBM);
    if (rc != ISY_OK) return rc;
    int4* cand = reinterpret_cast<int4*>(workspace_dev);
    float4* score = reinterpret_cast<float4*>(cand + Q);
    unsigned int* unresolved = reinterpret_cast<unsigned int*>((unsigned char*)workspace_dev + align_up((size_t)Q * (sizeof(int4) + sizeof(float4))));
    CUDA_TRY(cudaMemsetAsync(unresolved, 0, sizeof(unsigned int), st));
    const int ntiles = (int)((Q + fam::kBM - 1) / fam::kBM);
    const int grid = std::min(ntiles, device_sms(m->device));
    CUDA_TRY(cudaFuncSetAttribute(fam::fam_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fam::kSmemBytes));
    fam::fam_filter_kernel<<<grid, fam::kThreads, fam::kSmemBytes, st>>>(map_a, m->map_b, m->d_cn, (int)Q, (int)m->U, (int)m->K, cand, score);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    const int wpb = 8;
    fam::fam_recheck_kernel<<<(unsigned)((Q + wpb - 1) / wpb), wpb * 32, 0, st>>>(queries_dev, (long long)row_stride, m->d_stored, (long long)m->ldm,
                                                                                   (int)Q, (int)m->K, cand, score, m->mc_max,
                                                                                   familiarity_dev, nearest_dev, m->d_orig, unresolved);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return ISY_OK;
}

int isy_familiarity_unresolved(const isy_memory* m, int64_t Q, void* workspace_dev, void* stream, int64_t* count) {
    if (!m || !workspace_dev || !count || Q < 0) return fail(ISY_ERR_INVALID, "isy_familiarity_unresolved: bad argument");
    DeviceGuard g(m->device);
    unsigned int n = 0;
    CUDA_TRY(cudaMemcpyAsync(&n, (unsigned char*)workspace_dev + align_up((size_t)std::max<int64_t>(Q, 1) * (sizeof(int4) + sizeof(float4))),
                             sizeof n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    *count = n;
    return ISY_OK;
}

int isy_spectrum_influence_views(void* Y_dev, int64_t B, int64_t n, const double* irgbu_dev, int64_t rows, uint32_t flags,
                                 int device, void* stream) {
    if (!Y_dev || !irgbu_dev || B < 0 || n <= 0 || rows <= 0) return fail(ISY_ERR_INVALID, "isy_spectrum_influence_views: bad argument");
    if (n % rows) return fail(ISY_ERR_INVALID, "isy_spectrum_influence_views: %lld values cannot be tiled by %lld irgbu rows",
                              (long long)n, (long long)rows);
    if (B == 0) return ISY_OK;
    if (B > 0x7fffffff) return fail(ISY_ERR_INVALID, "isy_spectrum_influence_views: too many views for one call");
    DeviceGuard g(device);
    if (!g.ok) return fail(ISY_ERR_CUDA, "isy_spectrum_influence_views: cannot select CUDA device %d (no CPU fallback)", device);
    cudaStream_t st = (cudaStream_t)stream;
    if (flags & ISY_FLAG_OUT_F64)
        spectrum_kernel<double><<<(unsigned)B, 256, 0, st>>>((const double*)Y_dev, n, irgbu_dev, rows, (double*)Y_dev, (double*)nullptr);
    else
        spectrum_kernel<float><<<(unsigned)B, 256, 0, st>>>((const float*)Y_dev, n, irgbu_dev, rows, (float*)Y_dev, (float*)nullptr);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return ISY_OK;
}

int isy_familiarity_candidates(const isy_memory* m, int64_t Q, void* workspace_dev, void* stream, int32_t* cand_host,
                               float* score_host) {
    if (!m || !workspace_dev || Q < 0 || (!cand_host && !score_host))
        return fail(ISY_ERR_INVALID, "isy_familiarity_candidates: bad argument");
    DeviceGuard g(m->device);
    const int4* cand = reinterpret_cast<const int4*>(workspace_dev);
    if (cand_host) CUDA_TRY(cudaMemcpyAsync(cand_host, cand, (size_t)Q * sizeof(int4), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    if (score_host) CUDA_TRY(cudaMemcpyAsync(score_host, cand + Q, (size_t)Q * sizeof(float4), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return ISY_OK;
}

int isy_familiarity_host(const isy_memory* m, const float* queries_host, int64_t Q, float* familiarity_host,
                         int32_t* nearest_host) {
    if (!m || !queries_host || Q < 0 || (!familiarity_host && !nearest_host))
        return fail(ISY_ERR_INVALID, "isy_familiarity_host: bad argument");
    if (Q == 0) return ISY_OK;
    DeviceGuard g(m->device);
    if (!g.ok) return fail(ISY_ERR_CUDA, "isy_familiarity_host: cannot select CUDA device %d (no CPU fallback)", m->device);
    const int64_t ld = m->ldm;
    DevBuf<float> d_q, d_f;
    DevBuf<int32_t> d_n;
    DevBuf<unsigned char> d_ws;
    const size_t ws = isy_familiarity_workspace_bytes(m, Q);
    CUDA_TRY(d_q.alloc((size_t)Q * ld));
    CUDA_TRY(d_f.alloc((size_t)Q));
    CUDA_TRY(d_n.alloc((size_t)Q));
    CUDA_TRY(d_ws.alloc(ws));
    if (ld != m->K) CUDA_TRY(cudaMemset(d_q.p, 0, (size_t)Q * ld * sizeof(float)));
    CUDA_TRY(cudaMemcpy2D(d_q.p, (size_t)ld * sizeof(float), queries_host, (size_t)m->K * sizeof(float), (size_t)m->K * sizeof(float),
                          (size_t)Q, cudaMemcpyHostToDevice));
    const int rc = isy_familiarity(m, d_q.p, Q, ld, d_f.p, d_n.p, d_ws.p, ws, nullptr);
    if (rc != ISY_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(nullptr));
    if (familiarity_host) CUDA_TRY(cudaMemcpy(familiarity_host, d_f.p, (size_t)Q * sizeof(float), cudaMemcpyDeviceToHost));
    if (nearest_host) CUDA_TRY(cudaMemcpy(nearest_host, d_n.p, (size_t)Q * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return ISY_OK;
}

}  // extern "C"


// ================================================================================================
// Willshaw memory (isy_willshaw.cuh)
// ================================================================================================
struct isy_willshaw {
    int device;
    int64_t nb_pn, nb_kc;
    int32_t fan_in, k;
    uint32_t seed;
    unsigned int* d_mask;   // bit j: the KC -> MBON synapse of cell j is still 1
    size_t smem;
};

namespace {
size_t willshaw_smem(int64_t nb_pn, int64_t nb_kc) { return (size_t)((nb_pn + 3) & ~3ll) * 4 + (size_t)nb_kc * 4 + mb::kBins * 4; }
}  // namespace

extern "C" {

int isy_willshaw_create(int64_t nb_pn, int64_t nb_kc, int32_t fan_in, double sparseness, uint32_t seed, int device,
                        isy_willshaw** out) {
    if (!out || nb_pn <= 0 || nb_kc <= 0 || fan_in <= 0 || !(sparseness > 0) || sparseness > 1)
        return fail(ISY_ERR_INVALID, "isy_willshaw_create: bad argument");
    if (willshaw_smem(nb_pn, nb_kc) > 220 * 1024)
        return fail(ISY_ERR_INVALID, "isy_willshaw_create: %lld PNs + %lld KCs do not fit the shared memory of one SM (4 bytes each, 220 KB)",
                    (long long)nb_pn, (long long)nb_kc);
    DeviceGuard g(device);
    if (!g.ok) return fail(ISY_ERR_CUDA, "isy_willshaw_create: cannot select CUDA device %d (no CPU fallback)", device);
    isy_willshaw* m = new isy_willshaw{device, nb_pn, nb_kc, fan_in, 0, seed, nullptr, willshaw_smem(nb_pn, nb_kc)};
    m->k = (int32_t)std::max<int64_t>(1, std::min<int64_t>(nb_kc, (int64_t)std::llround(sparseness * (double)nb_kc)));
    if (cudaMalloc(&m->d_mask, (size_t)((nb_kc + 31) / 32) * 4) != cudaSuccess) {
        delete m;
        return fail(ISY_ERR_CUDA, "isy_willshaw_create: out of device memory");
    }
    *out = m;
    return isy_willshaw_reset(m);
}

int isy_willshaw_destroy(isy_willshaw* m) {
    if (!m) return ISY_OK;
    DeviceGuard g(m->device);
    cudaFree(m->d_mask);
    delete m;
    return ISY_OK;
}

int isy_willshaw_reset(isy_willshaw* m) {
    if (!m) return fail(ISY_ERR_INVALID, "isy_willshaw_reset: null handle");
    DeviceGuard g(m->device);
    CUDA_TRY(cudaMemset(m->d_mask, 0xff, (size_t)((m->nb_kc + 31) / 32) * 4));
    return ISY_OK;
}

int64_t isy_willshaw_active(const isy_willshaw* m) { return m ? m->k : -1; }

static int willshaw_launch(const isy_willshaw* m, bool train, const float* views_dev, int64_t Q, int64_t row_stride, float* fam_dev,
                           void* stream) {
    if (m && Q == 0) return ISY_OK;
    if (!m || !views_dev || Q < 0 || row_stride < m->nb_pn || (!train && !fam_dev))
        return fail(ISY_ERR_INVALID, "isy_willshaw: bad argument");
    if (Q > 0x7fffffff) return fail(ISY_ERR_INVALID, "isy_willshaw: too many views for one call");
    DeviceGuard g(m->device);
    if (!g.ok) return fail(ISY_ERR_CUDA, "isy_willshaw: cannot select CUDA device %d (no CPU fallback)", m->device);
    const int grid = (int)std::min<int64_t>(Q, device_sms(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (train) {
        CUDA_TRY(cudaFuncSetAttribute(mb::willshaw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem));
        mb::willshaw_kernel<true><<<grid, mb::kThreads, m->smem, st>>>(views_dev, (long long)row_stride, (int)Q, (int)m->nb_pn, (int)m->nb_kc,
                                                                       m->fan_in, m->k, m->seed, m->d_mask, nullptr);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(mb::willshaw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem));
        mb::willshaw_kernel<false><<<grid, mb::kThreads, m->smem, st>>>(views_dev, (long long)row_stride, (int)Q, (int)m->nb_pn, (int)m->nb_kc,
                                                                        m->fan_in, m->k, m->seed, m->d_mask, fam_dev);
    }
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return ISY_OK;
}

int isy_willshaw_train(isy_willshaw* m, const float* vectors_dev, int64_t M, int64_t row_stride, void* stream) {
    return willshaw_launch(m, true, vectors_dev, M, row_stride, nullptr, stream);
}

int isy_willshaw_familiarity(const isy_willshaw* m, const float* queries_dev, int64_t Q, int64_t row_stride, float* familiarity_dev,
                             void* stream) {
    return willshaw_launch(m, false, queries_dev, Q, row_stride, familiarity_dev, stream);
}

}  // extern "C"

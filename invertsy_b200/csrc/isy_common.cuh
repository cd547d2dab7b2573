// Shared definitions of the sm_100a view renderer: layouts, exact (reference-order) fp64
// helpers and the fp32 edge-function filter.  Reference = InvertSy env/world.py, env/sky.py.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/isy_b200.h"

// ------------------------------------------------------------------------------------------
// Bounds-checked build (-DISY_DEBUG_BOUNDS; `python tools/debug_bounds.py` builds it next to the product library and
// runs GPU tests through it): compute-sanitizer is not available on the B200 pool, so the indexed shared-memory and
// output accesses of the kernels carry their own asserts.  A violated bound is counted (isy_debug_bounds_violations)
// and the access is skipped by the caller where that is possible; the release build compiles the checks away.
// ------------------------------------------------------------------------------------------
#ifdef ISY_DEBUG_BOUNDS
__device__ unsigned int g_isy_bounds_violations = 0;
__device__ unsigned int g_isy_bounds_first_line = 0;
#define ISY_BOUND(cond)                                                  \
    do {                                                                 \
        if (!(cond)) {                                                   \
            if (atomicAdd(&g_isy_bounds_violations, 1u) == 0) g_isy_bounds_first_line = __LINE__; \
        }                                                                \
    } while (0)
#else
#define ISY_BOUND(cond) do { } while (0)
#endif

namespace isy {

constexpr double kPi = 3.141592653589793238462643383279502884;
constexpr double kHalfPi = kPi / 2;     // np.pi / 2
constexpr double kTwoPi = 2 * kPi;      // 2 * np.pi
constexpr double kEps = 2.220446049250313e-16;  // invertsy/__helpers.py:10

// ------------------------------------------------------------------------------------------
// One painted copy of a visible triangle, as seen from one position (world.py:109-120).
// `Entry` is the fp32 filter record staged in shared memory (48 B, three LDS.128 per test,
// broadcast to the warp); `Exact` is the fp64 record read only by the rare exact re-check.
//
//   edge k of the reference's same-side test (world.py:482-526), k = 0: B->C (ref A),
//   1: A->C (ref B), 2: A->B (ref C):
//       inside  <=>  for all k   cross2(u_k, p - a_k) * o_k >= 0
//   Entry holds E_k(p) = (A_k p_theta + B_k p_phi + C_k), the same cross product multiplied by
//   sign(o_k) and scaled so |A_k| + |B_k| = 1; then |E_k(fp32) - E_k(exact)| < kEdgeEps for every
//   query in [-pi/2, pi/2] x (-pi, pi] and the sign of E_k decides the test unless |E_k| <= kEdgeEps.
// ------------------------------------------------------------------------------------------
struct __align__(16) Entry {
    float A0, B0, C0, A1;
    float B1, C1, A2, B2;
    float C2;
    int32_t tri;       // triangle index in the world
    int32_t dist_lo;   // min-vertex distance (world.py:97), fp64 bits
    int32_t dist_hi;
};
static_assert(sizeof(Entry) == 48, "Entry must be 48 bytes");

struct Exact {
    double th[3];  // vertex pitch-space angle  (world.py:110)
    double ph[3];  // vertex azimuth, seam-shifted for this copy (world.py:109, :117-120)
    double o[3];   // cross2(u_k, ref_k - a_k) as the reference computes it
};

// rounding budget of the fp32 evaluation of a normalised edge function ("fp32
// filter bound"): with u = 2^-24, |A|+|B| = 1, |p_theta| <= pi/2, |p_phi| <= pi, |C| <= 2pi:
// coefficient rounding 3pi*u + query rounding pi*u + two FMA roundings (3pi + 3.5pi)*u < 2.0e-6.
// The rasteriser forms the query azimuth in fp32 (yaw_f + heading_f, one wrap by the float 2pi): four more
// roundings of at most 2^-24 * 2pi plus the float-2pi error 1.75e-7, < 1.1e-6 in total -> < 3.1e-6;
// evenly spaced headings are formed as fma(k, delta, h0), at most 7.5e-7 from the true heading (5e-7 admitted by the uniformity check) -> < 3.9e-6.
constexpr float kEdgeEps = 5.0e-6f;

// per-view sun constants (sky.py:304-321, 501-521), computed once per view by sun_setup_kernel
struct SunConst {
    double sx, sy, sz;   // unit vector towards the sun: (sin ths cos phs, sin ths sin phs, cos ths)
    double i00;          // L(0, theta_s)                       sky.py:309
    double K;            // i_00 * i_90 / (i_00 - i_90 + eps)   sky.py:312 (grouped as written there)
    double i90;
    double inv_i00;      // 1 / (i_00 + eps)
    double Yz;           // zenith luminance                    sky.py:501-510
    double Mp;           // max degree of polarisation          sky.py:513-521
    float inv_i00_f, K_f, Mp2_f, pad_;   // float copies for the fp32 polarisation finish (Mp2 = 2/pi * Mp)
};

struct RenderParams {
    // world (SoA planes: x0 y0 z0 x1 y1 z1 x2 y2 z2, each T floats) and colours T*3
    const float* tri_planes;
    const float4* tri_aos;   // the same vertices, 3 float4 per triangle (x0 y0 z0 x1 | y1 z1 x2 y2 | z2 - - -)
    const float* tri_rgb;
    int T;
    double horizon;
    double horizon_sq_max;                // largest s with sqrt_rn(s) <= horizon: `sqrt(s) <= horizon` without the sqrt
    // xy cell grid over the world's vertices (cull candidates)
    const unsigned int* wg_cell_start;   // [ny*nx + 1] row-major
    const unsigned int* wg_cell_tris;    // triangle index | (vertex tag << 30)
    float wg_x0, wg_y0, wg_inv_cell;
    int wg_nx, wg_ny;
    // eye
    const double* eye_py;    // [R][2] eye-frame (pitch, yaw)
    const double* eye_trig;  // [R][4] sin pitch, cos pitch, sin yaw, cos yaw (eye frame)
    const double* eye_ftheta;// [R]    Perez gradation f(theta) of each eye ray for this call's sky (workspace)
    const double* eye_quat;  // [R][4]
    const float* eye_w;      // [R]
    // the eye's rays sorted by pitch, for the rasteriser (null = not available)
    const unsigned int* grid_plut;        // [kPitchLut + 1] first sorted ray at or above each pitch slot
    const void* grid_sorted;              // [R] float2 (pitch, yaw) of the rays in pitch order
    const int* grid_sorted_ray;           // [R] ray index of each sorted position
    // the eye's static (pitch, yaw) ray grid (few-headings rasteriser)
    const unsigned int* cells_start;      // [rows*cols + 1]
    const unsigned int* cells_list;       // [R] pitch-sorted positions of the rays, cell by cell
    int cells_rows, cells_cols;
    const int* grid_pos_of;               // [R] sorted position of ray r
    int N, S, R;
    // views
    int P, H;
    const double* pos;       // [P][3]
    const double* yaw;       // [P][H] or null
    const double* vquat;     // [P][H][4] or null
    const double* init_c;    // [B][R] external background luminance (world.py:79-82) or null
    const SunConst* sun;     // [1] or [B]
    const double2* sky_gtab; // [2][kGTab + 1] Hermite table of exp(D acos c) for this call's sky (workspace)
    int sun_per_view;
    // row templates (isy_shade.cuh, "Row templates + patches"); row_bg == null: not set up for this launch
    const float* sky_tab;    // [3][sky_rows_h][R] Y, DoP, AoP rows: per heading (Perez sky, positions share their
                             // headings) or one constant row each (uniform sky)
    int sky_rows_h;          // H, 1, or 0 = no sky outputs
    const int* sky_tab_off;  // nonzero: the positions turned out NOT to share their headings (heading_rows_kernel)
    const float* row_bg;     // [R] hit = -1 | [R] depth = NaN | [3R] ground colour / NaN by the ray's pitch
    int rows_vec4;           // rows are 16-byte aligned (R a multiple of 4, output buffers aligned): 16 bytes per lane
    const float* resp_bg;    // rows mode, [R]: photoreceptor response of a ray that hits nothing under no / a uniform sky
                             // (under a Perez sky it is plane 3 of sky_tab, one row per heading)
    isy_sky_params sky;
    isy_sensor_params sensor; // photoreceptor transform behind out.resp
    uint32_t flags;
    isy_outputs out;
    // workspace
    unsigned int* queue;     // dynamic work counter
    int* status;             // sticky error flag
    unsigned long long* phase_cycles;  // [8] summed SM cycles per phase (thread 0 of every CTA), profiling aid
    int* slab_vis;           // [grid][slab_cap] visible triangle ids
    Entry* slab_entry;       // [grid][slab_cap]
    Exact* slab_exact;       // [grid][slab_cap]
    float* slab_stage;       // [grid][9][stage_cap] vertices of the visible triangles of a position culled ahead (rows mode) or null
    int stage_cap;
    unsigned int* slab_best; // [grid][R] winner buffer in global memory when a heading does not fit smem, else null
    int slab_cap;
    int Hg, ngroups;         // headings per group, number of groups (H = sum of group sizes)
    int shade_nch;           // fast final pass: heading chunks per ray block (chosen by the host so that the warps finish together)
    int splits;              // CTAs sharing one position (each redoes the cull, takes every splits-th group)
    int ray_splits;          // general-orientation eyes, few positions: CTAs sharing the rays of one (position, group)
    int items;               // P * splits * ray_splits
};

// ---- photoreceptor transform behind out.resp (isy_sensor_params in isy_b200.h) ----
// polarisation term of a photoreceptor of sensitivity rho whose microvilli lie along the azimuth (cphi, sphi) of its
// viewing direction:  rho (sin^2(A - phi) + cos^2(A - phi) (1 - P)^2) + (1 - rho).  (X, Yd) = view direction minus sun
// direction in the xy plane gives the angle of polarisation without any trigonometry: A = atan2(Yd, X) + pi/2 (sky.py:323)
// => (sin A, cos A) = (X, -Yd) / |(X, Yd)|.
__device__ __forceinline__ float sensor_pol_term(float rho, float P, float X, float Yd, float cphi, float sphi) {
    const float h2 = X * X + Yd * Yd;
    if (!(rho > 0.f) || !(h2 > 1e-36f)) return 1.f;
    float ih;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ih) : "f"(h2));
    const float sA = X * ih, cA = -Yd * ih;
    const float cd = cA * cphi + sA * sphi, c2 = cd * cd, q = 1.f - P;
    return rho * ((1.f - c2) + c2 * q * q) + (1.f - rho);
}
// colour of one cone sample on the 0..1 scale: triangle colour, else ground, else the sky colour of its luminance
// (luminance2skycolour / 255, world.py:465-479) times the polarisation term `pol`, else nothing.  channels == 0
// (polarisation photoreceptor): the sky's luminance times the term where the sample sees the sky, else 0.
__device__ __forceinline__ void sensor_colour(const isy_sensor_params& sp, bool hit, float tr, float tg, float tb, bool ground,
                                              bool has_sky, double Y, float pol, float& r, float& g, float& b) {
    const bool sees_sky = !hit && !ground && has_sky;
    if (sp.channels == 0) {
        r = g = b = sees_sky ? (float)Y * pol : 0.f;
    } else if (hit) {
        r = tr, g = tg, b = tb;
    } else if (ground) {
        r = (float)(229.0 / 255.0), g = (float)(183.0 / 255.0), b = (float)(90.0 / 255.0);
    } else if (has_sky) {
        r = (float)(fmin(fmax(19.0 * Y + 13.0, 0.0), 255.0) / 255.0) * pol;
        g = (float)(fmin(fmax(19.0 * Y + 135.0, 0.0), 255.0) / 255.0) * pol;
        b = (float)(fmin(fmax(19.0 * Y + 201.0, 0.0), 255.0) / 255.0) * pol;
    } else {
        r = 0.f, g = 0.f, b = 0.f;
    }
}
// (cone-averaged) colour -> responses: clip(colour * w * gain, 0, 1) per channel
__device__ __forceinline__ void sensor_finish(const isy_sensor_params& sp, float& r, float& g, float& b) {
    r = fminf(fmaxf((r * sp.w_rgb[0]) * sp.gain, 0.f), 1.f);
    g = fminf(fmaxf((g * sp.w_rgb[1]) * sp.gain, 0.f), 1.f);
    b = fminf(fmaxf((b * sp.w_rgb[2]) * sp.gain, 0.f), 1.f);
}
// PN input of the three responses: clip(mean over channels, 0, 1)  (agent.py:744)
__device__ __forceinline__ float sensor_pn(float r, float g, float b) {
    return fminf(fmaxf((r + g + b) * (1.0f / 3.0f), 0.f), 1.f);
}
// stores the response(s) of ommatidium `oi` (element index b*N + n) from its cone-averaged 0..1 colour
template <typename T>
__device__ __forceinline__ void sensor_store(const isy_sensor_params& sp, void* resp, size_t oi, float r, float g, float b) {
    T* q = reinterpret_cast<T*>(resp);
    if (sp.channels == 0) {      // polarisation photoreceptor: gain * sqrt(averaged polarised luminance)
        __stcs(q + oi, (T)(sp.gain * sqrtf(fmaxf(r, 0.f))));
        return;
    }
    sensor_finish(sp, r, g, b);
    if (sp.channels == 1) {
        __stcs(q + oi, (T)sensor_pn(r, g, b));
    } else {
        __stcs(q + 3 * oi, (T)r), __stcs(q + 3 * oi + 1, (T)g), __stcs(q + 3 * oi + 2, (T)b);
    }
}

// ---- exact fp64 arithmetic in the reference's order (no FMA contraction) ----
__device__ __forceinline__ double cross2_rn(double u0, double u1, double v0, double v1) {
    // np.cross on 2-vectors: u0*v1 - u1*v0, each product rounded (world.py:526)
    return __dsub_rn(__dmul_rn(u0, v1), __dmul_rn(u1, v0));
}

__device__ __forceinline__ double norm2_rn(double x, double y, double z) {
    // np.linalg.norm(axis=2) before the sqrt: (x*x + y*y) + z*z (world.py:97)
    return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}

__device__ __forceinline__ bool inside_exact(const Exact& t, double p, double y) {
    // same_side_technique (world.py:482-503) on one painted copy
    double e0 = cross2_rn(t.th[2] - t.th[1], t.ph[2] - t.ph[1], p - t.th[1], y - t.ph[1]);
    if (!(__dmul_rn(e0, t.o[0]) >= 0)) return false;
    double e1 = cross2_rn(t.th[2] - t.th[0], t.ph[2] - t.ph[0], p - t.th[0], y - t.ph[0]);
    if (!(__dmul_rn(e1, t.o[1]) >= 0)) return false;
    double e2 = cross2_rn(t.th[1] - t.th[0], t.ph[1] - t.ph[0], p - t.th[0], y - t.ph[0]);
    return __dmul_rn(e2, t.o[2]) >= 0;
}

__device__ __forceinline__ void edge_coeffs(double a_th, double a_ph, double b_th, double b_ph, double o, float& A,
                                            float& B, float& C) {
    double u_th = b_th - a_th, u_ph = b_ph - a_ph;
    double n = fabs(u_th) + fabs(u_ph);
    if (o != o) {           // NaN reference cross: the reference's `>= 0` is False for every point
        A = 0.f, B = 0.f, C = -1.f;
        return;
    }
    if (o == 0.0 || !(n > 0.0)) {  // degenerate: product is +-0 -> test passes for every point (world.py:526)
        A = 0.f, B = 0.f, C = 1.f;
        return;
    }
    double s = (o > 0.0 ? 1.0 : -1.0) / n;
    // cross2(u, p - a) = u_th (p_ph - a_ph) - u_ph (p_th - a_th)
    A = (float)(-u_ph * s);
    B = (float)(u_th * s);
    C = (float)((u_ph * a_th - u_th * a_ph) * s);
}

__device__ __forceinline__ void finish_copy(const double th[3], const double ph[3], double dist, int tri, Entry& e,
                                            Exact& x) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        x.th[k] = th[k];
        x.ph[k] = ph[k];
    }
    x.o[0] = cross2_rn(th[2] - th[1], ph[2] - ph[1], th[0] - th[1], ph[0] - ph[1]);
    x.o[1] = cross2_rn(th[2] - th[0], ph[2] - ph[0], th[1] - th[0], ph[1] - ph[0]);
    x.o[2] = cross2_rn(th[1] - th[0], ph[1] - ph[0], th[2] - th[0], ph[2] - ph[0]);
    edge_coeffs(th[1], ph[1], th[2], ph[2], x.o[0], e.A0, e.B0, e.C0);
    edge_coeffs(th[0], ph[0], th[2], ph[2], x.o[1], e.A1, e.B1, e.C1);
    edge_coeffs(th[0], ph[0], th[1], ph[1], x.o[2], e.A2, e.B2, e.C2);
    e.tri = tri;
    e.dist_lo = __double2loint(dist);
    e.dist_hi = __double2hiint(dist);
}

// wrap an angle to (-pi, pi] (the range of SciPy's as_euler yaw)
__device__ __forceinline__ double wrap_yaw(double y) {
    if (y > kPi) {
        y -= kTwoPi;                               // exact for y in (pi, 3pi] (Sterbenz)
        if (y > kPi) {                             // far out of range: rare, general path
            y = remainder(y, kTwoPi);
            if (y <= -kPi) y += kTwoPi;
        }
    } else if (y <= -kPi) {
        y += kTwoPi;
        if (y <= -kPi) {
            y = remainder(y, kTwoPi);
            if (y <= -kPi) y += kTwoPi;
        }
    }
    return y;
}

// (a + pi) % (2 pi) - pi with NumPy's floored modulo (sky.py:105, :324): result in [-pi, pi).
// Callers pass a in (-3pi, 3pi), where the modulo is one conditional shift of the rounded sum.
__device__ __forceinline__ double wrap_pi_numpy(double a) {
    double m = a + kPi;
    if (m >= kTwoPi) m -= kTwoPi;
    else if (m < 0) m += kTwoPi;
    if (m >= kTwoPi || m < 0) {                    // not reached for |a| < 3pi; keeps the function total
        m = fmod(m, kTwoPi);
        if (m < 0) m += kTwoPi;
    }
    return m - kPi;
}

// first column of the rotation matrix of unit quaternion (x,y,z,w) = ori.apply([1,0,0]) (sky.py:101)
__device__ __forceinline__ void quat_dir(double x, double y, double z, double w, double& dx, double& dy, double& dz) {
    dx = 1.0 - 2.0 * (y * y + z * z);
    dy = 2.0 * (x * y + w * z);
    dz = 2.0 * (x * z - w * y);
}

// ZYX Euler yaw/pitch of a direction (world.py:84): d = (cos p cos y, cos p sin y, -sin p)
__device__ __forceinline__ void dir_pitch_yaw(double dx, double dy, double dz, double& pitch, double& yaw) {
    yaw = atan2(dy, dx);
    pitch = atan2(-dz, sqrt(dx * dx + dy * dy));
}

// Perez luminance L(chi, z) with cos(z) and cos(chi) supplied (sky.py:348-374)
__device__ __forceinline__ double perez_f(double cos_z, bool above, double A, double B) {
    return above ? 1.0 + A * exp(B / (cos_z + kEps)) : 0.0;
}
__device__ __forceinline__ double perez_phi(double chi, double cos_chi, double C, double D, double E) {
    return 1.0 + C * exp(D * chi) + E * (cos_chi * cos_chi);
}

// Sky.__call__ for one direction (sky.py:304-334); d = unit view direction, theta = zenith angle.
// f_theta = perez_f(cos theta, theta < pi/2) is passed in: it depends on the elevation only and is
// tabulated per eye ray for yaw-only views.
// kPrecise = true : every step in fp64 in the reference's order (double outputs, ~1e-13 of the reference)
// kPrecise = false: float outputs. Luminance keeps the fp64 core (gamma, exp); the polarisation
//                   finish (one atan2, two divisions) runs in fp32 on fp64-rounded operands, which
//                   bounds the added error by ~1e-6 (DESIGN.md "sky error budget"), inside the 1e-5 bar.
constexpr int kGTab = 256;   // intervals of the G(c) = exp(D acos c) table, per half (c >= 0, c < 0)

// G(c) = exp(D * acos(c)), the only place the angular distance to the sun enters the luminance
// (sky.py:308 via :372), from a cubic Hermite table in u = sqrt(1 - |c|): acos(1 - u^2) = 2 asin(u / sqrt 2)
// is analytic on [0, 1], so 256 intervals give |error| < 1e-9 for any |D| < 4 (tests/test_oracle.py checks the
// construction on the CPU, DESIGN.md "sky error budget").  tab[half][i] = {G(u_i), dG/du(u_i) * h}.
__device__ __forceinline__ double g_table_eval(const double2* __restrict__ tab, double c) {
    // sqrt(x): float rsqrt seed (bare MUFU.RSQ) + one Newton step in fp64, relative error ~1e-14.
    // x is kept >= 1e-30 (u = 1e-15 instead of 0 moves G by 3e-15), which makes the step branch-free.
    const double x = fmax(1.0 - fabs(c), 1e-30);
    float r0f;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r0f) : "f"((float)x));
    const double r0 = (double)r0f;
    const double u = x * (r0 * (1.5 - 0.5 * x * r0 * r0));
    const double um = u * (double)kGTab;
    const int idx = min((int)um, kGTab - 1);
    const double t = um - (double)idx;
    const double2* tb = tab + (c < 0.0 ? kGTab + 1 : 0) + idx;
    const double2 a = tb[0], b = tb[1];
    const double c2 = 3.0 * (b.x - a.x) - 2.0 * a.y - b.y;
    const double c3 = 2.0 * (a.x - b.x) + a.y + b.y;
    return fma(t, fma(t, fma(t, c3, c2), a.y), a.x);
}

// atan2 for the float polarisation angle: odd degree-15 minimax polynomial on [0, 1] (max error 1.4e-7
// in fp32 evaluation) after one fast division, then the usual octant fix-ups; |error| < 5e-7 rad
// 1 / x by the bare MUFU.RCP (1 ulp, flushes denormals): for operands known to be normal floats, where
// __fdividef's range fix-ups for huge denominators only cost instructions
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float atan2_fast(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float hi = fmaxf(ax, ay), lo = fminf(ax, ay);
    const float a = hi > 1e-30f ? lo * rcp_fast(hi) : 0.f;   // (hi below 1e-30: the direction IS the sun's, any angle is right)
    const float s = a * a;
    float p = -0.004054563585668802f;
    p = fmaf(p, s, 0.021862946450710297f);
    p = fmaf(p, s, -0.0559123158454895f);
    p = fmaf(p, s, 0.09642196446657181f);
    p = fmaf(p, s, -0.1390862911939621f);
    p = fmaf(p, s, 0.19946566224098206f);
    p = fmaf(p, s, -0.33329859375953674f);
    p = fmaf(p, s, 0.9999993443489075f);
    float r = p * a;
    if (ay > ax) r = (float)kHalfPi - r;
    if (x < 0.f) r = (float)kPi - r;
    return y < 0.f ? -r : r;
}

template <bool kPrecise>
__device__ __forceinline__ void sky_eval(const isy_sky_params& sp, const SunConst& sc, double dx, double dy, double dz,
                                         double theta, double f_theta, double& Y, double& P, double& A,
                                         bool masked = false, const double2* __restrict__ gtab = nullptr) {
    double cg = dz * sc.sz + (dx * sc.sx + dy * sc.sy);      // cos(gamma), sky.py:304-305
    const double cos_t = dz;
    double gamma;
    if (kPrecise) {
        cg = fmin(fmax(cg, -1.0), 1.0);
        gamma = acos(cg);
        double i_prez = f_theta * perez_phi(gamma, cg, sp.C, sp.D, sp.E);           // sky.py:308
        double i = (1.0 / (i_prez + kEps) - 1.0 / (sc.i00 + kEps)) * sc.K;          // sky.py:312
        Y = fmax(sc.Yz * i_prez / (sc.i00 + kEps), 0.0);                            // sky.py:313
        double lp = fmax(1.0 - cg * cg, 0.0) / (1.0 + cg * cg);                     // sky.py:316
        double p = 2.0 / kPi * sc.Mp * lp * (theta * cos_t + (kHalfPi - theta) * i);  // sky.py:317
        P = fmin(fmax(p, 0.0), 1.0);
        A = wrap_pi_numpy(atan2(dy - sc.sy, dx - sc.sx) + kHalfPi);                 // sky.py:320-324
    } else {
        // gamma itself is never needed: exp(D gamma) comes from the table, the sun disc is a test on cos(gamma)
        gamma = cg > 0.99862953475457383 ? 0.0 : kPi;                    // cos(pi / 60)
        // f_theta is exactly 0 for every direction below the horizon (sky.py:366-370): the product is 0 whatever
        // the finite bracket holds, so those rays (half of a full-sphere eye, whole warps of it) skip the table
        const double g = f_theta != 0.0 ? g_table_eval(gtab, cg) : 0.0;
        const double i_prez = f_theta * (1.0 + sp.C * g + sp.E * (cg * cg));
        const double y = sc.Yz * i_prez * sc.inv_i00;
        Y = y > 0.0 ? y : 0.0;                                          // max(y, 0) of sky.py:313 (y is never NaN)
        const float cg2 = fminf((float)(cg * cg), 1.0f);
        const float lp = (1.0f - cg2) * rcp_fast(1.0f + cg2);
        const float i = (rcp_fast((float)i_prez + (float)kEps) - sc.inv_i00_f) * sc.K_f;   // >= 2.2e-16: normal
        const float p = sc.Mp2_f * lp * fmaf((float)(kHalfPi - theta), i, (float)(theta * cos_t));
        P = (double)fminf(fmaxf(p, 0.0f), 1.0f);
        float a = atan2_fast((float)(dy - sc.sy), (float)(dx - sc.sx)) + (float)kHalfPi;   // in (-pi/2, 3pi/2]
        if (a >= (float)kPi) a -= (float)kTwoPi;
        A = (double)a;
    }
    if (masked) Y = 0.0, P = 0.0, A = CUDART_NAN;                               // sky.py:330-332 (eta)
    if (gamma < kPi / 60) Y = 17.0;                                             // sky.py:334
}

}  // namespace isy

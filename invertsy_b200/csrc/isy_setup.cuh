// Small set-up kernels: eye tables, sky tables, per-view sun constants.
#pragma once
#include "isy_layout.cuh"

namespace isy {

// Head of a render workspace: [0] work counter, [32] "positions differ in their headings" flag (+128 B), [64] sticky
// status (+256 B), [66..67] cookie, [80..95] phase cycles (+320 B).  Every launch restarts the counter, the flag and
// the cycle sums; the status survives until it is read (isy_workspace_status), so that an overflow in any launch
// on this workspace is reported.  A workspace seen for the first time (no cookie) starts with a clean status.
constexpr unsigned int kWsCookie0 = 0x69737962u, kWsCookie1 = 0x32303062u;
__global__ void ws_reset_kernel(unsigned int* __restrict__ head) {
    const int t = threadIdx.x;
    if (t == 0) {
        head[0] = 0u, head[32] = 0u;
        if (head[66] != kWsCookie0 || head[67] != kWsCookie1) head[64] = 0u, head[66] = kWsCookie0, head[67] = kWsCookie1;
    }
    if (t < 16) head[80 + t] = 0u;
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

// Do all positions of the call look along the same row of headings?  (yaw[p][h] == yaw[0][h] bit for bit; grids
// and heat maps do.)  *differs is left at 0 if so.  One pass over the P*H headings, a few microseconds.
__global__ void heading_rows_kernel(const double* __restrict__ yaw, long long n, int H, int* __restrict__ differs) {
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        bad |= __double_as_longlong(yaw[i]) != __double_as_longlong(yaw[i % H]);
    if (bad) *differs = 1;
}

// The sky of a yaw-only view under one sun depends on (heading, ray) only: when the positions share their
// headings it is evaluated once per launch -- here, with the arithmetic of the per-view final pass -- into
// tab[3][H][R] (Y, DoP, AoP as the float outputs), and the render kernel's final pass copies it.
__global__ void sky_shared_kernel(const double* __restrict__ yaw, int H, int R, const double* __restrict__ eye_py,
                                  const double* __restrict__ eye_trig, const double* __restrict__ f_theta,
                                  isy_sky_params sp, const SunConst* __restrict__ sun,
                                  const double2* __restrict__ gtab, const int* __restrict__ differs,
                                  float* __restrict__ tab, int want_resp, isy_sensor_params sensor) {
    if (*differs) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * R) return;
    const int h = i / R, r = i - h * R;
    const double hw = wrap_yaw(yaw ? yaw[h] : 0.0);
    double hsn, hcs;
    sincos(hw, &hsn, &hcs);
    const double4 t4v = *reinterpret_cast<const double4*>(eye_trig + 4 * (size_t)r);
    const double sy = t4v.z * hcs + t4v.w * hsn, cy = t4v.w * hcs - t4v.z * hsn;   // yaw_e + heading
    double Y = 0, P = 0, A = CUDART_NAN;
    sky_eval<false>(sp, sun[0], t4v.y * cy, t4v.y * sy, -t4v.x, eye_py[2 * (size_t)r] + kHalfPi, f_theta[r], Y, P, A,
                    false, gtab);
    const size_t plane = (size_t)H * (size_t)R;
    tab[i] = (float)Y, tab[plane + i] = (float)P, tab[2 * plane + i] = (float)A;
    if (want_resp) {   // plane 3: the PN response of this (heading, ray) when the ray hits nothing
        float sr, sg, sb;
        const float pol = sensor_pol_term(sensor.pol_op, (float)P, (float)(t4v.y * cy - sun[0].sx), (float)(t4v.y * sy - sun[0].sy),
                                          (float)cy, (float)sy);
        sensor_colour(sensor, false, 0.f, 0.f, 0.f, eye_py[2 * (size_t)r] >= 0, true, Y, pol, sr, sg, sb);
        sensor_finish(sensor, sr, sg, sb);
        tab[3 * plane + i] = sensor_pn(sr, sg, sb);
    }
}

// Sky-only calls (an empty world, only Y / DoP / AoP asked for: sun sweeps over a dorsal-rim eye, examples/test_sky.py):
// no position has anything to cull or rasterise, so the views are a pure element-wise map (view, ray) -> sky.  One CTA
// takes a view at a time; consecutive threads take consecutive rays, so every store instruction covers 128 bytes of
// one output row.  The arithmetic is sky_eval's, the same as in the render kernel's final pass.
__global__ void __launch_bounds__(256) sky_views_kernel(const double* __restrict__ yaw, long long B, int R,
                                                        const double* __restrict__ eye_py, const double* __restrict__ eye_trig,
                                                        const double* __restrict__ f_theta, isy_sky_params sp,
                                                        const SunConst* __restrict__ sun, int sun_per_view,
                                                        const double2* __restrict__ gtab_g, float* __restrict__ oY,
                                                        float* __restrict__ oP, float* __restrict__ oA) {
    __shared__ double2 gtab[2 * (kGTab + 1)];
    const bool perez = sp.kind == ISY_SKY_PEREZ;
    if (perez)
        for (int i = threadIdx.x; i < 2 * (kGTab + 1); i += blockDim.x) gtab[i] = gtab_g[i];
    __syncthreads();
    const float nanf_ = __int_as_float(0x7fc00000);
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        double hsn = 0, hcs = 1;
        SunConst sc;
        if (perez) {
            sincos(wrap_yaw(yaw ? yaw[b] : 0.0), &hsn, &hcs);
            sc = sun[sun_per_view ? b : 0];
        }
        const size_t row = (size_t)b * (size_t)R;
        for (int r = threadIdx.x; r < R; r += blockDim.x) {
            double Y = sp.luminance, P = 0, A = CUDART_NAN;                      // UniformSky: sky.py:77-79
            if (perez) {
                const double4 t4v = *reinterpret_cast<const double4*>(eye_trig + 4 * (size_t)r);
                const double sy = t4v.z * hcs + t4v.w * hsn, cy = t4v.w * hcs - t4v.z * hsn;   // yaw_e + heading
                sky_eval<false>(sp, sc, t4v.y * cy, t4v.y * sy, -t4v.x, eye_py[2 * (size_t)r] + kHalfPi, f_theta[r], Y, P, A,
                                false, gtab);
            }
            if (oY) __stcs(oY + row + r, (float)Y);
            if (oP) __stcs(oP + row + r, (float)P);
            if (oA) __stcs(oA + row + r, perez ? (float)A : nanf_);
        }
    }
}

// Template rows of a launch in rows mode (isy_shade.cuh): what a ray that hits nothing gets -- hit = -1, depth = NaN,
// the ground colour or NaN by its own pitch (world.py:80, :90) -- and, under a uniform sky, the three constant sky
// rows (sky.py:77-79: y = luminance, p = 0, a = NaN).
__global__ void row_templates_kernel(const double* __restrict__ eye_py, int R, int uniform_sky, double luminance,
                                     float* __restrict__ bg, float* __restrict__ sky_rows, int want_resp, int has_sky,
                                     isy_sensor_params sensor) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float nanf_ = __int_as_float(0x7fc00000);
    bg[r] = __int_as_float(-1);
    bg[R + r] = nanf_;
    const bool ground = eye_py[2 * (size_t)r] >= 0;
    bg[2 * (size_t)R + 3 * r + 0] = ground ? (float)(229.0 / 255.0) : nanf_;
    bg[2 * (size_t)R + 3 * r + 1] = ground ? (float)(183.0 / 255.0) : nanf_;
    bg[2 * (size_t)R + 3 * r + 2] = ground ? (float)(90.0 / 255.0) : nanf_;
    if (uniform_sky) sky_rows[r] = (float)luminance, sky_rows[R + r] = 0.f, sky_rows[2 * (size_t)R + r] = nanf_;
    if (want_resp) {   // PN response of a ray that hits nothing under no sky / a uniform sky (constant luminance)
        float sr, sg, sb;
        sensor_colour(sensor, false, 0.f, 0.f, 0.f, ground, has_sky != 0, luminance, 1.f, sr, sg, sb);   // (uniform sky: DoP = 0)
        sensor_finish(sensor, sr, sg, sb);
        bg[5 * (size_t)R + r] = sensor_pn(sr, sg, sb);
    }
}

}  // namespace isy

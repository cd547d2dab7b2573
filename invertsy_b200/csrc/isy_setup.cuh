// Small set-up kernels: eye tables, sky tables, per-view sun constants.
#pragma once
#include "isy_layout.cuh"

namespace isy {

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

}  // namespace isy

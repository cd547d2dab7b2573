/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of the InvertSy eye-rendering hot path.
 *
 * This is the fast CPU checker used by tests/ (large parity cases), __graft_entry__.smoke()
 * and bench.py's cpu_baseline legs.  The product (invertsy_b200/csrc) never links or calls it.
 *
 * Parity status: pinned -- tests/test_oracle.py checks it against the golden fixtures made
 * from the unmodified reference (tests/golden/, oracle/gen_golden.py) and against the NumPy
 * restatement oracle/isy_oracle.py, which is itself bit-identical to the reference here.
 *
 * Follows /root/reference/src/invertsy/env/world.py:84-123, 482-526 and
 *         /root/reference/src/invertsy/env/sky.py:99-113, 269-334, 348-374, 501-521.
 *
 * All arithmetic is IEEE double with the reference's operation order; build with
 * -ffp-contract=off so that a*b - c*d is two rounded products and a rounded difference,
 * which is what NumPy's element-wise ufuncs do.  libm's atan2/acos/exp/cos/sin may differ
 * from NumPy's SIMD kernels by an ulp; decisions can then only differ for rays within
 * ~1e-16 rad of a triangle edge (documented near-tie exception).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ISY_PI 3.141592653589793238462643383279502884
#define ISY_EPS 2.220446049250313e-16 /* invertsy/__helpers.py:10 */

typedef struct {
    double th[3], ph[3];  /* vertex (theta, phi) of one painted copy */
    double o[3];          /* the three "reference point" crosses      */
    double dist;
    int32_t tri;
} otri_t;

static inline double cross2(double u0, double u1, double v0, double v1) { return u0 * v1 - u1 * v0; }

/* env/world.py:482-526: the three same-side tests, (theta, phi) component order */
static inline int inside_copy(const otri_t* t, double p, double y) {
    const double *th = t->th, *ph = t->ph;
    /* same_side(P, A; B, C): edge B->C */
    double e0 = cross2(th[2] - th[1], ph[2] - ph[1], p - th[1], y - ph[1]);
    if (!(e0 * t->o[0] >= 0)) return 0;
    /* same_side(P, B; A, C): edge A->C */
    double e1 = cross2(th[2] - th[0], ph[2] - ph[0], p - th[0], y - ph[0]);
    if (!(e1 * t->o[1] >= 0)) return 0;
    /* same_side(P, C; A, B): edge A->B */
    double e2 = cross2(th[1] - th[0], ph[1] - ph[0], p - th[0], y - ph[0]);
    if (!(e2 * t->o[2] >= 0)) return 0;
    return 1;
}

static void finish_copy(otri_t* t) {
    const double *th = t->th, *ph = t->ph;
    t->o[0] = cross2(th[2] - th[1], ph[2] - ph[1], th[0] - th[1], ph[0] - ph[1]);
    t->o[1] = cross2(th[2] - th[0], ph[2] - ph[0], th[1] - th[0], ph[1] - ph[0]);
    t->o[2] = cross2(th[1] - th[0], ph[1] - ph[0], th[2] - th[0], ph[2] - ph[0]);
}

static int cmp_otri(const void* a, const void* b) {
    const otri_t *x = (const otri_t*)a, *y = (const otri_t*)b;
    if (x->dist < y->dist) return -1;
    if (x->dist > y->dist) return 1;
    /* exact ties are the documented exception; the painter paints the later (argsort-reversed) one last */
    return (x->tri > y->tri) - (x->tri < y->tri);
}

/*
 * Nearest containing triangle per query direction (env/world.py:94-123).
 *   tri      T*9 floats  [t][vertex][x,y,z]   (WorldBase.polygons)
 *   pos      3 doubles   (nanmean(pos) of the reference call)
 *   pitch,yaw N doubles  (ori.as_euler('ZYX'))
 *   hit      N int32     winning triangle index or -1
 *   depth    N doubles   its min-vertex distance or NaN
 * returns the number of visible triangles, or -1 on allocation failure.
 */
long isy_oracle_world_hits(const float* tri, long T, double horizon, const double* pos, const double* pitch,
                           const double* yaw, long N, int32_t* hit, double* depth) {
    otri_t* list = (otri_t*)malloc(sizeof(otri_t) * (size_t)(2 * T + 1));
    if (!list) return -1;
    long n = 0, nvis = 0;
    for (long t = 0; t < T; ++t) {
        double v[3][3], d = INFINITY;
        int bad = 0;
        for (int k = 0; k < 3; ++k) {
            for (int c = 0; c < 3; ++c) v[k][c] = (double)tri[t * 9 + k * 3 + c] - pos[c];
            double r = sqrt(v[k][0] * v[k][0] + v[k][1] * v[k][1] + v[k][2] * v[k][2]); /* :97 */
            if (isnan(r)) bad = 1;
            if (r < d) d = r;
        }
        if (bad || !(d <= horizon)) continue; /* :98, :102 */
        ++nvis;
        otri_t a;
        a.dist = d;
        a.tri = (int32_t)t;
        double phmax = -INFINITY, phmin = INFINITY;
        for (int k = 0; k < 3; ++k) {
            a.ph[k] = atan2(v[k][1], v[k][0]);                                                   /* :109 */
            a.th[k] = atan2(sqrt(v[k][0] * v[k][0] + v[k][1] * v[k][1]), v[k][2]) - ISY_PI / 2; /* :110 */
            if (a.ph[k] > phmax) phmax = a.ph[k];
            if (a.ph[k] < phmin) phmin = a.ph[k];
        }
        if (phmax - phmin >= ISY_PI) { /* :116-120: two seam-side copies */
            otri_t p1 = a, p2 = a;
            for (int k = 0; k < 3; ++k) {
                if (a.ph[k] < 0) p1.ph[k] = a.ph[k] + 2 * ISY_PI;
                if (a.ph[k] > 0) p2.ph[k] = a.ph[k] - 2 * ISY_PI;
            }
            finish_copy(&p1);
            finish_copy(&p2);
            list[n++] = p1;
            list[n++] = p2;
        } else {
            finish_copy(&a);
            list[n++] = a;
        }
    }
    qsort(list, (size_t)n, sizeof(otri_t), cmp_otri);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) {
        int32_t h = -1;
        double dd = NAN;
        for (long j = 0; j < n; ++j) {
            if (inside_copy(&list[j], pitch[i], yaw[i])) {
                h = list[j].tri;
                dd = list[j].dist;
                break;
            }
        }
        hit[i] = h;
        depth[i] = dd;
    }
    free(list);
    return nvis;
}

/* env/sky.py:348-374 */
static inline double perez_L(double chi, double z, const double* k) {
    double f = 0.0;
    if (z < ISY_PI / 2) f = 1.0 + k[0] * exp(k[1] / (cos(z) + ISY_EPS));
    double cc = cos(chi);
    double phi = 1.0 + k[2] * exp(k[3] * chi) + k[4] * (cc * cc);
    return f * phi;
}

static inline double wrap_pi(double a) {
    /* (a + pi) % (2 pi) - pi with Python/NumPy floored-modulo semantics */
    double m = fmod(a + ISY_PI, 2 * ISY_PI);
    if (m < 0) m += 2 * ISY_PI;
    return m - ISY_PI;
}

/*
 * Sky.__call__ (env/sky.py:269-334) for N viewing directions given as unit vectors
 * (ori.apply([1,0,0])), noise = 0.
 *   coef = A,B,C,D,E ; tau_L after the pinv round trip ; c1,c2
 *   sun_xy = x,y of R.from_euler('ZY',[phi_s, pi/2-theta_s]).apply([1,0,0])  (sky.py:320-321)
 */
void isy_oracle_sky(const double* dir, long N, double theta_s, double phi_s, const double* coef, double tau_L,
                    double c1, double c2, const double* sun_xy, double* Y, double* P, double* A) {
    double chi = (4. / 9. - tau_L / 120.) * (ISY_PI - 2 * (ISY_PI / 2 - theta_s));
    double Y_z = (4.0453 * tau_L - 4.9710) * tan(chi) - 0.2155 * tau_L + 2.4192; /* :509-510 */
    double M_p = exp(-(tau_L - c1) / (c2 + ISY_EPS));                            /* :521 */
    double i_00 = perez_L(0., theta_s, coef);
    double i_90 = perez_L(ISY_PI / 2, fabs(theta_s - ISY_PI / 2), coef);
    double cs = cos(theta_s), ss = sin(theta_s);
#pragma omp parallel for schedule(static)
    for (long n = 0; n < N; ++n) {
        double x = fmin(fmax(dir[3 * n + 0], -1.), 1.), y = fmin(fmax(dir[3 * n + 1], -1.), 1.),
               z = fmin(fmax(dir[3 * n + 2], -1.), 1.);
        double phi = wrap_pi(atan2(y, x));
        double theta = acos(z);
        double gamma = acos(cos(theta) * cs + sin(theta) * ss * cos(phi - phi_s));
        double i_prez = perez_L(gamma, theta, coef);
        double i = (1. / (i_prez + ISY_EPS) - 1. / (i_00 + ISY_EPS)) * i_00 * i_90 / (i_00 - i_90 + ISY_EPS);
        double lum = fmax(Y_z * i_prez / (i_00 + ISY_EPS), 0.);
        double sg = sin(gamma), cg = cos(gamma);
        double lp = (sg * sg) / (1 + cg * cg);
        double p = 2. / ISY_PI * M_p * lp * (theta * cos(theta) + (ISY_PI / 2 - theta) * i);
        p = fmin(fmax(p, 0.), 1.);
        double a = wrap_pi(atan2(dir[3 * n + 1] - sun_xy[1], dir[3 * n + 0] - sun_xy[0]) + ISY_PI / 2);
        if (gamma < ISY_PI / 60) lum = 17;
        Y[n] = lum;
        P[n] = p;
        A[n] = a;
    }
}

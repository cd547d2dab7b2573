/*
 * isy_b200.h -- C ABI of the B200-native (sm_100a) compound-eye view renderer.
 *
 * This is the drop-in boundary for ONE hot path of InsectRobotics/InvertSy: what a set of
 * viewing directions sees of a triangle world (invertsy.env.world.WorldBase.__call__) fused
 * with the analytic sky (invertsy.env.sky.Sky.__call__ / UniformSky.__call__).
 *
 * The reference has no FFI: the interface it exposes for this path is the pair of Python
 * callables
 *     scene(pos, ori=..., init_c=..., noise=..., eta=..., rng=...) -> (N,3)   env/world.py:55
 *     sky(ori, irgbu=..., noise=..., eta=..., rng=...) -> (Y, DoP, AoP)        env/sky.py:269, :47
 * so each entry point below names the reference lines it replaces and the ctypes binding a
 * maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative isy_status and
 *     leaves a message in isy_last_error() (thread-local).  No exceptions cross the boundary.
 *   - "dev" pointers are CUDA device pointers on the handle's device; "host" pointers are
 *     ordinary (ideally pinned) host memory.  The caller owns every buffer it passes.
 *   - handles (isy_world, isy_eye) own their device copies and are immutable after creation;
 *     render calls are re-entrant on distinct streams with distinct workspaces.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All device
 *     entry points are asynchronous with respect to the host.
 *   - orientations are unit quaternions in SciPy order (x, y, z, w); angles are radians.
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     ISY_ERR_CUDA.
 */
#ifndef ISY_B200_H_
#define ISY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISY_ABI_VERSION 2

typedef enum {
    ISY_OK = 0,
    ISY_ERR_INVALID = -1,   /* bad argument (NULL, negative size, inconsistent flags)            */
    ISY_ERR_CUDA = -2,      /* CUDA runtime error / no device                                     */
    ISY_ERR_WORKSPACE = -3, /* workspace too small (use isy_render_workspace_bytes)               */
    ISY_ERR_OVERFLOW = -4   /* a position saw more triangles than the workspace slab can hold     */
} isy_status;

/* what the sky contributes (env/sky.py) */
typedef enum {
    ISY_SKY_NONE = 0,    /* no sky: Y/DoP/AoP are not produced                                    */
    ISY_SKY_UNIFORM = 1, /* UniformSky.__call__  env/sky.py:47-97: Y = luminance, DoP = 0, AoP = NaN */
    ISY_SKY_PEREZ = 2    /* Sky.__call__         env/sky.py:269-346                               */
} isy_sky_kind;

/* Scalars of one Sky instance (env/sky.py:245-267, 376-554). Filled by the Python mirror. */
typedef struct {
    int32_t kind;      /* isy_sky_kind                                                            */
    int32_t reserved;
    double luminance;  /* UniformSky luminance (sky.py:34)                                        */
    double A, B, C, D, E; /* Perez coefficients (sky.py:376-457)                                  */
    double tau_L;      /* turbidity after the pinv round trip (sky.py:535-554)                    */
    double c1, c2;     /* max-DoP coefficients (sky.py:264-265)                                   */
} isy_sky_params;

/* output selection / behaviour flags for the render calls */
enum {
    ISY_FLAG_INIT_FROM_SKY = 1 << 0, /* world starts from luminance2skycolour(Y) (world.py:82, :465-479)
                                        instead of NaN (world.py:80); rgb is then on the reference's
                                        mixed 0..255 / 0..1 scale                                     */
    ISY_FLAG_BRUTE_FORCE = 1 << 1,   /* test every visible triangle for every ray (no angular bins);
                                        same results, used as an on-device cross-check                */
    ISY_FLAG_SUN_PER_VIEW = 1 << 2,  /* `sun` holds one (theta_s, phi_s) per view instead of one pair  */
    ISY_FLAG_OUT_F64 = 1 << 3,       /* rgb / depth / Y / dop / aop buffers are double instead of float */
    ISY_FLAG_PER_VIEW_SKY = 1 << 4,  /* evaluate the sky for every view even when all positions share one row of
                                        headings under one sun (by default such calls evaluate it once per
                                        (heading, ray) and copy it into each view); same results to rounding,
                                        used as a cross-check                                             */
    ISY_FLAG_NO_TILES = 1 << 5       /* scans with evenly spaced headings are rasterised copy by copy (column walk)
                                        instead of tile by tile; same results, used as a cross-check        */
};

/* Per-ray / per-ommatidium outputs. Any pointer may be NULL (not produced).
 * B = number of views, R = rays per view (N ommatidia x S samples), N = ommatidia.        */
typedef struct {
    void* rgb;      /* [B][N][3]  world colour, cone-averaged over the S samples (world.py:79-123)   */
    int32_t* hit;   /* [B][R]     winning triangle index per ray, -1 = none (painter's top, world.py:106-123) */
    void* depth;    /* [B][R]     min-vertex distance of that triangle (world.py:97), NaN = none     */
    void* Y;        /* [B][N]     sky luminance      (sky.py:313, :334), cone-averaged               */
    void* dop;      /* [B][N]     degree of polarisation (sky.py:317)                                */
    void* aop;      /* [B][N]     angle of polarisation  (sky.py:320-324), circular cone average     */
    void* resp;     /* [B][N][C]  photoreceptor responses of the eye (needs isy_sensor_params), C = 3 or 1 (or 0 -> 1 value):
                                  what `eye(sky=, scene=)` leaves in eye.responses (simulation.py:2707, :2721),
                                  or, with C = 1, the PN input clip(responses.mean(axis=1), 0, 1) (agent.py:744)  */
} isy_outputs;      /* element type: float, or double with ISY_FLAG_OUT_F64                           */

/* Photoreceptor transform of a CompoundEye-compatible sensor (rows N1 / N2 of SURVEY.md 8f).  The reference takes
 * its eye from InvertPy (agent.py:20, simulation.py:24), which is not vendored: the transform below is this
 * library's own, documented one (invertsy_b200/sense.py).  Per cone sample the colour on the 0..1 scale is
 *     the triangle's colour where one is hit (world.py:123), else the ground colour where pitch >= 0 (world.py:90),
 *     else luminance2skycolour(Y) / 255 under a sky (world.py:465-479), else 0;
 * the samples of an ommatidium are averaged with the cone weights; then, per channel c in (R, G, B),
 *     resp_c = clip(mean_c * w_rgb[c] * gain, 0, 1)                (c_sensitive, omm_res of the eye's constructor)
 * and with channels == 1 the output is clip((resp_R + resp_G + resp_B) / 3, 0, 1).
 * Polarisation (omm_pol_op = rho in [0, 1]): a sample that sees the sky is multiplied by
 *     rho * (sin^2(A - phi) + cos^2(A - phi) * (1 - P)^2) + (1 - rho)
 * with P, A the degree and angle of polarisation of the sky in its direction (sky.py:317-324) and phi the azimuth of
 * the sample's own viewing direction (the microvilli of a dorsal-rim ommatidium lie along it); 1 for rho = 0.
 * channels == 0 is the polarisation photoreceptor of a PolarisationSensor (agent.py:835, :879): one value per
 * ommatidium, resp = gain * sqrt(cone average of Y * that term over the samples that see the sky, 0 for the others). */
typedef struct {
    float w_rgb[3];    /* spectral sensitivity to R, G, B: c_sensitive[1:4] of the IRGBU vector (agent.py:590)  */
    float gain;        /* omm_res / saturation (agent.py:589)                                                    */
    int32_t channels;  /* 3, 1 or 0 (see above)                                                                  */
    float pol_op;      /* omm_pol_op: polarisation sensitivity of the photoreceptors, 0 = none                   */
} isy_sensor_params;

typedef struct isy_world isy_world;
typedef struct isy_eye isy_eye;

/* ---- library ------------------------------------------------------------------------------- */
int isy_abi_version(void);
const char* isy_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
int64_t isy_launch_count(void);
/* Bounds-checked builds (-DISY_DEBUG_BOUNDS, tools/debug_bounds.py): synchronises `device` and returns how many of the
 * kernels' own index asserts have failed since load (*first_line = source line of the first); -1 in a release build. */
int64_t isy_debug_bounds_violations(int device, int32_t* first_line);

/* ---- world: WorldBase.__init__ / .polygons / .colours / .horizon  (world.py:48-53, 131-163) ---
 * tri_xyz: T*9 floats [t][vertex][x,y,z] (host) ; tri_rgb: T*3 floats (host) ; both copied.     */
int isy_world_create(const float* tri_xyz, const float* tri_rgb, int64_t T, double horizon, int device,
                     isy_world** out);
int isy_world_destroy(isy_world* w);
int64_t isy_world_triangles(const isy_world* w);

/* ---- eye layout: the ommatidia orientations a CompoundEye hands to scene()/sky() -------------
 * omm_quat: N*S*4 doubles (host), ray r = n*S + s is sample s of ommatidium n, eye frame.
 * weights : N*S floats (host) cone weights, each ommatidium's S weights sum to 1 (NULL = 1/S).
 * Replaces the `omm_ori` argument flow agent.py:742 -> InvertPy CompoundEye -> world.py:55 / sky.py:269. */
int isy_eye_create(const double* omm_quat, const float* weights, int64_t N, int32_t S, int device, isy_eye** out);
int isy_eye_destroy(isy_eye* e);
int64_t isy_eye_rays(const isy_eye* e);

/* ---- batched render: P positions x H orientations each = B = P*H views ----------------------
 * The unit of work is a *position*: cull + angular projection of the world (world.py:94-110)
 * is done once per position and shared by its H orientations.
 *   pos_dev      [P][3] doubles  eye position (the nanmean(pos) of world.py:94)
 *   yaw_dev      [P][H] doubles  heading about +Z in radians (R.from_euler('Z', yaw), simulation.py:836, 2706),
 *                                or NULL when view_quat_dev is given
 *   view_quat_dev[P][H][4] doubles full eye orientation, or NULL for yaw-only views
 *   sun_dev      [2] or [P][H][2] doubles (theta_s = zenith distance, phi_s), NULL unless PEREZ
 *   sensor       photoreceptor transform, NULL unless out->resp is asked for
 * Global ray orientation = view orientation * eye-frame ommatidium orientation (examples/test_sky.py:25).
 * Views are ordered b = p*H + h.                                                                  */
size_t isy_render_workspace_bytes(const isy_world* w, const isy_eye* e, int64_t P, int64_t H);
int isy_render(const isy_world* w, const isy_eye* e, int64_t P, int64_t H, const double* pos_dev,
               const double* yaw_dev, const double* view_quat_dev, const double* sun_dev,
               const isy_sky_params* sky, const isy_sensor_params* sensor, uint32_t flags,
               const isy_outputs* out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Synchronises `stream` and returns the sticky device-side status of the renders that used this
 * workspace since the status was last read (ISY_OK, or ISY_ERR_OVERFLOW when a position of ANY of them
 * exceeded the slab); reading clears it.  A render never clears it.                                */
int isy_workspace_status(void* workspace_dev, void* stream);
/* Profiling aid: SM cycles summed over CTAs of the LAST render that used this workspace, spent in
 * out4[0] cull+projection, out4[1] bin build, out4[2] ray phase; out4[3] = positions processed.  */
int isy_workspace_phase_cycles(void* workspace_dev, void* stream, uint64_t* out4);

/* Same call with HOST buffers: stages inputs H2D, renders, copies the selected outputs D2H in
 * chunks overlapped with compute (pinned staging owned by the library) and returns after the
 * results are in host memory.  This is the end-to-end path bench.py times as `e2e`.            */
int isy_render_host(const isy_world* w, const isy_eye* e, int64_t P, int64_t H, const double* pos_host,
                    const double* yaw_host, const double* view_quat_host, const double* sun_host,
                    const isy_sky_params* sky, const isy_sensor_params* sensor, uint32_t flags,
                    const isy_outputs* out_host);

/* ---- direct replacements of the two reference callables (explicit ray orientations) ----------
 * WorldBase.__call__(pos, ori, init_c, ...) world.py:55-129: N rays given as global quaternions.
 *   init_c_host: N doubles or NULL (world.py:79-82);  rgb_host: N*3 doubles (float64 like the
 *   reference's init_c branch; the caller casts to float32 for the NaN-init branch)
 *   hit_host / depth_host: optional N int32 / N doubles                                           */
int isy_world_call_host(const isy_world* w, const double* pos3_host, const double* ray_quat_host, int64_t N,
                        const double* init_c_host, double* rgb_host, int32_t* hit_host, double* depth_host);

/* Sky.__call__(ori) / UniformSky.__call__(ori) sky.py:269-346, :47-97 before spectrum_influence:
 *   eta_host: optional N bytes, non-zero = occluded ray (sky.py:327-332; applied before the sun disc :334)
 *   Y/dop/aop_host: N doubles each; theta/phi_host: optional N doubles (sky.py:99-113 caches)     */
int isy_sky_call_host(const isy_sky_params* sky, double theta_s, double phi_s, const double* ray_quat_host,
                      int64_t N, const uint8_t* eta_host, int device, double* Y_host, double* dop_host,
                      double* aop_host, double* theta_host, double* phi_host);

/* spectrum_influence(v, irgbu) sky.py:671-703 over ONE call of n values (call-wide nansum /
 * nanmax reductions); irgbu: rows*5 doubles with n % rows == 0 (tiled like sky.py:689).
 *   channels_host: optional n*5 doubles (the function's return value)
 *   sum_host     : optional n doubles   (.sum(axis=1), what Sky.__call__ returns as Y, sky.py:344) */
int isy_spectrum_influence_host(const double* v_host, int64_t n, const double* irgbu_host, int64_t rows,
                                int device, double* sum_host, double* channels_host);

/* The same applied to a batch of luminance rows on the device, in place and per view -- what `Sky.__call__(ori, irgbu=...)`
 * does to the luminance of each call (sky.py:342-344: y = spectrum_influence(y, irgbu).sum(axis=1)) for every view of a
 * batched render: Y_dev[B][n] (float, or double with ISY_FLAG_OUT_F64), irgbu_dev rows*5 doubles on the device.
 * One CTA per view: the call-wide nansum / nanmax reductions of sky.py:692-695 become per-view reductions.  Asynchronous. */
int isy_spectrum_influence_views(void* Y_dev, int64_t B, int64_t n, const double* irgbu_dev, int64_t rows, uint32_t flags,
                                 int device, void* stream);

/* ---- familiarity of rendered views: the consumer of every view in the reference's heat-map and
 * parallel-lines runs (agent/agent.py:741-744 -> memory -> sim/simulation.py:3085-3092) ------------------------
 * A perfect memory stores M vectors of K values (the PN inputs of the route views, i.e. `resp` with one channel):
 *     familiarity(q) = 1 - min_m sqrt(mean_k (q_k - m_k)^2)
 * (the memory model is InvertPy's, not vendored by the reference: this definition is the library's own).
 * isy_memory_create copies the vectors to the device (the caller keeps ownership of vectors_host).              */
typedef struct isy_memory isy_memory;
int isy_memory_create(const float* vectors_host, int64_t M, int64_t K, int device, isy_memory** out);
int isy_memory_destroy(isy_memory* m);
int64_t isy_memory_size(const isy_memory* m);

/* Familiarity of Q query vectors on the device: queries_dev[q * row_stride + k], k < K; row_stride (in floats) and the
 * pointer must be multiples of 4 floats / 16 bytes (TMA).  familiarity_dev[Q] and / or nearest_dev[Q] (index of the
 * nearest stored vector) may be NULL.  A TF32 tensor-core pass (tcgen05, accumulators in TMEM) keeps four candidates
 * per query, an fp32 pass re-checks them exactly; the Q x M distance matrix is never written.  Asynchronous.      */
size_t isy_familiarity_workspace_bytes(const isy_memory* m, int64_t Q);
int isy_familiarity(const isy_memory* m, const float* queries_dev, int64_t Q, int64_t row_stride,
                    float* familiarity_dev, int32_t* nearest_dev, void* workspace_dev, size_t workspace_bytes,
                    void* stream);
/* Synchronises `stream` and returns in *count how many queries of the LAST isy_familiarity call on this workspace had
 * all four candidates inside the filter's error margin (the exact re-check then cannot rule out a fifth; csrc/isy_fam.cuh).
 * 0 on every workload measured so far; exposed so that a caller can see it.                                     */
int isy_familiarity_unresolved(const isy_memory* m, int64_t Q, void* workspace_dev, void* stream, int64_t* count);
/* Diagnostics: the four candidates of every query of the LAST isy_familiarity call on this workspace and their filter
 * scores v = |m|^2 - 2 q.(m - mean) as the tensor cores produced them (cand_host[Q][4], score_host[Q][4], ascending;
 * either may be NULL).  Synchronises `stream`.  Tests use it to check the filter's error model.                    */
int isy_familiarity_candidates(const isy_memory* m, int64_t Q, void* workspace_dev, void* stream, int32_t* cand_host,
                               float* score_host);
/* the same with host buffers ([Q][K] contiguous); blocking */
int isy_familiarity_host(const isy_memory* m, const float* queries_host, int64_t Q, float* familiarity_host,
                         int32_t* nearest_host);

/* ---- Willshaw-type mushroom-body memory: the default memory of the reference's visual-navigation agent
 * (agent/agent.py:560-575: #KC = 40 x #PN, sparseness 1 %; the model is InvertPy's WillshawNetwork, not vendored --
 * this library's definition is stated in csrc/isy_willshaw.cuh and invertsy_b200/familiarity.py) -------------------
 * nb_kc Kenyon cells each sum fan_in pseudo-randomly chosen PNs; the round(sparseness * nb_kc) cells with the largest
 * input fire; storing a view clears the output synapses of its firing cells; familiarity = 1 - fraction of the firing
 * cells whose synapse is still set.  Vectors are device pointers, [.][row_stride] floats, nb_pn used per row.        */
typedef struct isy_willshaw isy_willshaw;
int isy_willshaw_create(int64_t nb_pn, int64_t nb_kc, int32_t fan_in, double sparseness, uint32_t seed, int device,
                        isy_willshaw** out);
int isy_willshaw_destroy(isy_willshaw* m);
int isy_willshaw_reset(isy_willshaw* m);                  /* forget everything (all synapses back to 1)             */
int64_t isy_willshaw_active(const isy_willshaw* m);       /* number of cells that fire per view                     */
int isy_willshaw_train(isy_willshaw* m, const float* vectors_dev, int64_t M, int64_t row_stride, void* stream);
int isy_willshaw_familiarity(const isy_willshaw* m, const float* queries_dev, int64_t Q, int64_t row_stride,
                             float* familiarity_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ISY_B200_H_ */

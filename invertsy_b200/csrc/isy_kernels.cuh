// Kernels of the sm_100a view renderer (persistent, one CTA per position work item).
//
//   isy_project.cuh  phase A1 cull    : candidates from the world's xy cell grid, min-vertex distance in fp64,
//                                       reference order (world.py:94-98)                   -> visible list
//                    phase A2 project : vertex (theta, phi) by atan2, seam rule, reference crosses, normalised
//                                       fp32 edge functions (world.py:109-120, 482-526)    -> tile in shared memory
//   isy_raster.cuh   phase B1 (yaw-only views): every projected copy finds the (ray, heading) pairs it can
//                                       contain -- pitch-run walk over the eye's pitch-sorted rays for scans of
//                                       many headings, ray-grid walk for few -- and posts itself into
//                                       best[heading][ray] if it is nearer; fp32 filter + exact fp64 re-check
//   isy_bins.cuh     phase B1' (general orientations): per-position (theta, phi) bins, rays test their bin
//   isy_shade.cuh    phase B2 : winner -> colour / ground / background (world.py:79-90, :123), analytic sky
//                                       (sky.py:304-334), ONE write of the view
//   isy_render.cuh   the persistent kernel that strings the phases together
//
// argmin of distance over containing copies == the painter's far->near overwrite (world.py:106-123).
// ISY_FLAG_BRUTE_FORCE (and any position whose copies do not fit the shared-memory tile) streams every copy
// past every ray instead; results are identical.
#pragma once
#include "isy_common.cuh"

#ifndef ISY_THREADS
#define ISY_THREADS 512
#endif

#include "isy_layout.cuh"
#include "isy_setup.cuh"
#include "isy_project.cuh"
#include "isy_bins.cuh"
#include "isy_raster.cuh"
#include "isy_shade.cuh"
#include "isy_render.cuh"
#include "isy_sky_kernels.cuh"

// Kernels of the sm_100a view renderer (persistent, one CTA per position work item).
//
//   isy_project.cuh  phase A1 cull    : candidates from the world's xy cell grid, min-vertex distance in fp64,
//                                       reference order (world.py:94-98)                   -> visible list
//                    phase A2 project : vertex (theta, phi) by atan2, seam rule, reference crosses, normalised
//                                       fp32 edge functions (world.py:109-120, 482-526)    -> tile in shared memory
//   isy_raster.cuh   phase B1 (yaw-only views): which projected copy contains each (ray, heading), nearest first.
//                                       Scans with evenly spaced headings: the copies are filed into tiles of 64
//                                       pitch-sorted rays x one azimuth column and a warp per tile tests its list,
//                                       winners in registers (tile walk; column walk if the lists do not fit);
//                                       irregular headings: pitch-run walk, a warp per copy posting into
//                                       best[heading][ray]; few headings: ray-grid walk.  fp32 filter + exact fp64
//                                       re-check everywhere
//   isy_bins.cuh     phase B1' (general orientations): per-position (theta, phi) bins, rays test their bin
//   isy_shade.cuh    phase B2 : winner -> colour / ground / background (world.py:79-90, :123), analytic sky
//                                       (sky.py:304-334), ONE write of the view; or, in rows mode (small eyes,
//                                       sky-less / uniform-sky / shared-heading calls): template and sky-table rows
//                                       streamed into the views while the copies are rasterised, then patches
//                                       for the rays that hit something
//   isy_render.cuh   the persistent kernel that strings the phases together
//
// argmin of distance over containing copies == the painter's far->near overwrite (world.py:106-123).
// ISY_FLAG_BRUTE_FORCE (and any position whose copies do not fit the shared-memory tile) streams every copy
// past every ray instead; results are identical.
#pragma once
#include "isy_common.cuh"

// CTA size.  One persistent CTA per SM either way (185-220 KB of shared memory); 896 threads = 28 warps at 72 registers.
// The rasteriser is latency-bound and scales with the warps (121 k -> 90 k cycles per position from 512 to 896
// threads), the final pass is throughput-bound and pays a little for the tighter register budget (73 k -> 81 k):
// 31.5 -> 27.9 ms per step on the benchmark grid.  640 / 768 / 1024 threads were measured too (README of profiles/).
#ifndef ISY_THREADS
#define ISY_THREADS 896
#endif

#include "isy_layout.cuh"
#include "isy_setup.cuh"
#include "isy_project.cuh"
#include "isy_bins.cuh"
#include "isy_raster.cuh"
#include "isy_shade.cuh"
#include "isy_render.cuh"
#include "isy_sky_kernels.cuh"

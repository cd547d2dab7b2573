// Familiarity of rendered views against a perfect memory of stored views:
//     familiarity(q) = 1 - min_m RMSE(q, m),   RMSE(q, m) = sqrt(mean_k (q_k - m_k)^2)
// (the reference's consumer of every rendered view: agent/agent.py:741-744 -> memory -> `familiarity`,
// sim/simulation.py:3085-3092; the memory model itself is InvertPy's, unpinned upstream -- the definition above
// is this library's, stated in invertsy_b200/familiarity.py).
//
// Q queries x M stored vectors x K ommatidia is a dense contraction (1.44 M x 812 x 1000 on the heat-map grid), the
// one place of this library where the tensor cores belong:
//
//   fam_filter_kernel   TF32 tcgen05.mma as the *filter*: for a tile of 128 queries the scores
//                           v(q, m) = |m|^2 - 2 q . (m - mean)          (= |q - m|^2 up to a per-query constant)
//                       are accumulated in TMEM (fp32), 256 stored vectors per pass, operands staged by TMA
//                       (128-byte swizzle, 4-stage mbarrier pipeline), accumulators double-buffered so that the
//                       epilogue warps scan one pass (tcgen05.ld, running top-4 per query) while the next is
//                       multiplied.  The Q x M score matrix never exists in HBM: 16 bytes per query leave the SM.
//                       Centring the stored vectors on their mean shrinks the TF32 rounding of the products
//                       (|m - mean| << |m| for views of one route) without changing the arg-min.
//   fam_recheck_kernel  the exact re-check: fp32 sum of squared differences of each query with its four
//                       candidates, one warp per query; familiarity and nearest index from the smallest.
//
// Roles in fam_filter_kernel (256 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane),
// warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM lane = query row of the tile; warp w reads lanes 32 (w % 4) ..).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "isy_common.cuh"

namespace isy {
namespace fam {

constexpr int kBM = 128;       // queries per tile = UMMA M = TMEM lanes
constexpr int kBK = 32;        // floats per k-block: 128 bytes = one row of the 128-byte swizzle
constexpr int kNB = 256;       // stored vectors per pass = UMMA N (max) = TMEM columns of one accumulator
constexpr int kStages = 4;     // smem pipeline depth
constexpr int kAccBufs = 2;    // accumulator buffers in TMEM (2 x 256 columns = all 512)
constexpr int kThreads = 256;
constexpr int kTopK = 4;       // candidates per query handed to the exact re-check
constexpr uint32_t kStageABytes = kBM * kBK * 4, kStageBBytes = kNB * kBK * 4;
constexpr size_t kSmemBytes = 1024 + (size_t)kStages * (kStageABytes + kStageBBytes) + 256;

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// waits for the phase with the given parity to complete; a pipeline that never completes traps instead of hanging
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 operands, fp32 accumulate; issued by ONE thread for the CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the mbarrier gets one arrival when every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor of a [rows x 32 floats] K-major tile written by TMA with the 128-byte swizzle:
// rows are 128 bytes apart, groups of 8 rows (one swizzle atom) 1024 bytes apart (SBO); LBO is unused for this
// layout; descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.  The tile starts on a 1024-byte boundary.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major (bits 15, 16 = 0),
// N >> 3 in bits 17-22, M >> 4 in bits 24-28
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// running top-4 (smallest scores) of one query
struct Top4 {
    float v[kTopK];
    int i[kTopK];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int k = 0; k < kTopK; ++k) v[k] = __int_as_float(0x7f800000), i[k] = -1;
    }
    __device__ __forceinline__ void push(float x, int n) {
        if (!(x < v[kTopK - 1])) return;
        v[kTopK - 1] = x, i[kTopK - 1] = n;     // replaces the worst, then bubbles up (ties keep the earlier index in front)
#pragma unroll
        for (int k = kTopK - 1; k > 0; --k) {
            const bool up = v[k] < v[k - 1];
            const float va = v[k], vb = v[k - 1];
            const int ia = i[k], ib = i[k - 1];
            v[k - 1] = up ? va : vb, v[k] = up ? vb : va;
            i[k - 1] = up ? ia : ib, i[k] = up ? ib : ia;
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// tmA: queries  [Q][K] floats, box 32 x 128;  tmB: centred stored vectors [M][K] floats, box 32 x 256; both with
// the 128-byte swizzle and zero fill outside the tensors (ragged Q, M, K cost nothing and contribute zeros).
// cn[m] = |m|^2 of the UNcentred stored vectors.  cand[q] / score[q] = the four smallest scores, ascending, and their
// stored vectors (-1 = none).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
fam_filter_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const float* __restrict__ cn, int Q, int M, int K, int4* __restrict__ cand, float4* __restrict__ score) {
    extern __shared__ uint8_t fam_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fam_smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sA = smem_u32(smem);                                   // [kStages][128 x 32] floats
    const uint32_t sB = sA + kStages * kStageABytes;                      // [kStages][256 x 32] floats
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * (kStageABytes + kStageBBytes));
    const uint32_t bar0 = smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (kStages + s); };
    auto acc_full = [&](int b) { return bar0 + 8u * (2 * kStages + b); };
    auto acc_empty = [&](int b) { return bar0 + 8u * (2 * kStages + kAccBufs + b); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAccBufs);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (Q + kBM - 1) / kBM, npass = (M + kNB - 1) / kNB, nkb = (K + kBK - 1) / kBK;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(full(s), 1), mbar_init(empty(s), 1);
        for (int b = 0; b < kAccBufs; ++b) mbar_init(acc_full(b), 1), mbar_init(acc_empty(b), 4 * 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // all 512 TMEM columns (one CTA per SM: 193 KB of shared memory)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ---- TMA producer ----
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int pass = 0; pass < npass; ++pass)
                    for (int kb = 0; kb < nkb; ++kb, ++it) {
                        const int s = it % kStages;
                        mbar_wait(empty(s), ((it / kStages) & 1) ^ 1);       // slot free (passes at once the first time round)
                        mbar_expect_tx(full(s), kStageABytes + kStageBBytes);
                        tma_load_2d(sA + s * kStageABytes, &tmA, kb * kBK, tile * kBM, full(s));
                        tma_load_2d(sB + s * kStageBBytes, &tmB, kb * kBK, pass * kNB, full(s));
                    }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ---- MMA issuer ----
            uint32_t it = 0, acc_it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int pass = 0; pass < npass; ++pass, ++acc_it) {
                    const int ab = acc_it % kAccBufs;
                    mbar_wait(acc_empty(ab), ((acc_it / kAccBufs) & 1) ^ 1);   // the epilogue has drained this accumulator
                    tc_fence_after();
                    const int ncols = min(kNB, ((M - pass * kNB) + 15) & ~15);  // UMMA N: a multiple of 16
                    const uint32_t idesc = umma_idesc_tf32(kBM, ncols);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(ab * kNB);
                    for (int kb = 0; kb < nkb; ++kb, ++it) {
                        const int s = it % kStages;
                        mbar_wait(full(s), (it / kStages) & 1);                // TMA has landed this stage
                        tc_fence_after();
                        const uint64_t ad = umma_desc_sw128(sA + s * kStageABytes), bd = umma_desc_sw128(sB + s * kStageBBytes);
#pragma unroll
                        for (int k = 0; k < kBK / 8; ++k)                      // UMMA K = 8 TF32 = 32 bytes along the swizzled row
                            umma_tf32(d_tmem, ad + 2u * k, bd + 2u * k, idesc, (kb | k) != 0);
                        umma_commit(empty(s));                                 // frees the slot when these MMAs are done
                    }
                    umma_commit(acc_full(ab));                                 // accumulator complete
                }
        }
    } else if (warp >= 4) {   // ---- epilogue: running top-4 per query ----
        uint32_t acc_it = 0;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            Top4 top;
            top.init();
            for (int pass = 0; pass < npass; ++pass, ++acc_it) {
                const int ab = acc_it % kAccBufs;
                mbar_wait(acc_full(ab), (acc_it / kAccBufs) & 1);
                tc_fence_after();
                const int n0 = pass * kNB, ncols = min(kNB, M - n0);
                for (int c0 = 0; c0 < ncols; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + lane_base + (uint32_t)(ab * kNB + c0), r);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = n0 + c0 + j;
                        if (n < M) top.push(fmaf(-2.0f, __uint_as_float(r[j]), __ldg(cn + n)), n);
                    }
                }
                tc_fence_before();
                mbar_arrive(acc_empty(ab));
            }
            const int q = tile * kBM + (warp & 3) * 32 + lane;
            if (q < Q) {
                ISY_BOUND(top.i[0] < M && top.i[1] < M && top.i[2] < M && top.i[3] < M && top.i[0] >= 0);
                cand[q] = make_int4(top.i[0], top.i[1], top.i[2], top.i[3]);
                score[q] = make_float4(top.v[0], top.v[1], top.v[2], top.v[3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// exact re-check: one warp per query; fp32 sum of squared differences with each candidate (lanes stride the
// vector 16 bytes at a time), the smallest wins (ties: lower index).  familiarity = 1 - sqrt(d2 / K).
//
// Which candidates need it: the filter's score of a stored vector is off by at most
//     2 * (2^-10 |q|_2 |m - mean|_2  +  K 2^-24 sum_k |q_k (m - mean)_k|)  <=  2.2 * 2^-10 |q|_2 mc_max     (K < 2^12)
// (TF32 operands rounded to nearest: 2^-11 relative each; Cauchy-Schwarz; fp32 accumulation), so a candidate whose
// score exceeds the best by more than twice that cannot be the nearest and is not loaded.  |q|_2 is formed here from
// the query the warp reads anyway; mc_max = max_m |m - mean|_2 comes from isy_memory_create.  If even the fourth
// candidate is inside the margin a fifth might be too: such queries are counted in *unresolved (none on the heat-map
// grid; the count is part of the result so that a caller can see it).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fam_recheck_kernel(const float* __restrict__ queries, long long ldq, const float* __restrict__ stored, long long lds, int Q, int K,
                   const int4* __restrict__ cand, const float4* __restrict__ score, float mc_max, float* __restrict__ fam_out,
                   int* __restrict__ nearest_out, const int* __restrict__ orig, unsigned int* __restrict__ unresolved) {
    const int lane = threadIdx.x & 31;
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q) return;
    const int4 c4 = cand[q];
    const float4 v4 = score[q];
    const int c[kTopK] = {c4.x, c4.y, c4.z, c4.w};
    const float v[kTopK] = {v4.x, v4.y, v4.z, v4.w};
    const float* x = queries + q * ldq;
    const bool vec = (K & 3) == 0 && (ldq & 3) == 0 && (lds & 3) == 0 && ((uintptr_t)queries & 15) == 0 && ((uintptr_t)stored & 15) == 0;
    float best = __int_as_float(0x7f800000), xx = 0.f, margin = __int_as_float(0x7f800000);
    int best_i = -1;
    // Vectors of up to 1024 floats (the usual eye): the query is read ONCE, eight 16-byte loads per lane issued
    // together (the kernel streams the queries from HBM: what bounds it is how many bytes a warp has in flight), and kept
    // in registers for all its candidates.  The sums run over the same elements in the same order as the general path.
    constexpr int kQV = 8;
    const int K4 = K >> 2;
    const bool in_regs = vec && K4 <= 32 * kQV;
    float4 xa[kQV];
    if (in_regs) {
#pragma unroll
        for (int u = 0; u < kQV; ++u) {
            const int k = lane + 32 * u;
            xa[u] = k < K4 ? __ldg(reinterpret_cast<const float4*>(x) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int t = 0; t < kTopK; ++t) {
        const int m = c[t];
        if (m < 0 || v[t] > v[0] + margin) continue;           // (warp-uniform: the candidates are the query's)
        const float* y = stored + (size_t)m * lds;
        float s0 = 0.f, s1 = 0.f, n0 = 0.f;
        if (in_regs) {
            float4 yb[kQV];
#pragma unroll
            for (int u = 0; u < kQV; ++u) {
                const int k = lane + 32 * u;
                yb[u] = k < K4 ? __ldg(reinterpret_cast<const float4*>(y) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kQV; ++u) {      // (elements beyond K are zeros in both: they leave the sums as they are)
                const float4 a = xa[u], b = yb[u];
                const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
                s0 = fmaf(d0, d0, s0), s1 = fmaf(d1, d1, s1), s0 = fmaf(d2, d2, s0), s1 = fmaf(d3, d3, s1);
                if (t == 0) n0 = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, n0))));
            }
        } else if (vec) {
            const float4* x4 = reinterpret_cast<const float4*>(x);
            const float4* y4 = reinterpret_cast<const float4*>(y);
            for (int k = lane; k < (K >> 2); k += 32) {
                const float4 a = __ldg(x4 + k), b = __ldg(y4 + k);
                const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
                s0 = fmaf(d0, d0, s0), s1 = fmaf(d1, d1, s1), s0 = fmaf(d2, d2, s0), s1 = fmaf(d3, d3, s1);
                if (t == 0) n0 = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, n0))));
            }
        } else {
            for (int k = lane; k < K; k += 32) {
                const float a = __ldg(x + k), d = a - __ldg(y + k);
                s0 = fmaf(d, d, s0);
                if (t == 0) n0 = fmaf(a, a, n0);
            }
        }
        float s = s0 + s1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (t == 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) n0 += __shfl_xor_sync(0xffffffffu, n0, o);
            xx = n0;
            margin = 2.0f * 2.2f * 9.765625e-4f * sqrtf(xx) * mc_max;
        }
        if (s < best || (s == best && m < best_i)) best = s, best_i = m;
    }
    if (lane == 0) {
        if (fam_out) fam_out[q] = 1.0f - sqrtf(best / (float)K);
        if (nearest_out) nearest_out[q] = best_i >= 0 ? __ldg(orig + best_i) : -1;   // index as the caller stored it
        if (unresolved && c[kTopK - 1] >= 0 && v[kTopK - 1] <= v[0] + margin) atomicAdd(unresolved, 1u);
    }
}

}  // namespace fam
}  // namespace isy

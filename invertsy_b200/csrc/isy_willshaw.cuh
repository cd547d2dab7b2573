// Willshaw-type mushroom-body memory: the default memory of the reference's visual-navigation
// agent (agent/agent.py:560-575: "#KC equal to 40 x #PN, sparseness is 1 %"); the model itself is InvertPy's
// (invertpy.brain.memory.WillshawNetwork), not vendored by the reference: **parity unpinned**.  What is implemented is
// stated in full in invertsy_b200/familiarity.py (`WillshawMemory`) and restated in NumPy by its tests:
//
//   PN -> KC   every Kenyon cell j sums `fan_in` projection neurons, input i being PN  umulhi(hash32(j * fan_in + i + seed),
//              nb_pn)  (a fixed pseudo-random wiring generated on the fly: a table would be 640 KB read once per view);
//              the sum runs in fp32 in input order;
//   KC         the k = round(sparseness * nb_kc) cells with the largest input fire (ties: the lower cell index);
//   KC -> MBON every synapse starts at 1; storing a view sets the synapses of its firing cells to 0;
//   familiarity(view) = 1 - (number of firing cells whose synapse is still 1) / k.
//
// One CTA per view at a time: the PN vector and the 40 000 KC inputs live in shared memory (as order-preserving
// 32-bit keys), the k-th largest key comes from a three-pass radix select over a shared-memory histogram.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "isy_common.cuh"

namespace isy {
namespace mb {

constexpr int kThreads = 512;
constexpr int kBins = 2048;   // 11 + 11 + 10 bits

__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {   // "lowbias32" integer hash
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t float_key(float f) {             // larger float <=> larger key
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// block-wide exclusive prefix sum of one value per thread (kThreads threads); returns the total in `total`
__device__ __forceinline__ unsigned int block_exclusive(unsigned int v, unsigned int* warp_sums, unsigned int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    __syncthreads();
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    unsigned int base = 0;
    total = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
        const unsigned int s = warp_sums[w];
        if (w < warp) base += s;
        total += s;
    }
    return base + inc - v;
}

// kTrain: the firing cells of each view have their synapse cleared in `mask` (bit j = synapse of cell j still 1);
// otherwise fam[q] = 1 - (firing cells with the bit set) / k.
template <bool kTrain>
__global__ void __launch_bounds__(kThreads, 1)
willshaw_kernel(const float* __restrict__ views, long long ld, int Q, int nb_pn, int nb_kc, int fan_in, int k, uint32_t seed,
                unsigned int* mask, float* __restrict__ fam) {
    extern __shared__ unsigned char mb_smem[];
    float* pn = reinterpret_cast<float*>(mb_smem);                                  // [nb_pn]
    uint32_t* key = reinterpret_cast<uint32_t*>(pn + ((nb_pn + 3) & ~3));           // [nb_kc]
    unsigned int* hist = key + nb_kc;                                               // [kBins]
    __shared__ unsigned int s_warp[kThreads / 32];
    __shared__ unsigned int s_prefix, s_want, s_count;
    const int tid = threadIdx.x;
    for (int q = blockIdx.x; q < Q; q += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < nb_pn; i += kThreads) pn[i] = views[(size_t)q * ld + i];
        __syncthreads();
        // KC inputs
        for (int j = tid; j < nb_kc; j += kThreads) {
            float s = 0.f;
            uint32_t h = (uint32_t)j * (uint32_t)fan_in + seed;
            for (int i = 0; i < fan_in; ++i) {
                ISY_BOUND(__umulhi(hash32(h + (uint32_t)i), (uint32_t)nb_pn) < (uint32_t)nb_pn);
                s += pn[__umulhi(hash32(h + (uint32_t)i), (uint32_t)nb_pn)];
            }
            key[j] = float_key(s);
        }
        // radix select of the k-th largest key: 11, 11, 10 bits from the top
        uint32_t prefix = 0, pmask = 0;
        unsigned int want = (unsigned int)k;      // rank (1 = largest) still to find among the keys matching the prefix
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
            const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0), bits = pass == 2 ? 10 : 11;
            const uint32_t nb = 1u << bits;
            __syncthreads();
            for (int i = tid; i < kBins; i += kThreads) hist[i] = 0;
            __syncthreads();
            for (int j = tid; j < nb_kc; j += kThreads) {
                const uint32_t kk = key[j];
                if ((kk & pmask) == prefix) atomicAdd(&hist[(kk >> shift) & (nb - 1)], 1u);
            }
            __syncthreads();
            if (tid < 32) {   // walk the bins from the top until `want` keys are covered: one warp, 64 bins per lane
                const int per = kBins / 32;
                unsigned int sum = 0;
                for (int i = 0; i < per; ++i) sum += hist[kBins - 1 - (tid * per + i)];
                unsigned int inc = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned int u = __shfl_up_sync(0xffffffffu, inc, d);
                    if (tid >= d) inc += u;
                }
                const unsigned int before = inc - sum;
                if (before < want && want <= inc) {   // the bin is in this lane's range
                    unsigned int run = before;
                    for (int i = 0; i < per; ++i) {
                        const int b = kBins - 1 - (tid * per + i);
                        ISY_BOUND(b >= 0 && b < kBins);
                        if (run + hist[b] >= want) {
                            s_prefix = (uint32_t)b, s_want = want - run;
                            break;
                        }
                        run += hist[b];
                    }
                }
            }
            __syncthreads();
            prefix |= s_prefix << shift;
            pmask |= (nb - 1) << shift;
            want = s_want;
        }
        // threshold key = prefix; cells above it fire, and the first `want` cells AT it in index order
        const uint32_t thr = prefix;
        const int per = (nb_kc + kThreads - 1) / kThreads;        // contiguous chunk of cells per thread: index order
        const int j0 = tid * per, j1 = min(nb_kc, j0 + per);
        unsigned int ties = 0;
        for (int j = j0; j < j1; ++j) ties += key[j] == thr;
        unsigned int total;
        unsigned int rank = block_exclusive(ties, s_warp, total);   // ties before this thread's chunk
        if (tid == 0) s_count = 0;
        __syncthreads();
        unsigned int novel = 0;
        for (int j = j0; j < j1; ++j) {
            const uint32_t kk = key[j];
            bool fire = kk > thr;
            if (kk == thr) fire = rank++ < want;
            if (!fire) continue;
            if (kTrain) atomicAnd(&mask[j >> 5], ~(1u << (j & 31)));
            else novel += (mask[j >> 5] >> (j & 31)) & 1u;
        }
        if (!kTrain) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) novel += __shfl_xor_sync(0xffffffffu, novel, o);
            if ((tid & 31) == 0 && novel) atomicAdd(&s_count, novel);
            __syncthreads();
            if (tid == 0) fam[q] = 1.0f - (float)s_count / (float)k;
        }
    }
}

}  // namespace mb
}  // namespace isy

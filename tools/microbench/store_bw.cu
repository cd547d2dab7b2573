// Per-SM global store throughput on sm_100a: what bounds a kernel whose CTAs each stream ~1 MB bursts to HBM?
// Variants: plain / streaming (.cs) stores of 4 and 16 bytes per lane, and TMA bulk copies shared -> global.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_bw store_bw.cu && ./store_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kThreads = 896;

template <int kVariant>
__global__ void __launch_bounds__(kThreads, 1) store_kernel(float* out, size_t bytes_per_cta, int chunk, unsigned long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = kThreads / 32;
    for (int i = threadIdx.x; i < chunk / 4; i += kThreads) reinterpret_cast<float*>(smem)[i] = (float)i;
    __syncthreads();
    unsigned char* base = reinterpret_cast<unsigned char*>(out) + (size_t)blockIdx.x * bytes_per_cta;
    const long long t0 = clock64();
    if (kVariant == 5 || kVariant == 6) {
        // 5: 4 B per lane, every warp store starting 32 B into a 128-byte line (rows of 1000 floats are like that)
        // 6: the colour pattern -- three stores of 4 B per lane at a 12-byte stride (96 floats per warp trip)
        const size_t per_warp_trip = kVariant == 5 ? 128 : 384;
        for (size_t off = (size_t)warp * per_warp_trip; off + per_warp_trip + 32 <= bytes_per_cta; off += (size_t)warps * per_warp_trip) {
            if (kVariant == 5) __stcs(reinterpret_cast<float*>(base + 32 + off) + lane, 1.f);
            else {
                float* q = reinterpret_cast<float*>(base + off) + 3 * lane;
                __stcs(q, 1.f), __stcs(q + 1, 2.f), __stcs(q + 2, 3.f);
            }
        }
    } else if (kVariant < 4) {
        constexpr int kB = (kVariant < 2) ? 16 : 4;                 // bytes per lane
        const size_t per_warp_trip = 32 * kB;
        for (size_t off = (size_t)warp * per_warp_trip; off < bytes_per_cta; off += (size_t)warps * per_warp_trip) {
            unsigned char* p = base + off + lane * kB;
            if (kVariant == 0) *reinterpret_cast<float4*>(p) = make_float4(1.f, 2.f, 3.f, 4.f);
            if (kVariant == 1) __stcs(reinterpret_cast<float4*>(p), make_float4(1.f, 2.f, 3.f, 4.f));
            if (kVariant == 2) *reinterpret_cast<float*>(p) = 1.f;
            if (kVariant == 3) __stcs(reinterpret_cast<float*>(p), 1.f);
        }
    } else {
        // TMA bulk stores: lane 0 of every warp issues chunk-byte copies from (constant) shared memory
        if (lane == 0) {
            const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
            for (size_t off = (size_t)warp * chunk; off + chunk <= bytes_per_cta; off += (size_t)warps * chunk) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off), "r"(s), "r"(chunk) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int kVariant>
void run(const char* name, float* buf, size_t bytes_per_cta, int ctas, int chunk, unsigned long long* d_cycles) {
    cudaFuncSetAttribute(store_kernel<kVariant>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    store_kernel<kVariant><<<ctas, kThreads, 64 * 1024>>>(buf, bytes_per_cta, chunk, d_cycles);
    cudaEventRecord(a);
    store_kernel<kVariant><<<ctas, kThreads, 64 * 1024>>>(buf, bytes_per_cta, chunk, d_cycles);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    unsigned long long h[148];
    cudaMemcpy(h, d_cycles, sizeof(unsigned long long) * ctas, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < ctas; ++i) mean += (double)h[i];
    mean /= ctas;
    printf("%-34s ctas=%3d chunk=%5d  %7.3f ms  %7.1f GB/s  %6.2f B/clk/SM  (%s)\n", name, ctas, chunk, ms,
           (double)bytes_per_cta * ctas / ms / 1e6, (double)bytes_per_cta / mean, cudaGetErrorString(e));
}

int main() {
    const size_t bytes_per_cta = 64u << 20;
    float* buf;
    unsigned long long* d_cycles;
    cudaMalloc(&buf, bytes_per_cta * 148);
    cudaMalloc(&d_cycles, 148 * sizeof(unsigned long long));
    for (int ctas : {148, 1}) {
        run<0>("st.global.v4 (16 B/lane)", buf, bytes_per_cta, ctas, 4096, d_cycles);
        run<1>("st.global.cs.v4 (16 B/lane, streaming)", buf, bytes_per_cta, ctas, 4096, d_cycles);
        run<2>("st.global.b32 (4 B/lane)", buf, bytes_per_cta, ctas, 4096, d_cycles);
        run<3>("st.global.cs.b32 (4 B/lane, streaming)", buf, bytes_per_cta, ctas, 4096, d_cycles);
        run<4>("TMA bulk smem->global", buf, bytes_per_cta, ctas, 4096, d_cycles);
        run<4>("TMA bulk smem->global", buf, bytes_per_cta, ctas, 16384, d_cycles);
        run<5>("st.global.cs.b32, 32 B off the line", buf, bytes_per_cta, ctas, 4096, d_cycles);
        run<6>("3 x st.global.cs.b32, 12-byte stride", buf, bytes_per_cta, ctas, 4096, d_cycles);
    }
    return 0;
}

// umma_probe.cu -- microbenchmark of tcgen05.mma issue / commit behaviour on one SM (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I skillshot_learning_b200/csrc -o /tmp/umma_probe tools/umma_probe.cu
// Prints cycles for NM back-to-back MMAs (M=128, N=n, K=16 each) with a commit every G MMAs.
#include <cstdio>
#include "ss_tc_common.cuh"
using namespace sstc;

template <int N, int G>
__global__ void probe(long long *out) {
    constexpr int NM = 64;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + 160 * 1024, tptr = bars + 64 * 8;
    for (uint32_t o = threadIdx.x * 16; o < 160 * 1024; o += blockDim.x * 16) *reinterpret_cast<uint4 *>(smem + o) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(bars + i * 8, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(tptr, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + 160 * 1024 + 64 * 8);
    if (threadIdx.x == 0) {
        const uint64_t ad = desc_kmajor(sbase, CHUNK_A), bd = desc_kmajor(sbase + 64 * 1024, N * 16);
        constexpr uint32_t idesc = umma_idesc(128, N);
        for (int rep = 0; rep < 3; ++rep) {
            long long t0 = clock64();
#pragma unroll
            for (int i = 0; i < NM; ++i) {
                umma_bf16(tmem, desc_advance(ad, (i & 15) * 4096), desc_advance(bd, (i & 7) * 2 * N * 16), idesc, (i % G) > 0);
                if ((i + 1) % G == 0) umma_commit(bars + ((i / G) & 63) * 8);
            }
            long long t1 = clock64();
            mbar_wait(bars + ((NM / G - 1) & 63) * 8, rep & 1);
            long long t2 = clock64();
            if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
            for (int b = 0; b < NM / G - 1; ++b) mbar_wait(bars + b * 8, rep & 1);
        }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int G>
void run() {
    long long *out; cudaMalloc(&out, 16);
    cudaFuncSetAttribute(probe<N, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    probe<N, G><<<1, 128, 200 * 1024>>>(out);
    long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d  64 mmas, commit every %2d : issue %6lld cyc, complete %6lld cyc  (%.1f cyc/mma; math floor %d)  %s\n",
           N, G, h[0], h[1], (double)h[1] / 64, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out);
}

// alternating instruction shapes / accumulators: MODE 0 = 4 x N128 into D2 then 2 x N64 into D1 (the block pipeline's
// issue pattern); MODE 1 = the same MMAs grouped (all N128 first, then all N64); MODE 2 = as 0 but N64 replaced by N128
template <int MODE>
__global__ void probe_mix(long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + 160 * 1024, tptr = bars + 64 * 8;
    for (uint32_t o = threadIdx.x * 16; o < 160 * 1024; o += blockDim.x * 16) *reinterpret_cast<uint4 *>(smem + o) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(bars + i * 8, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(tptr, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + 160 * 1024 + 64 * 8);
    if (threadIdx.x == 0) {
        const uint64_t ad = desc_kmajor(sbase, CHUNK_A), bd = desc_kmajor(sbase + 64 * 1024, 128 * 16), bd64 = desc_kmajor(sbase + 100 * 1024, 256 * 16);
        constexpr uint32_t i128 = umma_idesc(128, 128), i64 = umma_idesc(128, 64);
        for (int rep = 0; rep < 3; ++rep) {
            long long t0 = clock64();
            if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 64; ++i) umma_bf16(tmem + 256, desc_advance(ad, (i & 15) * 4096), desc_advance(bd, (i & 7) * 4096), i128, (i & 3) > 0);
#pragma unroll
                for (int i = 0; i < 32; ++i) umma_bf16(tmem + (i & 3) * 64, desc_advance(ad, (i & 1) * 4096), desc_advance(bd64, (i & 1) * 8192), i64, i & 1);
            } else {
#pragma unroll
                for (int b = 0; b < 16; ++b) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem + 256, desc_advance(ad, ((b * 4 + k) & 15) * 4096), desc_advance(bd, k * 4096), i128, k > 0);
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        umma_bf16(tmem + (b & 3) * 64, desc_advance(ad, k * 4096), desc_advance(MODE == 2 ? bd : bd64, k * 8192), MODE == 2 ? i128 : i64, k);
                }
            }
            umma_commit(bars);
            long long t1 = clock64();
            mbar_wait(bars, rep & 1);
            long long t2 = clock64();
            if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
template <int MODE>
void run_mix(const char *what) {
    long long *out; cudaMalloc(&out, 16);
    cudaFuncSetAttribute(probe_mix<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    probe_mix<MODE><<<1, 128, 200 * 1024>>>(out);
    long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-60s issue %6lld cyc, complete %6lld cyc  %s\n", what, h[0], h[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out);
}

// latency of a short MMA burst on an idle pipe: issue NM MMAs (N columns), commit, wait; plus the cost of a wait on an
// already-completed barrier and of the tcgen05 fences
template <int N, int NM>
__global__ void probe_lat(long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + 160 * 1024, tptr = bars + 64 * 8;
    for (uint32_t o = threadIdx.x * 16; o < 160 * 1024; o += blockDim.x * 16) *reinterpret_cast<uint4 *>(smem + o) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(bars + i * 8, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(tptr, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + 160 * 1024 + 64 * 8);
    if (threadIdx.x == 0) {
        const uint64_t ad = desc_kmajor(sbase, CHUNK_A), bd = desc_kmajor(sbase + 64 * 1024, N * 16);
        constexpr uint32_t idesc = umma_idesc(128, N);
        long long acc[5] = {0, 0, 0, 0, 0};
        for (int rep = 0; rep < 8; ++rep) {
            long long t0 = clock64();
#pragma unroll
            for (int i = 0; i < NM; ++i) umma_bf16(tmem, desc_advance(ad, i * 4096), desc_advance(bd, i * 2 * N * 16), idesc, i > 0);
            long long t1 = clock64();
            umma_commit(bars);
            long long t2 = clock64();
            mbar_wait(bars, rep & 1);
            long long t3 = clock64();
            mbar_wait(bars, rep & 1);            // already complete
            long long t4 = clock64();
            tc_fence_after();
            long long t5 = clock64();
            if (rep >= 4) { acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3; acc[4] += t5 - t4; }
        }
        for (int k = 0; k < 5; ++k) out[k] = acc[k] / 4;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
template <int N, int NM>
void run_lat() {
    long long *out; cudaMalloc(&out, 64);
    cudaFuncSetAttribute(probe_lat<N, NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    probe_lat<N, NM><<<1, 128, 200 * 1024>>>(out);
    long long h[5] = {0};
    cudaMemcpy(h, out, 40, cudaMemcpyDeviceToHost);
    printf("idle pipe, %2d x N=%3d: issue %5lld  commit %4lld  wait-until-retired %5lld  wait-on-complete %4lld  fence::after %4lld cycles\n",
           NM, N, h[0], h[1], h[2], h[3], h[4]);
    cudaFree(out);
}

int main() {
    run_lat<256, 1>(); run_lat<256, 2>(); run_lat<128, 1>(); run_lat<128, 4>(); run_lat<128, 17>();
    return 0;
}

// umma_probe3.cu -- what slows a stream of 128 x 256 x 16 tcgen05.mma when other units work beside it?  (development tool)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I skillshot_learning_b200/csrc -o tools/bin/umma_probe3 tools/umma_probe3.cu
// 64 back-to-back MMAs from thread 0 while  MODE & 1: another thread streams 16 KB bulk copies global -> shared memory,
// MODE & 2: three warps read the other 256 tensor-memory columns with tcgen05.ld,  MODE & 4: four warps store 16 bytes per
// thread to global memory in a loop.
#include <cstdio>
#include "ss_tc_common.cuh"
using namespace sstc;

__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

template <int MODE>
__global__ void probe(long long *out, const uint8_t *gsrc, uint8_t *gdst) {
    constexpr int NM = 64, N = 256;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + 216 * 1024, tptr = bars + 64;
    volatile int *stop = reinterpret_cast<volatile int *>(smem + 216 * 1024 + 128);
    for (uint32_t o = threadIdx.x * 16; o < 216 * 1024; o += blockDim.x * 16) {
        uint32_t h = o * 2654435761u;
        *reinterpret_cast<uint4 *>(smem + o) = make_uint4((h ^ 0x1234567u) & 0x3FFF3FFFu, (h * 7u) & 0x3FFF3FFFu, (h * 13u) & 0x3FFF3FFFu, (h * 29u) & 0x3FFF3FFFu);
    }
    if (threadIdx.x == 0) { mbar_init(bars, 1); mbar_init(bars + 8, 1); mbar_fence_init(); *stop = 0; }
    if (threadIdx.x < 32) tmem_alloc(tptr, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + 216 * 1024 + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        // operands as in the frame-stacked layer 1: A = 4 stages x 16 KB at 128 KB.., B = 128 KB image at 0
        const uint64_t ad = desc_kmajor(sbase + 128 * 1024, CHUNK_A), bd = desc_kmajor(sbase, N * 16);
        constexpr uint32_t idesc = umma_idesc_f16(128, N);
        for (int rep = 0; rep < 3; ++rep) {
            long long t0 = clock64();
#pragma unroll
            for (int i = 0; i < NM; ++i)
                umma_bf16(tmem, desc_advance(ad, (i & 15) * 4096), desc_advance(bd, (i & 15) * 2 * N * 16), idesc, i > 0);
            umma_commit(bars);
            long long t1 = clock64();
            mbar_wait(bars, rep & 1);
            long long t2 = clock64();
            if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
        *stop = 1;
    } else if (warp == 1 && lane == 0 && (MODE & 1)) {
        uint32_t ph = 0;
        long long n = 0;
        while (!*stop) {                                    // 16 KB bulk copies into a scratch region, one after the other
            mbar_expect_tx(bars + 8, 16384);
            bulk_load(sbase + 196 * 1024, gsrc + (n & 63) * 16384, 16384, bars + 8);
            mbar_wait(bars + 8, ph); ph ^= 1; ++n;
        }
        out[2] = n;
    } else if (warp >= 2 && warp < 5 && (MODE & 2)) {
        uint32_t v[32];
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
        uint32_t acc = 0;
        while (!*stop) {
            for (int j = 0; j < 8; ++j) { tmem_ld32(tl + j * 32, v); tmem_wait_ld(); acc += v[0] ^ v[31]; }
        }
        if (acc == 0x12345) out[3] = acc;
    } else if (warp >= 5 && (MODE & 4)) {
        uint4 *dst = reinterpret_cast<uint4 *>(gdst) + (warp - 5) * 32 + lane;
        int k = 0;
        while (!*stop) { dst[(k & 1023) * 128] = make_uint4(k, k, k, k); ++k; }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int MODE>
void run(const uint8_t *gsrc, uint8_t *gdst) {
    long long *out; cudaMalloc(&out, 64); cudaMemset(out, 0, 64);
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 217 * 1024);
    probe<MODE><<<1, 288, 217 * 1024>>>(out, gsrc, gdst);
    long long h[4] = {0, 0, 0, 0};
    cudaError_t e = cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
    printf("beside: %s%s%s%s -> 64 x N=256 fp16: %.1f cyc/mma  (bulk copies done: %lld)  %s\n", MODE & 1 ? "bulk-copies " : "", MODE & 2 ? "tcgen05.ld " : "",
           MODE & 4 ? "global-stores " : "", MODE ? "" : "nothing", (double)h[1] / 64, h[2], e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out);
}

int main() {
    uint8_t *gsrc, *gdst;
    cudaMalloc(&gsrc, 64 * 16384); cudaMemset(gsrc, 1, 64 * 16384);
    cudaMalloc(&gdst, 1024 * 128 * 16 + 4096);
    run<0>(gsrc, gdst); run<1>(gsrc, gdst); run<2>(gsrc, gdst); run<4>(gsrc, gdst); run<3>(gsrc, gdst); run<7>(gsrc, gdst);
    return 0;
}

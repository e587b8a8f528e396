// umma_probe2.cu -- does the data or the operand format change the tcgen05.mma rate?  (development tool)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I skillshot_learning_b200/csrc -o gpurun_out/umma_probe2 tools/umma_probe2.cu
// 64 back-to-back MMAs 128 x N x 16 from one thread, one commit at the end: cycles per MMA for
// {bf16, fp16} x {all-zero operands, random operands}.
#include <cstdio>
#include "ss_tc_common.cuh"
using namespace sstc;

template <int N, bool F16, bool RANDOM>
__global__ void probe(long long *out) {
    constexpr int NM = 64;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + 200 * 1024, tptr = bars + 64;
    for (uint32_t o = threadIdx.x * 16; o < 200 * 1024; o += blockDim.x * 16) {
        uint32_t h = o * 2654435761u;
        // random finite half / bfloat values: clear the top exponent bit of each 16-bit lane
        const uint4 v = RANDOM ? make_uint4((h ^ 0x1234567u) & 0x3FFF3FFFu, (h * 7u) & 0x3FFF3FFFu, (h * 13u) & 0x3FFF3FFFu, (h * 29u) & 0x3FFF3FFFu)
                               : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(smem + o) = v;
    }
    if (threadIdx.x == 0) { mbar_init(bars, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(tptr, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + 200 * 1024 + 64);
    if (threadIdx.x == 0) {
        const uint64_t ad = desc_kmajor(sbase, CHUNK_A), bd = desc_kmajor(sbase + 64 * 1024, N * 16);
        constexpr uint32_t idesc = F16 ? umma_idesc_f16(128, N) : umma_idesc(128, N);
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
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, bool F16, bool RANDOM>
void run() {
    long long *out; cudaMalloc(&out, 16);
    cudaFuncSetAttribute(probe<N, F16, RANDOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    probe<N, F16, RANDOM><<<1, 128, 201 * 1024>>>(out);
    long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d %s %s operands: issue %6lld cyc, complete %6lld cyc  (%.1f cyc/mma; math floor %d)  %s\n", N, F16 ? "fp16" : "bf16",
           RANDOM ? "random" : "zero  ", h[0], h[1], (double)h[1] / 64, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out);
}

int main() {
    run<256, false, false>(); run<256, false, true>(); run<256, true, false>(); run<256, true, true>();
    run<128, false, false>(); run<128, false, true>(); run<128, true, true>();
    return 0;
}

// tmem_probe.cu -- tcgen05.ld throughput per SM with 1..8 warps reading (development tool).
#include <cstdio>
#include "ss_tc_common.cuh"
using namespace sstc;

template <int SHAPE>   // 32 or 16 columns per load
__global__ void probe(int nwarps, int reps, long long *out) {
    __shared__ uint32_t tptr;
    __shared__ long long t0s[16], t1s[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(smem_u32(&tptr), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tl = tptr + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps) {
        for (int r = 0; r < reps; ++r) {
            if (SHAPE == 32) {
                uint32_t va[32], vb[32];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    tmem_ld32(tl + j * 32, va);
                    tmem_ld32(tl + (j + 1) * 32, vb);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 32; ++e) acc ^= va[e] ^ vb[e];
                }
            } else {
                uint32_t va[16], vb[16];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    tmem_ld16(tl + j * 16, va);
                    tmem_ld16(tl + (j + 1) * 16, vb);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc ^= va[e] ^ vb[e];
                }
            }
        }
    }
    long long t1 = clock64();
    if (lane == 0) { t0s[warp] = t0; t1s[warp] = t1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long a = t0s[0], b = t1s[0];
        for (int w = 1; w < nwarps; ++w) { a = min(a, t0s[w]); b = max(b, t1s[w]); }
        out[0] = b - a;
    }
    if (acc == 0x12345678u) out[1] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tptr, 512); }
}

int main() {
    long long *out; cudaMalloc(&out, 16);
    for (int shape : {32, 16})
        for (int nw : {1, 2, 4, 8}) {
            const int reps = 64;
            if (shape == 32) probe<32><<<1, 256>>>(nw, reps, out); else probe<16><<<1, 256>>>(nw, reps, out);
            long long h = 0;
            cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)nw * reps * 256 * 32 * 4;
            printf("x%d loads, %d warps: %lld cycles for %.0f KB -> %.1f B/cycle per SM, %.1f cycles per 256-column row sweep per warp\n",
                   shape, nw, h, bytes / 1024, bytes / h, (double)h / reps);
        }
    return 0;
}

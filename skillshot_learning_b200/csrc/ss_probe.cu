// ss_probe.cu -- measured instruction-rate ceilings of the GPU the library runs on, for the roofline record of the
// fused step kernel.  That kernel plays K ticks per launch out of registers: its DRAM traffic is ~1/8 of the algorithmic
// bytes and it is bound by the warp schedulers' issue rate and the float64 pipe (ncu: profiles/), so bench.py reports
// it against these two measured ceilings next to the HBM figure.  Two register-only kernels, no memory traffic:
//   issue  8 independent integer chains per thread, alternating between the ALU pipe (LOP3) and the FMA pipe (IMAD), 16
//          warps per SM sub-partition.  Each of those pipes alone takes one warp-instruction every 2 cycles
//          (tools/op_probe.cu); fed in alternate cycles they reach the scheduler's one warp-instruction per cycle, so the
//          rate is SMs x 4 x clock (measured 0.94-0.99 of it)
//   fp64   8 independent DFMA chains per thread: the rate of the float64 pipe
// Unlike every other entry point this one synchronises (it times its own launches with CUDA events) and returns HOST
// numbers; it is a measurement aid, not part of the game / learner path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/skillshot_b200.h"

namespace {

constexpr int kChains = 8;
constexpr int kInner = 64;          // unrolled instructions per chain per loop trip

__global__ void __launch_bounds__(512) issue_stream_kernel(int *sink, int trips, int a, int b) {
    int x[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = (int)threadIdx.x + c;
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int i = 0; i < kInner; ++i) {
#pragma unroll
            for (int c = 0; c < kChains; c += 2) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));          // ALU pipe
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[c + 1]) : "r"(a), "r"(b));         // FMA pipe
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += x[c];
    if (s == 0x12345678 && trips < 0) sink[0] = s;        // never true: keeps the chains alive
}

__global__ void __launch_bounds__(512) dfma_stream_kernel(double *sink, int trips, double a, double b) {
    double x[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = (double)(threadIdx.x + c);
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int i = 0; i < kInner; ++i) {
#pragma unroll
            for (int c = 0; c < kChains; ++c) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += x[c];
    if (s == 12345.678) sink[0] = s;
}

template <class F>
double time_ms(F launch, cudaStream_t st) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();                                   // warm-up
    cudaStreamSynchronize(st);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0, st);
        launch();
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return (double)best;
}

}  // namespace

extern "C" int ss_probe_rates(double *out_host, void *scratch, void *stream) {
    if (!out_host || !scratch) return SS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return SS_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 4, block = 512;     // 2,048 threads = 64 warps per SM: 16 per sub-partition
    const double warps = (double)grid * (block / 32);
    const int trips_f = 64, trips_d = 16;
    const double ms_f = time_ms([&] { issue_stream_kernel<<<grid, block, 0, st>>>((int *)scratch, trips_f, 0x5bd1e995, 12345); }, st);
    const double ms_d = time_ms([&] { dfma_stream_kernel<<<grid, block, 0, st>>>((double *)scratch, trips_d, 0.999, 1e-3); }, st);
    if (cudaGetLastError() != cudaSuccess) return SS_ERR_CUDA;
    out_host[0] = warps * trips_f * kInner * kChains / (ms_f * 1e-3);     // warp-instructions per second, LOP3 / IMAD stream
    out_host[1] = warps * trips_d * kInner * kChains / (ms_d * 1e-3);     // warp-instructions per second, DFMA stream
    out_host[2] = (double)sms;
    out_host[3] = 0.0;
    return SS_OK;
}

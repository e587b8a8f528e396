// op_probe.cu -- throughput (many warps, independent chains) and latency (one warp, one dependent chain) of the
// instructions the step kernel is made of, on the GPU it runs on.  Register-only kernels, clock64-timed per SM
// sub-partition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/op_probe tools/op_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

enum Op { DFMA, DADD, DMUL, I2F64, F2I64, F2F64, FFMA, IMAD, LOP, SHFL, DSETP, MAGIC_CVT, N_OPS };
const char *kNames[N_OPS] = {"DFMA", "DADD", "DMUL", "I2F.F64", "F2I.F64", "F2F.F64.F32", "FFMA", "IMAD", "LOP3", "SHFL", "DSETP+SEL",
                             "hiloint2double+DADD"};

template <int OP, int CHAINS>
__global__ void probe(long long *cycles, double *sink, int trips, double a, double b) {
    double x[CHAINS];
    int xi[CHAINS];
    float xf[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = 1.0 + threadIdx.x + c; xi[c] = threadIdx.x + c; xf[c] = 1.0f + threadIdx.x + c; }
    __syncthreads();
    long long t0 = clock64();
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (OP == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(a), "d"(b));
                if (OP == DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(b));
                if (OP == DMUL) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(a));
                if (OP == I2F64) { asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(x[c]) : "r"(xi[c])); asm volatile("mov.b64 {%0, _}, %1;" : "=r"(xi[c]) : "d"(x[c])); }
                if (OP == F2I64) { asm volatile("cvt.rni.s32.f64 %0, %1;" : "=r"(xi[c]) : "d"(x[c])); asm volatile("mov.b64 %0, {%1, %2};" : "=d"(x[c]) : "r"(xi[c]), "r"(0x40590000)); }
                if (OP == F2F64) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(x[c]) : "f"(xf[c])); asm volatile("mov.b64 {_, %0}, %1;" : "=f"(xf[c]) : "d"(x[c])); }
                if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(xf[c]) : "f"((float)a), "f"((float)b));
                if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(xi[c]) : "r"(trips), "r"(i));
                if (OP == LOP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(xi[c]) : "r"(trips), "r"(i));
                if (OP == SHFL) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(xi[c]));
                if (OP == DSETP) { asm volatile("{ .reg .pred p; setp.lt.f64 p, %0, %1; selp.b32 %2, %2, %3, p; }" : "+d"(x[c]) , "+d"(b), "+r"(xi[c]) : "r"(i)); asm volatile("mov.b64 %0, {%1, %2};" : "=d"(x[c]) : "r"(xi[c]), "r"(0x40590000)); }
                if (OP == MAGIC_CVT) { asm volatile("mov.b64 %0, {%1, %2};" : "=d"(x[c]) : "r"(xi[c]), "r"(0x43300000)); asm volatile("sub.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(4503599627370496.0)); asm volatile("mov.b64 {%0, _}, %1;" : "=r"(xi[c]) : "d"(x[c])); }
            }
        }
    }
    long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c] + xi[c] + xf[c];
    if (s == 12345.678) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(long long *d_cyc, double *d_sink) {
    long long h[2];
    const int trips = 64;
    // latency: one warp, one chain.  throughput: 16 warps per sub-partition (2,048 threads on one SM, 2 CTAs of 1,024), 4 chains
    probe<OP, 1><<<1, 32>>>(d_cyc, d_sink, trips, 0.999, 1e-3);
    CHK(cudaDeviceSynchronize());
    CHK(cudaMemcpy(h, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    const double lat = (double)h[0] / (trips * 16);
    probe<OP, 4><<<1, 1024>>>(d_cyc, d_sink, trips, 0.999, 1e-3);
    CHK(cudaDeviceSynchronize());
    CHK(cudaMemcpy(h, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    // 1,024 threads = 8 warps per sub-partition, 4 chains each: warp-instructions per cycle per sub-partition
    const double per_smsp = (double)trips * 16 * 4 * 8 / (double)h[0];
    printf("%-22s latency %6.1f cycles per dependent op   throughput %.3f warp-instr/cycle/sub-partition (%.1f cycles per warp-instr)\n",
           kNames[OP], lat, per_smsp, 1.0 / per_smsp);
    return 0;
}

// ---- mixed streams: can two half-rate pipes be fed in alternate cycles? ----
enum Mix { ALU_FMA, ALU_FP64, FMA_FP64, ALU_FMA_FP64, ALU_FP32, N_MIX };
const char *kMixNames[N_MIX] = {"LOP3 + IMAD", "LOP3 + DFMA", "IMAD + DFMA", "LOP3 + IMAD + DFMA", "LOP3 + FFMA"};

template <int MIX>
__global__ void mix_probe(long long *cycles, double *sink, int trips, double a, double b) {
    double x[4]; int xi[4], xj[4]; float xf[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { x[c] = 1.0 + threadIdx.x + c; xi[c] = threadIdx.x + c; xj[c] = threadIdx.x * 3 + c; xf[c] = 1.f + c; }
    __syncthreads();
    long long t0 = clock64();
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (MIX == ALU_FMA || MIX == ALU_FP64 || MIX == ALU_FMA_FP64 || MIX == ALU_FP32)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(xi[c]) : "r"(trips), "r"(i));
                if (MIX == ALU_FMA || MIX == FMA_FP64 || MIX == ALU_FMA_FP64)
                    asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(xj[c]) : "r"(trips), "r"(i));
                if (MIX == ALU_FP64 || MIX == FMA_FP64 || MIX == ALU_FMA_FP64)
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(a), "d"(b));
                if (MIX == ALU_FP32)
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(xf[c]) : "f"((float)a), "f"((float)b));
            }
        }
    }
    long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) s += x[c] + xi[c] + xj[c] + xf[c];
    if (s == 12345.678) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MIX>
int run_mix(long long *d_cyc, double *d_sink, int kinds) {
    long long h[1];
    const int trips = 64;
    mix_probe<MIX><<<1, 1024>>>(d_cyc, d_sink, trips, 0.999, 1e-3);
    CHK(cudaDeviceSynchronize());
    CHK(cudaMemcpy(h, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    const double per_smsp = (double)trips * 16 * 4 * kinds * 8 / (double)h[0];
    printf("%-22s mixed stream, 8 warps per sub-partition: %.3f warp-instr/cycle/sub-partition in total\n", kMixNames[MIX], per_smsp);
    return 0;
}

int main() {
    long long *d_cyc; double *d_sink;
    CHK(cudaMalloc(&d_cyc, 1024 * sizeof(long long)));
    CHK(cudaMalloc(&d_sink, 64));
    printf("(ops marked with a mov carry one extra MOV per op in the dependent chain)\n");
    if (run<DFMA>(d_cyc, d_sink)) return 1;
    if (run<DADD>(d_cyc, d_sink)) return 1;
    if (run<DMUL>(d_cyc, d_sink)) return 1;
    if (run<I2F64>(d_cyc, d_sink)) return 1;
    if (run<F2I64>(d_cyc, d_sink)) return 1;
    if (run<F2F64>(d_cyc, d_sink)) return 1;
    if (run<FFMA>(d_cyc, d_sink)) return 1;
    if (run<IMAD>(d_cyc, d_sink)) return 1;
    if (run<LOP>(d_cyc, d_sink)) return 1;
    if (run<SHFL>(d_cyc, d_sink)) return 1;
    if (run<DSETP>(d_cyc, d_sink)) return 1;
    if (run<MAGIC_CVT>(d_cyc, d_sink)) return 1;
    if (run_mix<ALU_FMA>(d_cyc, d_sink, 2)) return 1;
    if (run_mix<ALU_FP64>(d_cyc, d_sink, 2)) return 1;
    if (run_mix<FMA_FP64>(d_cyc, d_sink, 2)) return 1;
    if (run_mix<ALU_FMA_FP64>(d_cyc, d_sink, 3)) return 1;
    if (run_mix<ALU_FP32>(d_cyc, d_sink, 2)) return 1;
    return 0;
}

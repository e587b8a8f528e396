// ss_mlp_tc.cu -- the actor forward pass on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in tensor memory), sm_100a only.
//
// model_act* of the reference (SkillshotLearner.py:215-281) evaluates
// 12 -> 256 relu -> 128 relu -> 2 tanh for ONE observation per Keras predict().
// The batched rollout evaluates it for 2 x envs observations per tick
// (BASELINE.json configs 3-4: 524,288 .. 2,097,152 rows), which is two GEMMs:
//
//   layer 1   D1[128 x 256] = A0[128 x 32]  . B1[32 x 256]     (K = 12 padded; see below)
//   layer 2   D2[128 x 128] = A1[128 x 256] . B2[256 x 128]    (91 % of the MACs)
//   layer 3   128 -> 2 and tanh on the CUDA cores in the epilogue (fp32)
//
// One persistent CTA per SM keeps the bf16 weights resident in shared memory in
// the UMMA canonical K-major (no-swizzle) layout and streams 128-row tiles
// through two SLOTS.  Each slot has its own 64 KB activation buffer, its own 256
// tensor-memory columns and its own four epilogue warps; one extra warp issues
// every tcgen05.mma.  While the tensor core works on one slot's GEMM the other
// slot's warps run their epilogue (bias, ReLU, bf16 pack -> next A operand; or
// layer 3 + tanh -> global), so the tensor pipe and the CUDA cores overlap.
// Synchronisation is mbarrier-only (tcgen05.commit arrives when the MMAs retire).
//
// Precision: weights bf16 (fp32 master copy stays in HBM), accumulation fp32.
// The observation is split into a bf16 high part and a bf16 low part (x = hi + lo
// to 2^-17 relative) that occupy K = 0..11 and 16..27 of layer 1 against the same
// weight rows, because positions are multiples of 1/250 and a single bf16 (8 bits)
// would merge neighbouring pixels.  The hidden activations are rounded to bf16
// when they become the A operand of layer 2.
//
// Parameter noise (SkillshotLearner.py:260-265): the CTA perturbs the weights
// while staging them, w + w * (sd * eps), eps from the same Philox stream as the
// float32 path (ss_rng.cuh), one draw per noise group; groups are multiples of
// the 128-row tile so a tile never mixes two draws.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/skillshot_b200.h"
#include "ss_rng.cuh"

namespace {

using namespace ss;

constexpr int TM = 128;                  // rows per tile = UMMA M = tensor-memory lanes
constexpr int DS = SS_DIM_STATE, DA = SS_DIM_ACTION, H1 = SS_HIDDEN1, H2 = SS_HIDDEN2;
constexpr int K1 = 32;                   // layer-1 K: [obs hi 12 | 0 x4 | obs lo 12 | 0 x4]
constexpr int NSLOT = 2;
constexpr int EPI_WARPS = 4 * NSLOT;     // warp w serves slot w / 4 and TMEM lanes 32 (w % 4) ..
constexpr int MMA_WARP = EPI_WARPS;
constexpr int NTHREADS = 32 * (EPI_WARPS + 1);

constexpr int A_W1 = 0, A_B1 = A_W1 + DS * H1, A_W2 = A_B1 + H1, A_B2 = A_W2 + H1 * H2, A_W3 = A_B2 + H2,
              A_B3 = A_W3 + H2 * DA, A_N = A_B3 + DA;
static_assert(A_N == SS_ACTOR_PARAMS, "actor parameter count");

// Canonical K-major operand tile, no swizzle: [K/8][rows][8] bf16.  A core matrix is
// 8 rows x 16 bytes stored contiguously (128 B); consecutive 8-row groups follow at
// 128 B (the descriptor's stride-dimension byte offset), consecutive K chunks at
// rows * 16 B (its leading-dimension byte offset).
constexpr uint32_t CHUNK_A = TM * 16;    // 2048: K-chunk stride of an activation tile
constexpr uint32_t CHUNK_B1 = H1 * 16;   // 4096
constexpr uint32_t CHUNK_B2 = H2 * 16;   // 2048
constexpr uint32_t SBO = 128;

// shared-memory map (bytes)
constexpr uint32_t SM_B1 = 0;                               // [K1/8][256][8] bf16
constexpr uint32_t SM_B2 = SM_B1 + K1 * H1 * 2;             // [256/8][128][8] bf16
constexpr uint32_t SM_A = SM_B2 + H1 * H2 * 2;              // NSLOT x [256/8][128][8] bf16 (layer-1 A aliases its head)
constexpr uint32_t SM_BIAS1 = SM_A + NSLOT * TM * H1 * 2;
constexpr uint32_t SM_BIAS2 = SM_BIAS1 + H1 * 4;
constexpr uint32_t SM_W3 = SM_BIAS2 + H2 * 4;               // [128][2] f32
constexpr uint32_t SM_B3 = SM_W3 + H2 * DA * 4;
constexpr uint32_t SM_BAR = SM_B3 + 16;                     // NSLOT x {in, d1, h1, d2} mbarriers
constexpr uint32_t SM_TMEM = SM_BAR + NSLOT * 4 * 8;
constexpr uint32_t SM_TOTAL = SM_TMEM + 16;
static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");

enum { BAR_IN = 0, BAR_D1 = 1, BAR_H1 = 2, BAR_D2 = 3 };

// ---- PTX wrappers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_test(bar, parity)) {}
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: start address, leading / stride byte offsets (all >> 4),
// descriptor version 1 (sm_100), layout type 0 = no swizzle
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, shape M x N (x 16)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier once every tcgen05 operation this thread issued so far has retired
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 consecutive fp32 columns of this thread's tensor-memory lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);     // .x = lo (low half-word)
    return *reinterpret_cast<const uint32_t *>(&v);
}

struct TcArgs {
    const float *theta, *obs;
    float *act;
    int64_t n, group;
    float param_sd, action_sd;
    uint64_t seed, counter;
};

// one parameter -> its place in shared memory (bf16 operand tiles, fp32 biases / layer 3)
__device__ __forceinline__ void place_param(uint8_t *smem, int p, float w) {
    if (p < A_B1) {                        // W1[k][n]: against the high and the low half of the observation
        const int k = p >> 8, n = p & 255;
        const __nv_bfloat16 b = __float2bfloat16_rn(w);
        const uint32_t off = SM_B1 + (uint32_t)((k >> 3) * H1 + n) * 16 + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(smem + off) = b;
        *reinterpret_cast<__nv_bfloat16 *>(smem + off + 2 * CHUNK_B1) = b;
    } else if (p < A_W2) {
        reinterpret_cast<float *>(smem + SM_BIAS1)[p - A_B1] = w;
    } else if (p < A_B2) {                 // W2[k][n]
        const int e = p - A_W2, k = e >> 7, n = e & 127;
        *reinterpret_cast<__nv_bfloat16 *>(smem + SM_B2 + (uint32_t)((k >> 3) * H2 + n) * 16 + (k & 7) * 2) =
            __float2bfloat16_rn(w);
    } else if (p < A_W3) {
        reinterpret_cast<float *>(smem + SM_BIAS2)[p - A_B2] = w;
    } else if (p < A_B3) {
        reinterpret_cast<float *>(smem + SM_W3)[p - A_W3] = w;
    } else {
        reinterpret_cast<float *>(smem + SM_B3)[p - A_B3] = w;
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) actor_fwd_tc_kernel(const TcArgs A) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto bar = [&](int slot, int which) -> uint32_t { return sbase + SM_BAR + (uint32_t)(slot * 4 + which) * 8; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLOT; ++s) {
            mbar_init(bar(s, BAR_IN), 128);
            mbar_init(bar(s, BAR_D1), 1);
            mbar_init(bar(s, BAR_H1), 128);
            mbar_init(bar(s, BAR_D2), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {                // one warp owns the tensor-memory allocation (all 512 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SM_TMEM), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // rows k = 12..15 and 28..31 of B1 stay zero for the whole kernel
    for (uint32_t o = threadIdx.x * 16; o < (uint32_t)K1 * H1 * 2; o += NTHREADS * 16)
        *reinterpret_cast<uint4 *>(smem + SM_B1 + o) = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<const volatile uint32_t *>(smem + SM_TMEM);

    const bool noisy = A.param_sd > 0.f;
    const int64_t group = noisy ? A.group : A.n;
    const int64_t upg = (group + TM - 1) / TM;                 // tiles per noise group
    const int64_t n_groups = (A.n + group - 1) / group;
    const int64_t units = n_groups * upg;
    const int64_t u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;

    uint32_t phase[NSLOT] = {0, 0};        // parity of the slot's barriers: each completes once per tile

    for (int64_t u = u0; u < u1;) {
        const int64_t g = u / upg;
        const int64_t seg_end = min(u1, (g + 1) * upg);
        const int64_t ntiles = seg_end - u;

        // ---- stage (and perturb) the weights of noise group g -----------------------------
        for (int q = threadIdx.x; q < (A_N + 3) / 4; q += NTHREADS) {
            float w[4], z[4] = {0.f, 0.f, 0.f, 0.f};
            if (4 * q + 3 < A_N) {
                const float4 v = *reinterpret_cast<const float4 *>(A.theta + 4 * q);
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            } else {
                for (int e = 0; e < 4; ++e) w[e] = (4 * q + e < A_N) ? A.theta[4 * q + e] : 0.f;
            }
            if (noisy) normal4(A.seed, kTagParamNoise, (uint32_t)q, (uint32_t)g, A.counter, z);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (4 * q + e < A_N) place_param(smem, 4 * q + e, noisy ? w[e] + w[e] * (A.param_sd * z[e]) : w[e]);
        }
        fence_proxy_async();               // generic-proxy writes -> visible to the tensor core's async proxy
        __syncthreads();

        if (warp < EPI_WARPS) {
            // ================= epilogue / producer warps of one slot =================
            const int slot = warp >> 2, qd = warp & 3;
            const int r = qd * 32 + lane;                                // row of the tile = tensor-memory lane
            uint8_t *abuf = smem + SM_A + (uint32_t)slot * (TM * H1 * 2);
            const uint32_t tlane = tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)slot * 256;
            const float *bias1 = reinterpret_cast<const float *>(smem + SM_BIAS1);
            const float *bias2 = reinterpret_cast<const float *>(smem + SM_BIAS2);
            const float *w3 = reinterpret_cast<const float *>(smem + SM_W3);
            const float *b3 = reinterpret_cast<const float *>(smem + SM_B3);
            const int64_t end = min(A.n, (g + 1) * group);
            uint32_t ph = phase[slot];
            for (int64_t k = slot; k < ntiles; k += NSLOT) {
                const int64_t base = g * group + (u + k - g * upg) * TM;
                const int64_t row = base + r;
                // ---- 1. observation -> A0 = [hi | lo] bf16, K = 32 ----
                float x[12];
                if (row < end) {
                    const float4 *src = reinterpret_cast<const float4 *>(A.obs + row * DS);
                    const float4 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2);
                    x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y;
                    x[6] = v1.z; x[7] = v1.w; x[8] = v2.x; x[9] = v2.y; x[10] = v2.z; x[11] = v2.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 12; ++e) x[e] = 0.f;
                }
                float lo[12];
#pragma unroll
                for (int e = 0; e < 12; ++e) {
                    const float hi = __bfloat162float(__float2bfloat16_rn(x[e]));
                    lo[e] = x[e] - hi;
                    x[e] = hi;
                }
                *reinterpret_cast<uint4 *>(abuf + 0 * CHUNK_A + r * 16) =
                    make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
                *reinterpret_cast<uint4 *>(abuf + 1 * CHUNK_A + r * 16) =
                    make_uint4(pack_bf16(x[8], x[9]), pack_bf16(x[10], x[11]), 0u, 0u);
                *reinterpret_cast<uint4 *>(abuf + 2 * CHUNK_A + r * 16) =
                    make_uint4(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]), pack_bf16(lo[4], lo[5]), pack_bf16(lo[6], lo[7]));
                *reinterpret_cast<uint4 *>(abuf + 3 * CHUNK_A + r * 16) =
                    make_uint4(pack_bf16(lo[8], lo[9]), pack_bf16(lo[10], lo[11]), 0u, 0u);
                fence_proxy_async();
                tc_fence_before();         // orders this thread's earlier tcgen05.ld (previous tile) before the next MMA
                mbar_arrive(bar(slot, BAR_IN));

                // ---- 2. D1 -> bias, ReLU, bf16 -> A1 ----
                mbar_wait(bar(slot, BAR_D1), ph);
                tc_fence_after();
#pragma unroll 1
                for (int j = 0; j < H1 / 32; ++j) {
                    uint32_t v[32];
                    tmem_ld32(tlane + j * 32, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int col = j * 32 + c * 8;
                        const float4 b0 = *reinterpret_cast<const float4 *>(bias1 + col);
                        const float4 b1 = *reinterpret_cast<const float4 *>(bias1 + col + 4);
                        const float h0 = fmaxf(__uint_as_float(v[c * 8 + 0]) + b0.x, 0.f);
                        const float h1 = fmaxf(__uint_as_float(v[c * 8 + 1]) + b0.y, 0.f);
                        const float h2 = fmaxf(__uint_as_float(v[c * 8 + 2]) + b0.z, 0.f);
                        const float h3 = fmaxf(__uint_as_float(v[c * 8 + 3]) + b0.w, 0.f);
                        const float h4 = fmaxf(__uint_as_float(v[c * 8 + 4]) + b1.x, 0.f);
                        const float h5 = fmaxf(__uint_as_float(v[c * 8 + 5]) + b1.y, 0.f);
                        const float h6 = fmaxf(__uint_as_float(v[c * 8 + 6]) + b1.z, 0.f);
                        const float h7 = fmaxf(__uint_as_float(v[c * 8 + 7]) + b1.w, 0.f);
                        *reinterpret_cast<uint4 *>(abuf + (uint32_t)(col >> 3) * CHUNK_A + r * 16) =
                            make_uint4(pack_bf16(h0, h1), pack_bf16(h2, h3), pack_bf16(h4, h5), pack_bf16(h6, h7));
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(bar(slot, BAR_H1));

                // ---- 3. D2 -> bias, ReLU, layer 3 (128 -> 2), tanh -> actions ----
                mbar_wait(bar(slot, BAR_D2), ph);
                tc_fence_after();
                float z0 = b3[0], z1 = b3[1];
#pragma unroll 1
                for (int j = 0; j < H2 / 32; ++j) {
                    uint32_t v[32];
                    tmem_ld32(tlane + j * 32, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const int col = j * 32 + c * 2;
                        const float2 b = *reinterpret_cast<const float2 *>(bias2 + col);
                        const float4 w = *reinterpret_cast<const float4 *>(w3 + col * 2);
                        const float ha = fmaxf(__uint_as_float(v[c * 2 + 0]) + b.x, 0.f);
                        const float hb = fmaxf(__uint_as_float(v[c * 2 + 1]) + b.y, 0.f);
                        z0 = fmaf(ha, w.x, z0); z1 = fmaf(ha, w.y, z1);
                        z0 = fmaf(hb, w.z, z0); z1 = fmaf(hb, w.w, z1);
                    }
                }
                if (row < end) {
                    float a0 = tanhf(z0), a1 = tanhf(z1);
                    if (A.action_sd > 0.f) {                             // SkillshotLearner.py:238
                        float zn[4];
                        normal4(A.seed, kTagActionNoise, (uint32_t)row, (uint32_t)(row >> 32), A.counter, zn);
                        a0 += A.action_sd * zn[0];
                        a1 += A.action_sd * zn[1];
                    }
                    reinterpret_cast<float2 *>(A.act)[row] = make_float2(a0, a1);
                }
                ph ^= 1;
            }
            tc_fence_before();
            phase[slot] = ph;
        } else {
            // ================= the MMA-issuing warp =================
            int64_t left[NSLOT];
            int stage[NSLOT];
            for (int s = 0; s < NSLOT; ++s) {
                left[s] = ntiles > s ? (ntiles - s + NSLOT - 1) / NSLOT : 0;
                stage[s] = 0;
            }
            constexpr uint32_t kIdesc1 = umma_idesc(TM, H1), kIdesc2 = umma_idesc(TM, H2);
            while (left[0] > 0 || left[1] > 0) {
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    if (left[s] <= 0) continue;
                    const uint32_t a_addr = sbase + SM_A + (uint32_t)s * (TM * H1 * 2);
                    const uint32_t d_addr = tmem + (uint32_t)s * 256;
                    if (stage[s] == 0) {
                        if (!mbar_test(bar(s, BAR_IN), phase[s])) continue;
                        tc_fence_after();
                        if (lane == 0) {
                            const uint64_t ad = umma_desc(a_addr, CHUNK_A, SBO), bd = umma_desc(sbase + SM_B1, CHUNK_B1, SBO);
#pragma unroll
                            for (int ks = 0; ks < K1 / 16; ++ks)
                                umma_bf16(d_addr, ad + (uint64_t)((2 * CHUNK_A * ks) >> 4),
                                          bd + (uint64_t)((2 * CHUNK_B1 * ks) >> 4), kIdesc1, ks > 0);
                            umma_commit(bar(s, BAR_D1));
                        }
                        __syncwarp();
                        stage[s] = 1;
                    } else {
                        if (!mbar_test(bar(s, BAR_H1), phase[s])) continue;
                        tc_fence_after();
                        if (lane == 0) {
                            const uint64_t ad = umma_desc(a_addr, CHUNK_A, SBO), bd = umma_desc(sbase + SM_B2, CHUNK_B2, SBO);
#pragma unroll
                            for (int ks = 0; ks < H1 / 16; ++ks)
                                umma_bf16(d_addr, ad + (uint64_t)((2 * CHUNK_A * ks) >> 4),
                                          bd + (uint64_t)((2 * CHUNK_B2 * ks) >> 4), kIdesc2, ks > 0);
                            umma_commit(bar(s, BAR_D2));
                        }
                        __syncwarp();
                        stage[s] = 0;
                        left[s] -= 1;
                        phase[s] ^= 1;
                    }
                }
            }
        }
        u = seg_end;
        __syncthreads();                   // every MMA of the segment has retired (the epilogue waited on D2)
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

}  // namespace

extern "C" int ss_actor_forward_tc(const float *actor_params, const float *obs, float *act_out, int64_t n,
                                   float param_noise_sd, int64_t noise_group, float action_noise_sd,
                                   uint64_t seed, uint64_t counter, void *stream) {
    if (!actor_params || !obs || !act_out || n <= 0 || param_noise_sd < 0.f || action_noise_sd < 0.f)
        return SS_ERR_INVALID_ARG;
    if (((uintptr_t)actor_params | (uintptr_t)obs) & 15 || ((uintptr_t)act_out & 7)) return SS_ERR_INVALID_ARG;
    const bool noisy = param_noise_sd > 0.f;
    if (noisy && (noise_group <= 0 || noise_group % TM != 0)) return SS_ERR_INVALID_ARG;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return SS_ERR_CUDA;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return SS_ERR_CUDA;
    if (cudaFuncSetAttribute(actor_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL) != cudaSuccess)
        return SS_ERR_CUDA;
    const int64_t group = noisy ? noise_group : n;
    const int64_t units = ((n + group - 1) / group) * ((group + TM - 1) / TM);
    const int grid = (int)(units < sms ? units : sms);
    TcArgs A{actor_params, obs, act_out, n, group, param_noise_sd, action_noise_sd, seed, counter};
    actor_fwd_tc_kernel<<<grid, NTHREADS, SM_TOTAL, (cudaStream_t)stream>>>(A);
    return cudaGetLastError() == cudaSuccess ? SS_OK : SS_ERR_CUDA;
}

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
// when they become the A operand of layer 2.  Both biases ride through the tensor
// core as extra K rows against a constant 1 in the A operand (bias as a bf16 high +
// low pair: fp32-accurate to 2^-17), so the epilogues are ReLU + pack only.
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
constexpr int K1 = 32;                   // layer-1 K: [obs hi 12 | 1 1 0 0 | obs lo 12 | 0 x4]; rows 12, 13 of B1 = b1 hi, lo
constexpr int K2 = H1 + 16;              // layer-2 K: [h1 256 | 1 1 0 x14];                rows 256, 257 of B2 = b2 hi, lo
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
constexpr uint32_t A_BYTES = (K2 / 8) * CHUNK_A;            // 69,632 per slot
constexpr uint32_t ONES = 0x3F803F80u;                      // bf16 {1, 1}

// shared-memory map (bytes)
constexpr uint32_t SM_B1 = 0;                               // [K1/8][256][8] bf16
constexpr uint32_t SM_B2 = SM_B1 + (K1 / 8) * CHUNK_B1;     // [K2/8][128][8] bf16
constexpr uint32_t SM_A = SM_B2 + (K2 / 8) * CHUNK_B2;      // NSLOT x [K2/8][128][8] bf16 (layer-1 A aliases its head)
constexpr uint32_t SM_W3 = SM_A + NSLOT * A_BYTES;          // [128][2] f32
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
// non-blocking probe (the MMA warp polls two slots; try_wait would park it on one of them)
__device__ __forceinline__ bool mbar_probe(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
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
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));        // first source -> upper half
    return d;
}
// ReLU fused into the conversion: max(x, 0) rounded to bf16, two at a time
__device__ __forceinline__ uint32_t pack_relu_bf16(uint32_t lo, uint32_t hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
    return d;
}
// 32 accumulator columns -> ReLU -> bf16 -> four 16-byte K chunks of this thread's A-operand row
__device__ __forceinline__ void relu_pack_store(const uint32_t (&v)[32], uint8_t *dst) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4 *>(dst + c * CHUNK_A) =
            make_uint4(pack_relu_bf16(v[c * 8 + 0], v[c * 8 + 1]), pack_relu_bf16(v[c * 8 + 2], v[c * 8 + 3]),
                       pack_relu_bf16(v[c * 8 + 4], v[c * 8 + 5]), pack_relu_bf16(v[c * 8 + 6], v[c * 8 + 7]));
}
// 32 accumulator columns (hidden units col0 ..) -> ReLU -> the two 128 -> 2 dot products, split accumulators
__device__ __forceinline__ void relu_dot(const uint32_t (&v)[32], const float *w3, float (&z0)[4], float (&z1)[4]) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float4 w = *reinterpret_cast<const float4 *>(w3 + c * 4);        // W3[col][0..1], W3[col+1][0..1]
        const float ha = fmaxf(__uint_as_float(v[c * 2 + 0]), 0.f), hb = fmaxf(__uint_as_float(v[c * 2 + 1]), 0.f);
        z0[(c & 1) * 2 + 0] = fmaf(ha, w.x, z0[(c & 1) * 2 + 0]);
        z1[(c & 1) * 2 + 0] = fmaf(ha, w.y, z1[(c & 1) * 2 + 0]);
        z0[(c & 1) * 2 + 1] = fmaf(hb, w.z, z0[(c & 1) * 2 + 1]);
        z1[(c & 1) * 2 + 1] = fmaf(hb, w.w, z1[(c & 1) * 2 + 1]);
    }
}
__device__ __forceinline__ uint32_t hi_lo_bf16(float w) {                     // {hi, lo} with hi + lo = w to 2^-17
    const float hi = __bfloat162float(__float2bfloat16_rn(w));
    return pack_bf16(hi, w - hi);
}

struct TcArgs {
    const float *theta, *obs;
    float *act;
    int64_t n, group;
    float param_sd, action_sd;
    uint64_t seed, counter;
};

// ---- weight staging ------------------------------------------------------------
// fast N(0,1) quad for the staging loop (same Philox draw as normal4; intrinsic log / sincos:
// differs from the float32 path's draw by ~1e-6, far below the bf16 rounding that follows)
__device__ __forceinline__ void normal4_fast(uint64_t seed, uint32_t q, uint32_t g, uint64_t counter, float *z) {
    const U4 u = draw4(seed, kTagParamNoise, q, g, counter);
    const float r0 = sqrtf(-2.0f * __logf(unit_open(u.x))), r1 = sqrtf(-2.0f * __logf(unit_open(u.z)));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * unit_open(u.y), &s0, &c0);
    __sincosf(6.283185307179586f * unit_open(u.w), &s1, &c1);
    z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
}

struct Stager {
    const float *theta;
    uint8_t *smem;
    bool noisy;
    float sd;
    uint64_t seed, counter;
    uint32_t group;
    // four consecutive parameters starting at p (p % 4 == 0), perturbed if asked
    __device__ __forceinline__ void perturb(int p, float4 &v) const {
        if (!noisy) return;
        float z[4];
        normal4_fast(seed, (uint32_t)(p >> 2), group, counter, z);
        v.x += v.x * (sd * z[0]); v.y += v.y * (sd * z[1]); v.z += v.z * (sd * z[2]); v.w += v.w * (sd * z[3]);
    }
    __device__ __forceinline__ float4 load(int p) const { return __ldg(reinterpret_cast<const float4 *>(theta + p)); }
};

// Weights are [k][n] with n contiguous in HBM and [k/8][n][k%8] bf16 in shared memory.  A task takes
// one K chunk (8 rows) of four consecutive columns: 8 independent 16-byte loads (coalesced over the
// lanes), 8 Philox quads if the weights are perturbed, then one 16-byte store per column.
__device__ __forceinline__ void stage_weights(const Stager &S) {
    constexpr int T_W2 = (H1 / 8) * (H2 / 4), T_W1 = 2 * (H1 / 4), T_B2 = H2 / 4, T_W3 = H2 * DA / 4;
    for (int t = threadIdx.x; t < T_W2 + T_W1 + T_B2 + T_W3 + 1; t += NTHREADS) {
        if (t < T_W2) {
            const int kc = t / (H2 / 4), n = (t % (H2 / 4)) * 4;
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = S.load(A_W2 + (kc * 8 + i) * H2 + n);
#pragma unroll
            for (int i = 0; i < 8; ++i) S.perturb(A_W2 + (kc * 8 + i) * H2 + n, v[i]);
            uint8_t *dst = S.smem + SM_B2 + (uint32_t)(kc * H2 + n) * 16;
            *reinterpret_cast<uint4 *>(dst + 0) = make_uint4(pack_bf16(v[0].x, v[1].x), pack_bf16(v[2].x, v[3].x), pack_bf16(v[4].x, v[5].x), pack_bf16(v[6].x, v[7].x));
            *reinterpret_cast<uint4 *>(dst + 16) = make_uint4(pack_bf16(v[0].y, v[1].y), pack_bf16(v[2].y, v[3].y), pack_bf16(v[4].y, v[5].y), pack_bf16(v[6].y, v[7].y));
            *reinterpret_cast<uint4 *>(dst + 32) = make_uint4(pack_bf16(v[0].z, v[1].z), pack_bf16(v[2].z, v[3].z), pack_bf16(v[4].z, v[5].z), pack_bf16(v[6].z, v[7].z));
            *reinterpret_cast<uint4 *>(dst + 48) = make_uint4(pack_bf16(v[0].w, v[1].w), pack_bf16(v[2].w, v[3].w), pack_bf16(v[4].w, v[5].w), pack_bf16(v[6].w, v[7].w));
        } else if (t < T_W2 + T_W1) {
            // W1 rows 0..7 (chunk 0) or rows 8..11 + the bias pair b1 hi, lo at K = 12, 13 (chunk 1);
            // the same rows serve the low half of the observation two chunks further on
            const int tt = t - T_W2, kc = tt / (H1 / 4), n = (tt % (H1 / 4)) * 4;
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = kc * 8 + i;
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < DS) { v[i] = S.load(A_W1 + k * H1 + n); S.perturb(A_W1 + k * H1 + n, v[i]); }
            }
            uint32_t e45[4] = {0u, 0u, 0u, 0u};
            if (kc == 1) {
                float4 b = S.load(A_B1 + n);
                S.perturb(A_B1 + n, b);
                e45[0] = hi_lo_bf16(b.x); e45[1] = hi_lo_bf16(b.y); e45[2] = hi_lo_bf16(b.z); e45[3] = hi_lo_bf16(b.w);
            }
            uint8_t *dst = S.smem + SM_B1 + (uint32_t)(kc * H1 + n) * 16;
            const float c[4][8] = {{v[0].x, v[1].x, v[2].x, v[3].x, v[4].x, v[5].x, v[6].x, v[7].x},
                                   {v[0].y, v[1].y, v[2].y, v[3].y, v[4].y, v[5].y, v[6].y, v[7].y},
                                   {v[0].z, v[1].z, v[2].z, v[3].z, v[4].z, v[5].z, v[6].z, v[7].z},
                                   {v[0].w, v[1].w, v[2].w, v[3].w, v[4].w, v[5].w, v[6].w, v[7].w}};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t w01 = pack_bf16(c[j][0], c[j][1]), w23 = pack_bf16(c[j][2], c[j][3]);
                const uint32_t w45 = pack_bf16(c[j][4], c[j][5]), w67 = pack_bf16(c[j][6], c[j][7]);
                *reinterpret_cast<uint4 *>(dst + j * 16) = make_uint4(w01, w23, kc == 1 ? e45[j] : w45, w67);
                *reinterpret_cast<uint4 *>(dst + j * 16 + 2 * CHUNK_B1) = make_uint4(w01, w23, kc == 1 ? 0u : w45, w67);
            }
        } else if (t < T_W2 + T_W1 + T_B2) {        // b2 hi, lo -> rows 256, 257 of B2
            const int n = (t - T_W2 - T_W1) * 4;
            float4 b = S.load(A_B2 + n);
            S.perturb(A_B2 + n, b);
            uint8_t *dst = S.smem + SM_B2 + (uint32_t)((H1 / 8) * H2 + n) * 16;
            *reinterpret_cast<uint4 *>(dst + 0) = make_uint4(hi_lo_bf16(b.x), 0u, 0u, 0u);
            *reinterpret_cast<uint4 *>(dst + 16) = make_uint4(hi_lo_bf16(b.y), 0u, 0u, 0u);
            *reinterpret_cast<uint4 *>(dst + 32) = make_uint4(hi_lo_bf16(b.z), 0u, 0u, 0u);
            *reinterpret_cast<uint4 *>(dst + 48) = make_uint4(hi_lo_bf16(b.w), 0u, 0u, 0u);
        } else if (t < T_W2 + T_W1 + T_B2 + T_W3) {  // W3 [128][2] fp32
            const int e = (t - T_W2 - T_W1 - T_B2) * 4;
            float4 w = S.load(A_W3 + e);
            S.perturb(A_W3 + e, w);
            *reinterpret_cast<float4 *>(S.smem + SM_W3 + e * 4) = w;
        } else {                                     // b3: the last, partial quad
            float4 w = make_float4(S.theta[A_B3], S.theta[A_B3 + 1], 0.f, 0.f);
            S.perturb(A_B3, w);
            *reinterpret_cast<float2 *>(S.smem + SM_B3) = make_float2(w.x, w.y);
        }
    }
}

// observation row -> registers (zeros beyond the end of the noise group / batch)
__device__ __forceinline__ void load_obs(const float *obs, int64_t row, int64_t end, float4 (&x)[3]) {
    if (row < end) {
        const float4 *src = reinterpret_cast<const float4 *>(obs + row * DS);
        x[0] = __ldg(src); x[1] = __ldg(src + 1); x[2] = __ldg(src + 2);
    } else {
        x[0] = x[1] = x[2] = make_float4(0.f, 0.f, 0.f, 0.f);
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
    // constant parts of the operand tiles: the padding rows of B1 / B2 stay zero, the bias rows of
    // the activation tiles (K = 256, 257) stay one, for the whole kernel
    for (uint32_t o = threadIdx.x * 16; o < SM_A; o += NTHREADS * 16)
        *reinterpret_cast<uint4 *>(smem + o) = make_uint4(0, 0, 0, 0);
    if (warp < EPI_WARPS) {
        uint8_t *abuf = smem + SM_A + (uint32_t)(warp >> 2) * A_BYTES + (uint32_t)((warp & 3) * 32 + lane) * 16;
        *reinterpret_cast<uint4 *>(abuf + (H1 / 8) * CHUNK_A) = make_uint4(ONES, 0, 0, 0);
        *reinterpret_cast<uint4 *>(abuf + (H1 / 8 + 1) * CHUNK_A) = make_uint4(0, 0, 0, 0);
    }
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
        const int64_t end = min(A.n, (g + 1) * group);

        // the epilogue warps start fetching their first observation rows under the weight staging
        float4 xin[3];
        if (warp < EPI_WARPS)
            load_obs(A.obs, g * group + (u + (warp >> 2) - g * upg) * TM + (warp & 3) * 32 + lane,
                     (warp >> 2) < ntiles ? end : 0, xin);

        // ---- stage (and perturb) the weights of noise group g ----
        stage_weights(Stager{A.theta, smem, noisy, A.param_sd, A.seed, A.counter, (uint32_t)g});
        fence_proxy_async();               // generic-proxy writes -> visible to the tensor core's async proxy
        __syncthreads();

        if (warp < EPI_WARPS) {
            // ================= epilogue / producer warps of one slot =================
            const int slot = warp >> 2, qd = warp & 3;
            const int r = qd * 32 + lane;                                // row of the tile = tensor-memory lane
            uint8_t *arow = smem + SM_A + (uint32_t)slot * A_BYTES + (uint32_t)r * 16;
            const uint32_t tlane = tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)slot * 256;
            const float *w3 = reinterpret_cast<const float *>(smem + SM_W3);
            const float *b3 = reinterpret_cast<const float *>(smem + SM_B3);
            uint32_t ph = phase[slot];
            for (int64_t k = slot; k < ntiles; k += NSLOT) {
                const int64_t row = g * group + (u + k - g * upg) * TM + r;
                // ---- 1. observation -> A0 = [hi | 1 1 | lo] bf16, K = 32 ----
                {
                    const float x[12] = {xin[0].x, xin[0].y, xin[0].z, xin[0].w, xin[1].x, xin[1].y,
                                         xin[1].z, xin[1].w, xin[2].x, xin[2].y, xin[2].z, xin[2].w};
                    float hi[12], lo[12];
#pragma unroll
                    for (int e = 0; e < 12; ++e) {
                        hi[e] = __bfloat162float(__float2bfloat16_rn(x[e]));
                        lo[e] = x[e] - hi[e];
                    }
                    *reinterpret_cast<uint4 *>(arow + 0 * CHUNK_A) = make_uint4(
                        pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], hi[3]), pack_bf16(hi[4], hi[5]), pack_bf16(hi[6], hi[7]));
                    *reinterpret_cast<uint4 *>(arow + 1 * CHUNK_A) =
                        make_uint4(pack_bf16(hi[8], hi[9]), pack_bf16(hi[10], hi[11]), ONES, 0u);
                    *reinterpret_cast<uint4 *>(arow + 2 * CHUNK_A) = make_uint4(
                        pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]), pack_bf16(lo[4], lo[5]), pack_bf16(lo[6], lo[7]));
                    *reinterpret_cast<uint4 *>(arow + 3 * CHUNK_A) =
                        make_uint4(pack_bf16(lo[8], lo[9]), pack_bf16(lo[10], lo[11]), 0u, 0u);
                }
                fence_proxy_async();
                tc_fence_before();         // orders this thread's earlier tcgen05.ld (previous tile) before the next MMA
                mbar_arrive(bar(slot, BAR_IN));
                // next tile's observation row: in flight during this tile's two GEMMs
                load_obs(A.obs, row + NSLOT * TM, (k + NSLOT < ntiles) ? end : 0, xin);

                // ---- 2. D1 (bias included) -> ReLU, bf16 -> A1; tensor-memory loads double-buffered ----
                mbar_wait(bar(slot, BAR_D1), ph);
                tc_fence_after();
                {
                    uint32_t va[32], vb[32];
                    tmem_ld32(tlane, va);
#pragma unroll
                    for (int j = 0; j < H1 / 32; j += 2) {
                        tmem_wait_ld();
                        tmem_ld32(tlane + (j + 1) * 32, vb);
                        relu_pack_store(va, arow + (uint32_t)(j * 4) * CHUNK_A);
                        tmem_wait_ld();
                        if (j + 2 < H1 / 32) tmem_ld32(tlane + (j + 2) * 32, va);
                        relu_pack_store(vb, arow + (uint32_t)((j + 1) * 4) * CHUNK_A);
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(bar(slot, BAR_H1));

                // ---- 3. D2 (bias included) -> ReLU, layer 3 (128 -> 2), tanh -> actions ----
                mbar_wait(bar(slot, BAR_D2), ph);
                tc_fence_after();
                float z0[4] = {b3[0], 0.f, 0.f, 0.f}, z1[4] = {b3[1], 0.f, 0.f, 0.f};
                {
                    uint32_t va[32], vb[32];
                    tmem_ld32(tlane, va);
#pragma unroll
                    for (int j = 0; j < H2 / 32; j += 2) {
                        tmem_wait_ld();
                        tmem_ld32(tlane + (j + 1) * 32, vb);
                        relu_dot(va, w3 + j * 64, z0, z1);
                        tmem_wait_ld();
                        if (j + 2 < H2 / 32) tmem_ld32(tlane + (j + 2) * 32, va);
                        relu_dot(vb, w3 + (j + 1) * 64, z0, z1);
                    }
                }
                if (row < end) {
                    float a0 = tanhf((z0[0] + z0[1]) + (z0[2] + z0[3])), a1 = tanhf((z1[0] + z1[1]) + (z1[2] + z1[3]));
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
                    const uint32_t a_addr = sbase + SM_A + (uint32_t)s * A_BYTES;
                    const uint32_t d_addr = tmem + (uint32_t)s * 256;
                    if (stage[s] == 0) {
                        if (!mbar_probe(bar(s, BAR_IN), phase[s])) continue;
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
                        if (!mbar_probe(bar(s, BAR_H1), phase[s])) continue;
                        tc_fence_after();
                        if (lane == 0) {
                            const uint64_t ad = umma_desc(a_addr, CHUNK_A, SBO), bd = umma_desc(sbase + SM_B2, CHUNK_B2, SBO);
#pragma unroll
                            for (int ks = 0; ks < K2 / 16; ++ks)
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

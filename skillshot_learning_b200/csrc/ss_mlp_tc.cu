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
#include "ss_tc_common.cuh"

namespace {

using namespace sstc;

constexpr int NSLOT = 2;
constexpr int EPI_WARPS = 4 * NSLOT;     // warp w serves slot w / 4 and TMEM lanes 32 (w % 4) ..
constexpr int MMA_WARP = EPI_WARPS;
constexpr int NTHREADS = 32 * (EPI_WARPS + 1);

// shared-memory map (bytes)
constexpr uint32_t SM_B1 = 0;                               // [K1/8][256][8] bf16
constexpr uint32_t SM_B2 = SM_B1 + B1_BYTES;                // [K2/8][128][8] bf16
constexpr uint32_t SM_A = SM_B2 + B2_BYTES;                 // NSLOT x [K2/8][128][8] bf16 (layer-1 A aliases its head)
constexpr uint32_t SM_W3 = SM_A + NSLOT * X2_BYTES;         // [128] float4 (layer 3, see Stager::w3x)
constexpr uint32_t SM_B3 = SM_W3 + H2 * 16;
constexpr uint32_t SM_BAR = SM_B3 + 16;                     // NSLOT x {in, d1, h1, d2} mbarriers
constexpr uint32_t SM_TMEM = SM_BAR + NSLOT * 4 * 8;
constexpr uint32_t SM_TOTAL = SM_TMEM + 16;
static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");
static_assert(NSLOT == 2, "the MMA warp services exactly two slots");

enum { BAR_IN = 0, BAR_D1 = 1, BAR_H1 = 2, BAR_D2 = 3 };

struct FwdArgs {
    const float *params, *obs;
    int64_t n, group;
    // actor: actions out, exploration noise
    float *act_out;
    float param_sd, action_sd;
    uint64_t seed, counter;
    // critic: actions in; any of q_out / up_out / y_out
    const float *act_in;
    float *q_out;            // [n]     Q(s, a)
    float *up_out;           // [n][2]  -dQ/da, the upstream gradient of the actor's policy-gradient step
    float *y_out;            // [n]     reward + gamma * (1 - done) * Q   (TD target when params are the target critic)
    const float *reward;
    const uint8_t *done;
    float gamma;
};

// 32 accumulator columns (hidden-2 units) -> ReLU -> layer 3, four split accumulators per output.
// actor: the two 128 -> 2 dot products (W3 as stored, two units per 16-byte load).
// critic: q, and dq/da_m = sum_k [z_k > 0] W3[k] W2[256+m][k].
template <int NET>
__device__ __forceinline__ void layer3(const uint32_t (&v)[32], const float4 *w3x, float (&acc)[3][4]) {
    if (NET == NET_ACTOR) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float4 w = w3x[c];                                           // W3[2c][0..1], W3[2c+1][0..1]
            const float ha = fmaxf(__uint_as_float(v[c * 2 + 0]), 0.f), hb = fmaxf(__uint_as_float(v[c * 2 + 1]), 0.f);
            acc[0][(c & 1) * 2 + 0] = fmaf(ha, w.x, acc[0][(c & 1) * 2 + 0]);
            acc[1][(c & 1) * 2 + 0] = fmaf(ha, w.y, acc[1][(c & 1) * 2 + 0]);
            acc[0][(c & 1) * 2 + 1] = fmaf(hb, w.z, acc[0][(c & 1) * 2 + 1]);
            acc[1][(c & 1) * 2 + 1] = fmaf(hb, w.w, acc[1][(c & 1) * 2 + 1]);
        }
    } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const float4 w = w3x[c];
            const float z = __uint_as_float(v[c]);
            const float gate = z > 0.f ? 1.f : 0.f;
            acc[0][c & 3] = fmaf(fmaxf(z, 0.f), w.x, acc[0][c & 3]);
            acc[1][c & 3] = fmaf(gate, w.y, acc[1][c & 3]);
            acc[2][c & 3] = fmaf(gate, w.z, acc[2][c & 3]);
        }
    }
}

template <int NET>
__global__ void __launch_bounds__(NTHREADS, 1) mlp_fwd_tc_kernel(const FwdArgs A) {
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
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(sbase + SM_TMEM, 512);     // one warp owns the allocation (all 512 columns)
    // constant parts of the operand tiles: the padding rows of B1 / B2 stay zero, the bias columns of
    // the activation tiles (K = 256.. ) stay one, for the whole kernel
    for (uint32_t o = threadIdx.x * 16; o < SM_A; o += NTHREADS * 16)
        *reinterpret_cast<uint4 *>(smem + o) = make_uint4(0, 0, 0, 0);
    if (warp < EPI_WARPS) {
        uint8_t *arow = smem + SM_A + (uint32_t)(warp >> 2) * X2_BYTES + (uint32_t)((warp & 3) * 32 + lane) * 16;
        *reinterpret_cast<uint4 *>(arow + (H1 / 8) * CHUNK_A) = tail_chunk_actor();
        *reinterpret_cast<uint4 *>(arow + (H1 / 8 + 1) * CHUNK_A) = make_uint4(0, 0, 0, 0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<const volatile uint32_t *>(smem + SM_TMEM);

    const bool noisy = NET == NET_ACTOR && A.param_sd > 0.f;
    const int64_t group = noisy ? A.group : A.n;
    const int64_t upg = (group + TM - 1) / TM;                 // tiles per noise group
    const int64_t n_groups = (A.n + group - 1) / group;
    const int64_t units = n_groups * upg;
    const int64_t u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;

    // parity of a slot's barriers (each completes once per tile): the epilogue warps track their own
    // slot, the MMA warp both
    uint32_t ph = 0, mma_phase0 = 0, mma_phase1 = 0;

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
        stage_weights<NET, NTHREADS>(Stager{A.params, smem + SM_B1, smem + SM_B2, reinterpret_cast<float4 *>(smem + SM_W3),
                                            reinterpret_cast<float *>(smem + SM_B3), noisy, A.param_sd, A.seed,
                                            A.counter, (uint32_t)g});
        fence_proxy_async();               // generic-proxy writes -> visible to the tensor core's async proxy
        __syncthreads();

        if (warp < EPI_WARPS) {
            // ================= epilogue / producer warps of one slot =================
            const int slot = warp >> 2, qd = warp & 3;
            const int r = qd * 32 + lane;                                // row of the tile = tensor-memory lane
            uint8_t *arow = smem + SM_A + (uint32_t)slot * X2_BYTES + (uint32_t)r * 16;
            const uint32_t tlane = tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)slot * 256;
            const float4 *w3x = reinterpret_cast<const float4 *>(smem + SM_W3);
            const float *b3 = reinterpret_cast<const float *>(smem + SM_B3);
            for (int64_t k = slot; k < ntiles; k += NSLOT) {
                const int64_t row = g * group + (u + k - g * upg) * TM + r;
                // ---- 1. observation -> layer-1 A operand (K = 32); critic: the action into the layer-2 tail chunk ----
                store_obs_row(xin, arow);
                if (NET == NET_CRITIC) {
                    float2 a = make_float2(0.f, 0.f);
                    if (row < end) a = __ldg(reinterpret_cast<const float2 *>(A.act_in) + row);
                    *reinterpret_cast<uint4 *>(arow + (H1 / 8) * CHUNK_A) = tail_chunk_critic(a.x, a.y);
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

                // ---- 3. D2 (bias included) -> ReLU, layer 3, output ----
                mbar_wait(bar(slot, BAR_D2), ph);
                tc_fence_after();
                float acc[3][4] = {{b3[0], 0.f, 0.f, 0.f}, {NET == NET_ACTOR ? b3[1] : 0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                {
                    constexpr int kW3Step = NET == NET_ACTOR ? 16 : 32;        // float4 entries per 32 units
                    uint32_t va[32], vb[32];
                    tmem_ld32(tlane, va);
#pragma unroll
                    for (int j = 0; j < H2 / 32; j += 2) {
                        tmem_wait_ld();
                        tmem_ld32(tlane + (j + 1) * 32, vb);
                        layer3<NET>(va, w3x + j * kW3Step, acc);
                        tmem_wait_ld();
                        if (j + 2 < H2 / 32) tmem_ld32(tlane + (j + 2) * 32, va);
                        layer3<NET>(vb, w3x + (j + 1) * kW3Step, acc);
                    }
                }
                const float out0 = (acc[0][0] + acc[0][1]) + (acc[0][2] + acc[0][3]);
                const float out1 = (acc[1][0] + acc[1][1]) + (acc[1][2] + acc[1][3]);
                const float out2 = (acc[2][0] + acc[2][1]) + (acc[2][2] + acc[2][3]);
                if (row < end) {
                    if (NET == NET_ACTOR) {
                        float a0 = tanhf(out0), a1 = tanhf(out1);
                        if (A.action_sd > 0.f) {                         // SkillshotLearner.py:238
                            float zn[4];
                            normal4(A.seed, kTagActionNoise, (uint32_t)row, (uint32_t)(row >> 32), A.counter, zn);
                            a0 += A.action_sd * zn[0];
                            a1 += A.action_sd * zn[1];
                        }
                        reinterpret_cast<float2 *>(A.act_out)[row] = make_float2(a0, a1);
                    } else {
                        const float q = out0;
                        if (A.q_out) A.q_out[row] = q;
                        if (A.up_out)                                    // output_gradients = -dq/da, SkillshotLearner.py:410
                            reinterpret_cast<float2 *>(A.up_out)[row] =
                                make_float2(-out1, -out2);
                        if (A.y_out)
                            A.y_out[row] = A.reward[row] + A.gamma * ((A.done && A.done[row]) ? 0.f : 1.f) * q;
                    }
                }
                ph ^= 1;
            }
            tc_fence_before();
        } else {
            // ================= the MMA-issuing warp =================
            // per slot: tiles left in this segment and which barrier the slot waits on next (0: inputs, 1: hidden 1)
            int64_t left0 = ntiles > 0 ? (ntiles + 1) / 2 : 0, left1 = ntiles > 1 ? ntiles / 2 : 0;
            int stage0 = 0, stage1 = 0;
            constexpr uint32_t kIdesc1 = umma_idesc(TM, H1), kIdesc2 = umma_idesc(TM, H2);
            auto service = [&](const int s, int64_t &left, int &stage, uint32_t &phase) {
                if (left <= 0) return;
                const uint32_t a_addr = sbase + SM_A + (uint32_t)s * X2_BYTES;
                const uint32_t d_addr = tmem + (uint32_t)s * 256;
                if (stage == 0) {
                    if (!mbar_test(bar(s, BAR_IN), phase)) return;
                    tc_fence_after();
                    if (lane == 0) {
                        const uint64_t ad = desc_kmajor(a_addr, CHUNK_A), bd = desc_kmajor(sbase + SM_B1, CHUNK_B1);
#pragma unroll
                        for (int ks = 0; ks < K1 / 16; ++ks)
                            umma_bf16(d_addr, desc_advance(ad, 2 * CHUNK_A * ks), desc_advance(bd, 2 * CHUNK_B1 * ks),
                                      kIdesc1, ks > 0);
                        umma_commit(bar(s, BAR_D1));
                    }
                    __syncwarp();
                    stage = 1;
                } else {
                    if (!mbar_test(bar(s, BAR_H1), phase)) return;
                    tc_fence_after();
                    if (lane == 0) {
                        const uint64_t ad = desc_kmajor(a_addr, CHUNK_A), bd = desc_kmajor(sbase + SM_B2, CHUNK_B2);
#pragma unroll
                        for (int ks = 0; ks < K2 / 16; ++ks)
                            umma_bf16(d_addr, desc_advance(ad, 2 * CHUNK_A * ks), desc_advance(bd, 2 * CHUNK_B2 * ks),
                                      kIdesc2, ks > 0);
                        umma_commit(bar(s, BAR_D2));
                    }
                    __syncwarp();
                    stage = 0;
                    left -= 1;
                    phase ^= 1;
                }
            };
            while (left0 > 0 || left1 > 0) {
                service(0, left0, stage0, mma_phase0);
                service(1, left1, stage1, mma_phase1);
            }
        }
        u = seg_end;
        __syncthreads();                   // every MMA of the segment has retired (the epilogue waited on D2)
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <int NET>
int launch_fwd(const FwdArgs &A, void *stream) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return SS_ERR_CUDA;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return SS_ERR_CUDA;
    if (cudaFuncSetAttribute(mlp_fwd_tc_kernel<NET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL) != cudaSuccess)
        return SS_ERR_CUDA;
    const int64_t units = ((A.n + A.group - 1) / A.group) * ((A.group + TM - 1) / TM);
    const int grid = (int)(units < sms ? units : sms);
    mlp_fwd_tc_kernel<NET><<<grid, NTHREADS, SM_TOTAL, (cudaStream_t)stream>>>(A);
    return cudaGetLastError() == cudaSuccess ? SS_OK : SS_ERR_CUDA;
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
    FwdArgs A{};
    A.params = actor_params; A.obs = obs; A.n = n; A.group = noisy ? noise_group : n;
    A.act_out = act_out; A.param_sd = param_noise_sd; A.action_sd = action_noise_sd; A.seed = seed; A.counter = counter;
    return launch_fwd<NET_ACTOR>(A, stream);
}

extern "C" int ss_critic_forward_tc(const float *critic_params, const float *obs, const float *act, int64_t n,
                                    float *q_out, float *neg_dq_da_out, const float *reward, const uint8_t *done,
                                    float gamma, float *y_out, void *stream) {
    if (!critic_params || !obs || !act || n <= 0 || (!q_out && !neg_dq_da_out && !y_out)) return SS_ERR_INVALID_ARG;
    if (y_out && !reward) return SS_ERR_INVALID_ARG;
    if (((uintptr_t)critic_params | (uintptr_t)obs) & 15 || (((uintptr_t)act | (uintptr_t)neg_dq_da_out) & 7))
        return SS_ERR_INVALID_ARG;
    FwdArgs A{};
    A.params = critic_params; A.obs = obs; A.n = n; A.group = n;
    A.act_in = act; A.q_out = q_out; A.up_out = neg_dq_da_out; A.y_out = y_out; A.reward = reward; A.done = done;
    A.gamma = gamma;
    return launch_fwd<NET_CRITIC>(A, stream);
}

// ss_peer.cu -- the multi-GPU exchange step of the update, fused with its neighbours over
// NVLink peer memory: [fixed-order reduction of the per-CTA gradient slices] -> all-reduce ->
// [tf.keras Adam + soft target update].
//
// The path has exactly one exchange per network step: a flat float32 gradient of 36,482 or
// 36,609 values (146 KB) must be summed over the ranks before Adam.  At that
// size a collective is pure latency, so instead of calling one:
//
//   push    the reduction kernel that produces rank r's gradient writes every value directly
//           into slot r of EVERY rank's inbox (peer-to-peer stores over NVLink / NVSwitch);
//           the last CTA to finish publishes "epoch e from rank r is complete" in every
//           inbox's flag word (release at system scope);
//   reduce  the Adam kernel on each rank waits for the world's flags of epoch e (acquire),
//           sums the inbox slots in rank order -- the same order on every rank, so all ranks
//           apply bit-identical updates and the weights never need a broadcast -- and applies
//           Adam and the soft target update in the same pass.
//
// Inboxes are double-buffered by epoch parity: a rank can only push epoch e+2 after it has
// reduced epoch e+1, which needed every peer's epoch-e+1 flag, which each peer publishes only
// after its own epoch-e reduction kernel has finished (stream order).  No other barrier exists.
//
// Memory: one allocation per rank {flags[2][world] | inbox[2][world][capacity] | per-CTA flags[2][world][ctas] of the
// one-kernel form}, created here
// with cudaMalloc and shared between the processes of one node through CUDA IPC handles
// (ss_peer_alloc / ss_peer_export / ss_peer_import), because the library's other entry points
// never allocate: these are the explicit exception.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/skillshot_b200.h"
#include "ss_launch.cuh"

namespace {

constexpr int kMaxWorld = SS_PEER_MAX_WORLD;

struct PeerTable {
    void *base[kMaxWorld];      // every rank's exchange allocation, as mapped in THIS process
};

__host__ __device__ inline int64_t flags_bytes(int world) { return (int64_t)(2 * world * sizeof(uint32_t) + 255) / 256 * 256; }
__device__ inline uint32_t *flag_ptr(void *base, int world, int parity, int src) {
    return reinterpret_cast<uint32_t *>(base) + parity * world + src;
}
__device__ inline float *inbox_ptr(void *base, int world, int64_t capacity, int parity, int src) {
    return reinterpret_cast<float *>(reinterpret_cast<char *>(base) + flags_bytes(world)) + ((int64_t)parity * world + src) * capacity;
}
// the one-kernel form's flags, one per (parity, source rank, 64-parameter CTA), behind the inboxes
__host__ __device__ inline int64_t cta_count(int64_t capacity) { return (capacity + 1 + 63) / 64; }
__device__ inline uint32_t *cta_flag_ptr(void *base, int world, int64_t capacity, int parity, int src, int cta) {
    uint32_t *f = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(base) + flags_bytes(world) +
                                               (int64_t)2 * world * capacity * (int64_t)sizeof(float));
    return f + ((int64_t)parity * world + src) * cta_count(capacity) + cta;
}

// grad[p] = sum of the CTA slices (fixed order), written into slot `rank` of every rank's inbox; aux[0] = extra slot.
// The last CTA to finish raises this rank's flag for `epoch` in every inbox.
__global__ void reduce_push_kernel(const float *work, int parts, int n_params, float *aux, PeerTable T, int world, int rank,
                                   int64_t capacity, uint32_t epoch, unsigned int *done_counter) {
    __shared__ float red[4][64];
    __shared__ bool last;
    const int p = blockIdx.x * 64 + threadIdx.x, q = threadIdx.y;
    const int parity = (int)(epoch & 1u);
    sslaunch::griddep_wait();            // ss_launch.cuh: placed beside the gradient kernel's last CTAs, held here until it is done
    sslaunch::griddep_launch();
    float s = 0.f;
    if (p <= n_params) {
#pragma unroll 10        // ten independent loads in flight per thread; the additions keep their order
        for (int c = q; c < parts; c += 4) s += __ldcg(work + (int64_t)c * (n_params + 1) + p);
    }
    red[q][threadIdx.x] = s;
    __syncthreads();
    if (q == 0 && p <= n_params) {
        const float t = ((red[0][threadIdx.x] + red[1][threadIdx.x]) + red[2][threadIdx.x]) + red[3][threadIdx.x];
        if (p < n_params) {
            for (int d = 0; d < world; ++d) inbox_ptr(T.base[d], world, capacity, parity, rank)[p] = t;
        } else if (aux) {
            aux[0] = t;
        }
    }
    // publish: every store of this CTA is visible system-wide before the counter moves
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) last = atomicAdd(done_counter, 1u) + 1u == gridDim.x;
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.y == 0 && threadIdx.x < world) {
            uint32_t *f = flag_ptr(T.base[threadIdx.x], world, parity, rank);
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
        }
        if (threadIdx.x == 0 && threadIdx.y == 0) *done_counter = 0;     // ready for the next launch on this stream
    }
}

// Wait for every rank's flag of `epoch`, sum the inbox slots in rank order, apply
// tf.keras Adam (SkillshotLearner.py:68, 118, 417) and the soft target update.
__global__ void peer_adam_kernel(void *base, int world, int64_t capacity, uint32_t epoch, float *params, float *m, float *v,
                                 float *target, float *grad_out, int64_t n, float lr_t, float beta1, float beta2, float eps,
                                 float tau, float grad_scale, uint32_t *status) {
    const int parity = (int)(epoch & 1u);
    __shared__ int timed_out;
    sslaunch::griddep_wait();            // ss_launch.cuh
    sslaunch::griddep_launch();
    // A timeout is sticky: once a rank's gradient has failed to arrive, no later step is applied either (the ranks would no
    // longer hold the same weights) until the host has seen the status word and decided what to do (PeerExchange.check_status).
    if (threadIdx.x == 0) timed_out = (status && (*(volatile uint32_t *)status & SS_STATUS_PEER_TIMEOUT)) ? 1 : 0;
    __syncthreads();
    if (timed_out) return;
    if (threadIdx.x < world) {
        const uint32_t *f = flag_ptr(base, world, parity, threadIdx.x);
        uint32_t seen = 0, spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
            if (seen == epoch) break;
            if (++spins > (1u << 26)) {                        // a peer never arrived: report instead of hanging
                if (status) atomicOr(status, SS_STATUS_PEER_TIMEOUT);
                *(volatile int *)&timed_out = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (timed_out) return;     // never sum a stale or partial inbox into Adam and the targets
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float g = 0.f;
    for (int r = 0; r < world; ++r) g += __ldcg(inbox_ptr(base, world, capacity, parity, r) + p);
    if (grad_out) grad_out[p] = g;
    const float gr = g * grad_scale;
    const float mm = beta1 * m[p] + (1.0f - beta1) * gr;
    const float vv = beta2 * v[p] + (1.0f - beta2) * gr * gr;
    m[p] = mm;
    v[p] = vv;
    const float w = params[p] - lr_t * mm / (sqrtf(vv) + eps);
    params[p] = w;
    if (target) target[p] = tau * w + (1.0f - tau) * target[p];
}

// push + wait + Adam as ONE kernel.  A CTA owns 64 parameters from the slice reduction to the parameter update: it sums the
// gradient slices (fixed order), stores the sums into slot `rank` of every rank's inbox, raises ITS flag of this epoch in
// every inbox (release at system scope, after a system-scope fence over the CTA's stores), waits for the same CTA's flag
// of every rank (acquire), sums the inbox slots in rank order -- the same order on every rank -- and applies Adam and the
// soft target update.  Against the two-kernel form: no kernel boundary between push and reduce, no grid-wide completion
// counter, and a CTA proceeds as soon as ITS 64 values have arrived from everybody instead of after everybody's last CTA.
// Every CTA pushes before it waits, so the ranks cannot wait for each other in a cycle.  The inbox parity argument of the
// header comment holds unchanged: epoch e + 2 is pushed only by the kernel after the one that waited for every rank's
// epoch e + 1 flags of EVERY CTA, and a rank publishes those only in the kernel after its own epoch-e reads.
// A timeout is sticky and applies nothing in the CTA that saw it (a rank that never launched times out every CTA alike).
__global__ void peer_reduce_adam_kernel(const float *work, int parts, int n_params, float *aux, PeerTable T, int world, int rank,
                                        int64_t capacity, uint32_t epoch, float *params, float *m, float *v, float *target,
                                        float *grad_out, float lr_t, float beta1, float beta2, float eps, float tau,
                                        float grad_scale, uint32_t *status) {
    __shared__ float red[4][64];
    __shared__ int timed_out;
    const int p = blockIdx.x * 64 + threadIdx.x, q = threadIdx.y;
    const int parity = (int)(epoch & 1u);
    sslaunch::griddep_wait();            // ss_launch.cuh
    sslaunch::griddep_launch();
    if (threadIdx.x == 0 && q == 0) timed_out = (status && (*(volatile uint32_t *)status & SS_STATUS_PEER_TIMEOUT)) ? 1 : 0;
    float s = 0.f;
    if (p <= n_params) {
#pragma unroll 10        // ten independent loads in flight per thread; the additions keep their order
        for (int c = q; c < parts; c += 4) s += __ldcg(work + (int64_t)c * (n_params + 1) + p);
    }
    red[q][threadIdx.x] = s;
    __syncthreads();
    if (timed_out) return;
    if (q == 0 && p <= n_params) {
        const float t = ((red[0][threadIdx.x] + red[1][threadIdx.x]) + red[2][threadIdx.x]) + red[3][threadIdx.x];
        if (p < n_params) {
            for (int d = 0; d < world; ++d) inbox_ptr(T.base[d], world, capacity, parity, rank)[p] = t;
        } else if (aux) {
            aux[0] = t;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (q == 0 && threadIdx.x < world) {
        uint32_t *mine = cta_flag_ptr(T.base[threadIdx.x], world, capacity, parity, rank, blockIdx.x);
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(mine), "r"(epoch) : "memory");
        const uint32_t *f = cta_flag_ptr(T.base[rank], world, capacity, parity, threadIdx.x, blockIdx.x);
        uint32_t seen = 0, spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
            if (seen == epoch) break;
            if (++spins > (1u << 26)) {                        // a peer never arrived: report instead of hanging
                if (status) atomicOr(status, SS_STATUS_PEER_TIMEOUT);
                *(volatile int *)&timed_out = 1;
                break;
            }
            __nanosleep(32);
        }
    }
    __syncthreads();
    if (timed_out || q != 0 || p >= n_params) return;
    float g = 0.f;
    for (int r = 0; r < world; ++r) g += __ldcg(inbox_ptr(T.base[rank], world, capacity, parity, r) + p);
    if (grad_out) grad_out[p] = g;
    const float gr = g * grad_scale;
    const float mm = beta1 * m[p] + (1.0f - beta1) * gr;
    const float vv = beta2 * v[p] + (1.0f - beta2) * gr * gr;
    m[p] = mm;
    v[p] = vv;
    const float w = params[p] - lr_t * mm / (sqrtf(vv) + eps);
    params[p] = w;
    if (target) target[p] = tau * w + (1.0f - tau) * target[p];
}

}  // namespace

extern "C" {

int64_t ss_peer_bytes(int world, int64_t capacity) {
    if (world < 1 || world > kMaxWorld || capacity < 1) return -1;
    return flags_bytes(world) + (int64_t)2 * world * capacity * (int64_t)sizeof(float) +
           (int64_t)2 * world * cta_count(capacity) * (int64_t)sizeof(uint32_t);
}

int ss_peer_alloc(int world, int64_t capacity, void **base_out) {
    const int64_t bytes = ss_peer_bytes(world, capacity);
    if (bytes < 0 || !base_out) return SS_ERR_INVALID_ARG;
    if (cudaMalloc(base_out, (size_t)bytes) != cudaSuccess) return SS_ERR_CUDA;
    if (cudaMemset(*base_out, 0, (size_t)bytes) != cudaSuccess) return SS_ERR_CUDA;
    return cudaDeviceSynchronize() == cudaSuccess ? SS_OK : SS_ERR_CUDA;
}

int ss_peer_free(void *base) { return cudaFree(base) == cudaSuccess ? SS_OK : SS_ERR_CUDA; }

int ss_peer_export(void *base, void *handle_out_host) {
    if (!base || !handle_out_host) return SS_ERR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == SS_PEER_HANDLE_BYTES, "IPC handle size");
    return cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle_out_host), base) == cudaSuccess ? SS_OK : SS_ERR_CUDA;
}

int ss_peer_import(const void *handle_host, void **base_out) {
    if (!handle_host || !base_out) return SS_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h = *reinterpret_cast<const cudaIpcMemHandle_t *>(handle_host);
    return cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? SS_OK : SS_ERR_CUDA;
}

int ss_peer_close(void *imported_base) { return cudaIpcCloseMemHandle(imported_base) == cudaSuccess ? SS_OK : SS_ERR_CUDA; }

int ss_peer_reduce_push(const void *workspace, int parts, int n_params, float *aux_out, void *const *peer_bases_host,
                        int world, int rank, int64_t capacity, uint32_t epoch, uint32_t *done_counter, void *stream) {
    if (!workspace || parts < 1 || n_params < 1 || !peer_bases_host || world < 1 || world > kMaxWorld || rank < 0 ||
        rank >= world || capacity < n_params || epoch == 0 || !done_counter)
        return SS_ERR_INVALID_ARG;
    PeerTable T{};
    for (int d = 0; d < world; ++d) {
        if (!peer_bases_host[d]) return SS_ERR_INVALID_ARG;
        T.base[d] = peer_bases_host[d];
    }
    if (sslaunch::launch(reduce_push_kernel, dim3((n_params + 1 + 63) / 64), dim3(64, 4), 0, (cudaStream_t)stream,
                         (const float *)workspace, parts, n_params, aux_out, T, world, rank, capacity, epoch,
                         (unsigned int *)done_counter) != cudaSuccess)
        return SS_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? SS_OK : SS_ERR_CUDA;
}

int ss_peer_reduce_adam_tf(const void *workspace, int parts, int n_params, float *aux_out, void *const *peer_bases_host,
                           int world, int rank, int64_t capacity, uint32_t epoch, float *params, float *m, float *v,
                           float *target_params, float *grad_out, int64_t step, float lr, float beta1, float beta2, float eps,
                           float tau, float grad_scale, uint32_t *status, void *stream) {
    if (!workspace || parts < 1 || n_params < 1 || !peer_bases_host || world < 1 || world > kMaxWorld || rank < 0 ||
        rank >= world || capacity < n_params || epoch == 0 || !params || !m || !v || step < 1)
        return SS_ERR_INVALID_ARG;
    PeerTable T{};
    for (int d = 0; d < world; ++d) {
        if (!peer_bases_host[d]) return SS_ERR_INVALID_ARG;
        T.base[d] = peer_bases_host[d];
    }
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
    if (sslaunch::launch(peer_reduce_adam_kernel, dim3((n_params + 1 + 63) / 64), dim3(64, 4), 0, (cudaStream_t)stream,
                         (const float *)workspace, parts, n_params, aux_out, T, world, rank, capacity, epoch, params, m, v,
                         target_params, grad_out, (float)lr_t, beta1, beta2, eps, tau, grad_scale, status) != cudaSuccess)
        return SS_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? SS_OK : SS_ERR_CUDA;
}

int ss_peer_adam_tf(void *own_base, int world, int64_t capacity, uint32_t epoch, float *params, float *m, float *v,
                    float *target_params, float *grad_out, int64_t n, int64_t step, float lr, float beta1, float beta2,
                    float eps, float tau, float grad_scale, uint32_t *status, void *stream) {
    if (!own_base || world < 1 || world > kMaxWorld || capacity < n || epoch == 0 || !params || !m || !v || n <= 0 || step < 1)
        return SS_ERR_INVALID_ARG;
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
    if (sslaunch::launch(peer_adam_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, own_base, world,
                         capacity, epoch, params, m, v, target_params, grad_out, n, (float)lr_t, beta1, beta2, eps, tau, grad_scale,
                         status) != cudaSuccess)
        return SS_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? SS_OK : SS_ERR_CUDA;
}

}  // extern "C"

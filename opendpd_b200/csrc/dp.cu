// dp.cu — data-parallel gradient exchange fused into the optimiser step, over NVLink peer memory (include/odpd.h).
//
// PUSH design, one NVLink traversal per step (round 1 pulled: publish + flag round trip + remote loads = >= 2 round trips):
// every rank owns a receive buffer  ll[2 parity][world sources][stride]  of 8-byte words {fp32 bits, step tag}.  One CTA per GPU
//   1. pushes its flat gradient (+ loss) as such words straight into slot [parity][own rank] of EVERY peer's buffer — plain
//      8-byte remote stores, fire and forget; data and flag travel in the same store (the "LL" idea of NCCL), so no
//      __threadfence_system, no separate flag and nothing is ever read across the link;
//   2. polls its OWN (local) buffer until every word of every source carries this step's tag;
//   3. sums the sources in rank order (bit-identical replicas), clips, applies AdamW.
// Slot reuse is safe with two parities: a rank can only publish step s+2 after it received every peer's step s+1, which a
// peer publishes after its step-s kernel (the consumer of parity s) completed in stream order.
#include <cstring>
#include "cells.h"

namespace odpd {

static constexpr int DP_MAX_WORLD = ODPD_DP_MAX_WORLD;    // 8 = one NVSwitch node; keeps the per-thread gather arrays in registers
struct DpPtrs { uint2 *buf[DP_MAX_WORLD]; };

__host__ __device__ inline int64_t dp_stride(int64_t n) { return (n + 1 + 3) & ~(int64_t)3; }

__device__ __forceinline__ void st_ll(uint2 *p, float v, unsigned tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const uint2 *p) {
    uint2 w;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

static constexpr unsigned long long DP_TIMEOUT_NS = 2000000000ull;   // a peer that has not published within 2 s is reported, not waited for

__global__ void __launch_bounds__(1024) dp_clip_adamw_kernel(float *__restrict__ p, DpPtrs bufs, int world, int rank, int64_t n,
                                                             const float *__restrict__ grad_local, const double *loss_local,
                                                             float *__restrict__ m, float *__restrict__ v, const float *__restrict__ lr_dev, float b1,
                                                             float b2, float eps, float wd, float max_norm, int64_t *step_dev, float *gnorm_out,
                                                             float *loss_out, int *status_dev) {
    __shared__ float red[32];
    __shared__ float s_coef;
    pdl_enter();
    const int64_t step = *step_dev + 1;
    const int64_t stride = dp_stride(n);
    const int par = (int)(step & 1);
    const unsigned tag = (unsigned)step;
    // ---- push: this rank's gradient (+ loss at element n) into slot [par][rank] of every peer
    float own[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int64_t i = threadIdx.x + (int64_t)e * blockDim.x;
        own[e] = 0.f;
        if (grad_local && i <= n) {      // grad_local == NULL: the backward's reduction already published (odpd_dp_publish_next_bwd)
            own[e] = (i < n) ? grad_local[i] : (loss_local ? (float)(*loss_local) : 0.f);
            const int64_t off = ((int64_t)par * world + rank) * stride + i;
#pragma unroll
            for (int r = 0; r < DP_MAX_WORLD; ++r)
                if (r < world && r != rank) st_ll(bufs.buf[r] + off, own[e], tag);
        }
    }
    // ---- gather from the LOCAL buffer + ordered sum (each thread owns up to 4 elements: n <= 4095 on this path)
    const uint2 *mine = bufs.buf[rank] + (int64_t)par * world * stride;
    const unsigned long long t0 = global_ns();
    int bad = 0;
    float g[4];
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int64_t i = threadIdx.x + (int64_t)e * blockDim.x;
        float acc = 0.f;
        if (i <= n) {
            float pv[DP_MAX_WORLD];
            unsigned pending = 0;
#pragma unroll
            for (int r = 0; r < DP_MAX_WORLD; ++r) {
                pv[r] = 0.f;
                if (r < world && (r != rank || !grad_local)) pending |= 1u << r;
            }
            while (pending && !bad) {
                uint2 w[DP_MAX_WORLD];
#pragma unroll
                for (int r = 0; r < DP_MAX_WORLD; ++r)         // all loads of a round in flight together
                    if ((pending >> r) & 1u) w[r] = ld_ll(mine + (int64_t)r * stride + i);
#pragma unroll
                for (int r = 0; r < DP_MAX_WORLD; ++r)
                    if (((pending >> r) & 1u) && w[r].y == tag) { pv[r] = __uint_as_float(w[r].x); pending &= ~(1u << r); }
                if (pending && global_ns() - t0 > DP_TIMEOUT_NS) bad = __ffs(pending);   // 1 + lowest rank still missing
            }
#pragma unroll
            for (int r = 0; r < DP_MAX_WORLD; ++r)
                if (r < world) acc += (r == rank && grad_local) ? own[e] : pv[r];
        }
        g[e] = acc;
        if (i < n) ss = fmaf(acc, acc, ss);
    }
    // a peer that never published: report it and leave the parameters and the step counter untouched (the host reads status_dev
    // on the chunk-controller cadence and raises; see NativeTrainStep._check_exchange)
    if (__syncthreads_or(bad)) {
        if (bad && status_dev) atomicMax(status_dev, bad);
        return;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (threadIdx.x + (int64_t)e * blockDim.x == n && loss_out) *loss_out = g[e];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) {
            const float norm = sqrtf(t);
            float coef = 1.f;
            if (max_norm > 0.f) { coef = max_norm / (norm + 1e-6f); coef = coef < 1.f ? coef : 1.f; }
            s_coef = coef;
            if (gnorm_out) *gnorm_out = norm;
        }
    }
    __syncthreads();
    const float coef = s_coef, lr = *lr_dev;
    const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
    const float step_size = lr / bc1, bc2s = sqrtf(bc2);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int64_t i = threadIdx.x + (int64_t)e * blockDim.x;
        if (i < n) {
            const float gi = g[e] * coef;
            float pi = p[i] * (1.f - lr * wd);
            const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
            const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
            pi -= step_size * (mi / (sqrtf(vi) / bc2s + eps));
            p[i] = pi; m[i] = mi; v[i] = vi;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *step_dev = step;
}

}  // namespace odpd

namespace odpd {
static thread_local DpPushArgs g_armed_push;
static thread_local bool g_armed = false;
bool dp_take_armed_push(DpPushArgs &out) {
    if (!g_armed) return false;
    out = g_armed_push;
    g_armed = false;
    return true;
}
}  // namespace odpd

using namespace odpd;

extern "C" {

int odpd_dp_publish_next_bwd(void *const *bufs, int world, int rank, int64_t n, const int64_t *step_dev, const double *loss_local) {
    ODPD_CHECK(bufs && step_dev, "odpd_dp_publish_next_bwd: NULL buffer");
    ODPD_CHECK(world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world, "odpd_dp_publish_next_bwd: bad world/rank (%d,%d)", world, rank);
    ODPD_CHECK(n >= 1 && n + 1 <= 4096, "odpd_dp_publish_next_bwd: n=%lld outside 1..4095", (long long)n);
    DpPushArgs p{};
    for (int r = 0; r < world; ++r) { ODPD_CHECK(bufs[r] != nullptr, "odpd_dp_publish_next_bwd: bufs[%d] is NULL", r); p.buf[r] = (uint2 *)bufs[r]; }
    p.world = world; p.rank = rank; p.stride = dp_stride(n); p.step_dev = step_dev; p.loss_local = loss_local;
    g_armed_push = p;
    g_armed = true;
    return 0;
}

int64_t odpd_dp_buffer_bytes(int64_t n_params) { return 2 * DP_MAX_WORLD * dp_stride(n_params) * (int64_t)sizeof(uint2); }

int odpd_dp_alloc(int64_t bytes, void **out_ptr) {
    ODPD_CHECK(out_ptr && bytes > 0, "odpd_dp_alloc: bad arguments");
    cudaError_t e = cudaMalloc(out_ptr, (size_t)bytes);
    ODPD_CHECK(e == cudaSuccess, "cudaMalloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
    e = cudaMemset(*out_ptr, 0, (size_t)bytes);
    ODPD_CHECK(e == cudaSuccess, "cudaMemset: %s", cudaGetErrorString(e));
    e = cudaDeviceSynchronize();
    ODPD_CHECK(e == cudaSuccess, "cudaDeviceSynchronize: %s", cudaGetErrorString(e));
    return 0;
}
int odpd_dp_free(void *ptr) {
    cudaError_t e = cudaFree(ptr);
    ODPD_CHECK(e == cudaSuccess, "cudaFree: %s", cudaGetErrorString(e));
    return 0;
}
int odpd_dp_ipc_handle(void *ptr, unsigned char handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    ODPD_CHECK(e == cudaSuccess, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    memcpy(handle_out, &h, 64);
    return 0;
}
int odpd_dp_ipc_open(const unsigned char handle[64], void **out_peer_ptr) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    cudaError_t e = cudaIpcOpenMemHandle(out_peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    ODPD_CHECK(e == cudaSuccess, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    return 0;
}
int odpd_dp_ipc_close(void *peer_ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
    ODPD_CHECK(e == cudaSuccess, "cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
    return 0;
}
int odpd_dp_clip_adamw(float *param, void *const *bufs, int world, int rank, int64_t n, const float *grad_local, const double *loss_local, float *exp_avg,
                       float *exp_avg_sq, const float *lr_dev, float beta1, float beta2, float eps, float weight_decay, float max_norm,
                       int64_t *step_dev, float *gnorm_out, float *loss_out, int *status_dev, void *stream) {
    ODPD_CHECK(param && bufs && exp_avg && exp_avg_sq && lr_dev && step_dev, "odpd_dp_clip_adamw: NULL buffer");
    ODPD_CHECK(world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world, "odpd_dp_clip_adamw: bad world/rank (%d,%d)", world, rank);
    ODPD_CHECK(n >= 1 && n + 1 <= 4096, "odpd_dp_clip_adamw: n=%lld outside 1..4095", (long long)n);
    DpPtrs p{};
    for (int r = 0; r < world; ++r) { ODPD_CHECK(bufs[r] != nullptr, "odpd_dp_clip_adamw: bufs[%d] is NULL", r); p.buf[r] = (uint2 *)bufs[r]; }
    launch_pdl(dp_clip_adamw_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, param, p, world, rank, n, grad_local, loss_local, exp_avg, exp_avg_sq,
               lr_dev, beta1, beta2, eps, weight_decay, max_norm, step_dev, gnorm_out, loss_out, status_dev);
    return check_launch("dp_clip_adamw_kernel");
}

}  // extern "C"

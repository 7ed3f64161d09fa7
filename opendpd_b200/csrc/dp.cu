// dp.cu — data-parallel gradient exchange fused into the optimiser step, over NVLink peer memory (include/odpd.h).
// One CTA per GPU: publish (system-scope release of a step counter), wait for every peer's counter (bounded spin), read all
// peers' flat gradients with cache-volatile loads straight across NVLink/NVSwitch, sum in rank order, clip, AdamW.
#include <cstring>
#include "cells.h"

namespace odpd {

static constexpr int DP_MAX_WORLD = 16;
struct DpPtrs { float *buf[DP_MAX_WORLD]; };

__host__ __device__ inline int64_t dp_stride(int64_t n) { return (n + 1 + 3) & ~(int64_t)3; }

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(1024) dp_clip_adamw_kernel(float *__restrict__ p, DpPtrs bufs, int world, int rank, int64_t n,
                                                             const float *__restrict__ grad_local, const double *loss_local,
                                                             float *__restrict__ m, float *__restrict__ v, const float *__restrict__ lr_dev, float b1,
                                                             float b2, float eps, float wd, float max_norm, int64_t *step_dev, float *gnorm_out,
                                                             float *loss_out, int *status_dev) {
    __shared__ float red[32];
    __shared__ float s_coef;
    __shared__ int s_bad;
    const int64_t step = *step_dev + 1;
    const int64_t stride = dp_stride(n);
    const int par = (int)(step & 1);
    float *own = bufs.buf[rank] + par * stride;
    if (grad_local) {   // publish: copy this rank's gradient into the slot of this step's parity (chosen on the device: graph-replayable)
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) own[i] = grad_local[i];
        __syncthreads();   // CTA-scope ordering; thread 0's system-scope fence + release below is cumulative over these writes
    }
    if (threadIdx.x == 0) {
        s_bad = 0;
        own[n] = loss_local ? (float)(*loss_local) : 0.f;
        __threadfence_system();
        unsigned long long *flag = reinterpret_cast<unsigned long long *>(bufs.buf[rank] + 2 * stride);
        st_release_sys(flag, (unsigned long long)step);
    }
    // wait for the peers: warp w polls peer w
    if ((threadIdx.x & 31) == 0) {
        const int peer = threadIdx.x >> 5;
        if (peer < world && peer != rank) {
            const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(bufs.buf[peer] + 2 * stride);
            long long spins = 0;
            while (ld_acquire_sys(flag) < (unsigned long long)step) {
                if (++spins > (1ll << 27)) { atomicExch(&s_bad, peer + 1); break; }
                __nanosleep(20);
            }
        }
    }
    __syncthreads();
    if (s_bad) { if (threadIdx.x == 0 && status_dev) *status_dev = s_bad; return; }   // do not touch the parameters on a failed exchange
    // gather + ordered sum (each thread owns up to 4 elements: n <= 4096 on this path)
    float g[4];
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int64_t i = threadIdx.x + (int64_t)e * blockDim.x;
        float acc = 0.f;
        if (i <= n) {
            // issue every peer read before the first add (independent NVLink loads in flight together), then sum in rank order
            float pv[DP_MAX_WORLD];
#pragma unroll
            for (int r = 0; r < DP_MAX_WORLD; ++r) pv[r] = r < world ? __ldcv(bufs.buf[r] + par * stride + i) : 0.f;
#pragma unroll
            for (int r = 0; r < DP_MAX_WORLD; ++r) if (r < world) acc += pv[r];
        }
        g[e] = acc;
        if (i < n) ss = fmaf(acc, acc, ss);
        if (i == n && loss_out) *loss_out = acc;
    }
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

using namespace odpd;

extern "C" {

int64_t odpd_dp_buffer_bytes(int64_t n_params) { return (2 * dp_stride(n_params) + 4) * (int64_t)sizeof(float); }

int odpd_dp_alloc(int64_t bytes, void **out_ptr) {
    ODPD_CHECK(out_ptr && bytes > 0, "odpd_dp_alloc: bad arguments");
    cudaError_t e = cudaMalloc(out_ptr, (size_t)bytes);
    ODPD_CHECK(e == cudaSuccess, "cudaMalloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
    e = cudaMemset(*out_ptr, 0, (size_t)bytes);
    ODPD_CHECK(e == cudaSuccess, "cudaMemset: %s", cudaGetErrorString(e));
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
    for (int r = 0; r < world; ++r) { ODPD_CHECK(bufs[r] != nullptr, "odpd_dp_clip_adamw: bufs[%d] is NULL", r); p.buf[r] = (float *)bufs[r]; }
    dp_clip_adamw_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(param, p, world, rank, n, grad_local, loss_local, exp_avg, exp_avg_sq, lr_dev, beta1, beta2,
                                                              eps, weight_decay, max_norm, step_dev, gnorm_out, loss_out, status_dev);
    return check_launch("dp_clip_adamw_kernel");
}

}  // extern "C"

// api.cu — the extern "C" surface of libodpd.so (include/odpd.h): argument checking, dispatch, error strings,
// the ordered gradient-partials reduction and the fused clip+AdamW step.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include "cells.h"
#include "chunking.cuh"

namespace odpd {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("ODPD_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return -2;
    }
    return 0;
}

// clip_grad_norm_ + AdamW on the flat buffers, executed by ONE CTA of any size (train_funcs.py:41-44, project.py:283):
// p *= 1-lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// `g` is read with L1-bypassing loads: in the fused form other CTAs of the same kernel wrote it.
__device__ __forceinline__ void clip_adamw_body(float *__restrict__ p, float *g, float *__restrict__ m, float *__restrict__ v, int64_t n,
                                                const float *__restrict__ lr_dev, float b1, float b2, float eps, float wd, float max_norm,
                                                int64_t *step_dev, float *gnorm_out, int zero_grad, float *red, float *s_coef) {
    float ss = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const float x = __ldcg(g + i); ss = fmaf(x, x, ss); }
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
            *s_coef = coef;
            if (gnorm_out) *gnorm_out = norm;
        }
    }
    __syncthreads();
    const float coef = *s_coef;
    const int64_t step = *step_dev + 1;
    const float lr = *lr_dev;
    const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
    const float step_size = lr / bc1, bc2s = sqrtf(bc2);
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float gi = __ldcg(g + i) * coef;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
        const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
        const float denom = sqrtf(vi) / bc2s + eps;
        pi -= step_size * (mi / denom);
        p[i] = pi; m[i] = mi; v[i] = vi;
        if (zero_grad) g[i] = 0.f; else g[i] = gi;
    }
    __syncthreads();
    if (threadIdx.x == 0) *step_dev = step;
}

// g[p] (+)= sum_{b=0..nrows-1} part[b][p].  Block = 32 parameters x RED_TY row groups: thread (tx,ty) sums rows ty, ty+RED_TY, ... in
// row order (8 independent loads in flight), the RED_TY group sums are then added in group order — a fixed summation tree, so the result
// is bit-reproducible run to run; no float atomics.
static constexpr int RED_TY = 16;
__global__ void __launch_bounds__(32 * RED_TY) reduce_partials_kernel(const float *__restrict__ part, int nrows, int64_t P, float *__restrict__ g,
                                                                      int overwrite, DpPushArgs push, AdamFuseArgs adam) {
    __shared__ float red[RED_TY][33];
    pdl_enter();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t p = (int64_t)blockIdx.x * 32 + tx;
    float acc = 0.f;
    if (p < P) {
        int b = ty;
        for (; b + 7 * RED_TY < nrows; b += 8 * RED_TY) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = part[(int64_t)(b + u * RED_TY) * P + p];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u];
        }
        for (; b < nrows; b += RED_TY) acc += part[(int64_t)b * P + p];
    }
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && p < P) {
        float s = red[0][tx];
#pragma unroll
        for (int k = 1; k < RED_TY; ++k) s += red[k][tx];
        s = overwrite ? s : g[p] + s;
        g[p] = s;
        if (push.world > 1) {      // data-parallel publish (dp.cu): {value, step tag} words into slot [parity][rank] of every rank
            const int64_t step = *push.step_dev + 1;
            const unsigned tag = (unsigned)step;
            const int64_t off = ((int64_t)(step & 1) * push.world + push.rank) * push.stride;
#pragma unroll
            for (int r = 0; r < ODPD_DP_MAX_WORLD; ++r)
                if (r < push.world) asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(push.buf[r] + off + p), "r"(__float_as_uint(s)), "r"(tag) : "memory");
            if (p == 0) {           // the loss rides along as element P
                const float lv = push.loss_local ? (float)(*push.loss_local) : 0.f;
#pragma unroll
                for (int r = 0; r < ODPD_DP_MAX_WORLD; ++r)
                    if (r < push.world) asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(push.buf[r] + off + P), "r"(__float_as_uint(lv)), "r"(tag) : "memory");
            }
        }
    }
    if (adam.p) {       // optimiser fused into the reduction: the last CTA to arrive sees every g[] (fence + atomic ticket)
        __shared__ int s_last;
        __shared__ float s_coef;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(adam.ticket, 1) == (int)gridDim.x - 1);
        __syncthreads();
        if (s_last) {
            __threadfence();
            clip_adamw_body(adam.p, g, adam.m, adam.v, P, adam.lr_dev, adam.b1, adam.b2, adam.eps, adam.wd, adam.max_norm, adam.step_dev, adam.gnorm_out,
                            0, &red[0][0], &s_coef);
            if (threadIdx.x == 0) *adam.ticket = 0;
        }
    }
}

int reduce_partials(const float *part, int nrows, int64_t P, float *g, int overwrite, cudaStream_t st, const DpPushArgs *push, const AdamFuseArgs *adam) {
    if (P <= 0) return 0;
    DpPushArgs pa{};
    if (push) pa = *push;
    AdamFuseArgs aa{};
    if (adam) aa = *adam;
    launch_pdl(reduce_partials_kernel, dim3((unsigned)((P + 31) / 32)), dim3(32 * RED_TY), 0, st, part, nrows, P, g, overwrite, pa, aa);
    return check_launch("reduce_partials_kernel");
}

// ---------------------------------------------------------------- clip_grad_norm_ + AdamW as its own launch (single CTA; n is ~1e3)
__global__ void __launch_bounds__(1024) clip_adamw_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m,
                                                          float *__restrict__ v, int64_t n, const float *__restrict__ lr_dev, float b1,
                                                          float b2, float eps, float wd, float max_norm, int64_t *step_dev,
                                                          float *gnorm_out, int zero_grad) {
    __shared__ float red[32];
    __shared__ float s_coef;
    pdl_enter();
    clip_adamw_body(p, g, m, v, n, lr_dev, b1, b2, eps, wd, max_norm, step_dev, gnorm_out, zero_grad, red, &s_coef);
}

// one-shot, thread-local: arms the next weight-gradient reduction of this host thread to run the optimiser in its last CTA
static thread_local AdamFuseArgs g_armed_adam;
static thread_local bool g_adam_armed = false;
static bool take_armed_adam(AdamFuseArgs &out) {
    if (!g_adam_armed) return false;
    out = g_armed_adam;
    g_adam_armed = false;
    return true;
}

static int check_dims(const OdpdDims *d) {
    ODPD_CHECK(d != nullptr, "dims is NULL");
    ODPD_CHECK(d->cell >= 0 && d->cell < ODPD_CELL_COUNT, "unknown cell %d", d->cell);
    ODPD_CHECK(d->B >= 0 && d->T >= 0, "negative B/T (%d,%d)", d->B, d->T);
    if (wide_supported(d->cell)) {
        ODPD_CHECK(d->H >= 1 && d->H <= 64, "hidden_size %d outside 1..64", d->H);
        ODPD_CHECK(d->K >= 0 && d->K <= 8, "num_layers %d outside 1..8", d->K);
    } else if (d->cell == ODPD_CELL_RVTDCNN || d->cell == ODPD_CELL_TCNN || d->cell == ODPD_CELL_NEURALTX) {
        ODPD_CHECK(d->H >= 1 && d->H <= 64, "hidden size %d outside 1..64", d->H);
    } else if (d->cell == ODPD_CELL_APNRRU) {
        ODPD_CHECK(d->H >= 1 && d->H <= 14, "APNRRU hidden_size %d outside 1..14", d->H);
    } else if (d->cell == ODPD_CELL_MCLDNN) {
        ODPD_CHECK(d->H >= 1 && d->H <= 12, "MCLDNN hidden size (conv channels) %d outside 1..12", d->H);
    } else if (d->cell == ODPD_CELL_DELTAJANET) {
        ODPD_CHECK(d->H >= 1 && d->H <= 16, "DeltaJANET hidden_size %d outside 1..16", d->H);
    } else if (d->cell == ODPD_CELL_TRES_QAT) {
        ODPD_CHECK(d->H >= 1 && d->H <= 16, "fake-quantised TRes-DeltaGRU hidden_size %d outside 1..16", d->H);
    } else if (d->cell == ODPD_CELL_BOJANET) {
        ODPD_CHECK(d->H >= 1 && d->H <= 18, "BOJANET hidden_size %d outside 1..18 (bojanet.py:41-52)", d->H);
    } else if (d->cell != ODPD_CELL_GMP) {
        ODPD_CHECK(d->H >= 1 && d->H <= 32, "hidden_size %d outside the fused range 1..32", d->H);
    }
    if (d->cell == ODPD_CELL_DVRJANET) ODPD_CHECK(d->K >= 1 && d->K <= 8, "num_dvr_units %d outside 1..8", d->K);
    return 0;
}

// GRU / LSTM / DGRU / QGRU / QGRU_AMP1: OdpdDims.K = num_layers (0 or 1 = one layer).  Hidden sizes above the fused tiers and stacked
// layers take the layered kernels of wide.cu.
static int wide_layers(int cell, int K) { return wide_supported(cell) && K > 1 ? K : 1; }
static bool is_wide(int cell, int H, int K) { return wide_supported(cell) && (H > 32 || K > 1); }
static GruArgs wide_args(const OdpdDims *d) {
    GruArgs a{};
    a.B = d->B; a.T = d->T; a.H = d->H; a.cell = d->cell;
    a.x_bf16 = (d->flags & ODPD_F_X_BF16) != 0; a.target_bf16 = (d->flags & ODPD_F_TARGET_BF16) != 0;
    a.x_starts = d->x_starts; a.target_starts = d->target_starts;
    return a;
}

static bool is_tcn(int c) { return c == ODPD_CELL_TCNN || c == ODPD_CELL_NEURALTX; }
static bool is_gru_family(int c) { return c == ODPD_CELL_GRU || c == ODPD_CELL_DGRU || c == ODPD_CELL_QGRU || c == ODPD_CELL_QGRU_AMP1; }

}  // namespace odpd

using namespace odpd;

extern "C" {

int odpd_version(void) { return ODPD_VERSION; }
const char *odpd_last_error(void) { return g_err; }

int64_t odpd_n_params(int32_t cell, int32_t H, int32_t K) {
    if (is_wide(cell, H, K)) return (H >= 1 && H <= 64 && K <= 8) ? wide_nparams(cell, H, wide_layers(cell, K)) : -1;
    if (cell == ODPD_CELL_RVTDCNN) return rvtdcnn_nparams(H);
    if (cell == ODPD_CELL_BOJANET) return bojanet_nparams(H);
    if (is_tcn(cell)) return tcnn_nparams(cell, H);
    if (cell == ODPD_CELL_APNRRU) return apnrru_nparams(H);
    if (cell == ODPD_CELL_MCLDNN) return mcldnn_nparams(H);
    if (cell == ODPD_CELL_DELTAJANET) return deltajanet_nparams(H);
    if (cell == ODPD_CELL_TRES_QAT) return tresq_nparams(H);
    if (is_gru_family(cell)) return gru_family_nparams(cell, H);
    return other_nparams(cell, H, K);
}

int64_t odpd_saved_bytes(const OdpdDims *d) {
    if (check_dims(d)) return -1;
    const bool save = (d->flags & ODPD_F_SAVE) != 0;
    if (is_wide(d->cell, d->H, d->K)) return 4 * wide_saved_floats(d->cell, d->B, d->T, d->H, wide_layers(d->cell, d->K), save);
    if (d->cell == ODPD_CELL_BOJANET) return 4 * bojanet_saved_floats(d->B, d->T, d->H);
    if (is_tcn(d->cell)) return save ? 4 * tcnn_saved_floats(d->B, d->T, d->H) : 0;
    if (d->cell == ODPD_CELL_APNRRU) return 4 * apnrru_saved_floats(d->B, d->T, d->H);
    if (d->cell == ODPD_CELL_MCLDNN) return 4 * mcldnn_saved_floats(d->B, d->T, d->H);
    if (d->cell == ODPD_CELL_DELTAJANET) return 4 * deltajanet_saved_floats(d->B, d->T, d->H);
    if (d->cell == ODPD_CELL_TRES_QAT) return save ? 4 * tresq_saved_floats(d->B, d->T) : 0;
    if (d->cell == ODPD_CELL_RVTDCNN) return save ? 16 : 0;      // nothing is saved (the backward recomputes); a token buffer keeps callers uniform
    if (is_gru_family(d->cell)) {
        const int64_t n = gru_family_saved_floats(d->cell, d->B, d->T, d->H, save, d->tchunks);
        if (n < 0) { set_error("GRU-family kernels support hidden_size <= 32 (got %d)", d->H); return -1; }
        return 4 * n;
    }
    return other_saved_bytes(d);
}

int64_t odpd_bwd_workspace_bytes(const OdpdDims *d) {
    if (check_dims(d)) return -1;
    if (is_wide(d->cell, d->H, d->K)) return 4 * wide_workspace_floats(d->cell, d->B, d->T, d->H, wide_layers(d->cell, d->K)) + 64;
    if (d->cell == ODPD_CELL_RVTDCNN) return 4 * rvtdcnn_workspace_floats(d->B, d->T, d->H) + 64;
    if (d->cell == ODPD_CELL_BOJANET) return 4 * bojanet_workspace_floats(d->B, d->T, d->H) + 64;
    if (is_tcn(d->cell)) return 4 * tcnn_workspace_floats(d->cell, d->B, d->T, d->H) + 64;
    if (d->cell == ODPD_CELL_APNRRU) return 4 * apnrru_workspace_floats(d->B, d->T, d->H) + 64;
    if (d->cell == ODPD_CELL_MCLDNN) return 4 * mcldnn_workspace_floats(d->B, d->T, d->H) + 64;
    if (d->cell == ODPD_CELL_DELTAJANET) return 4 * deltajanet_workspace_floats(d->B, d->T, d->H) + 64;
    if (d->cell == ODPD_CELL_TRES_QAT) return 4 * tresq_workspace_floats(d->B, d->T, d->H) + 64;
    if (is_gru_family(d->cell)) {
        const int64_t n = gru_family_workspace_floats(d->cell, d->B, d->T, d->H, d->tchunks);
        if (n < 0) { set_error("GRU-family kernels support hidden_size <= 32 (got %d)", d->H); return -1; }
        return 4 * n + 64;
    }
    const int64_t n = other_workspace_floats(d);
    if (n < 0) { set_error("cell %d: hidden_size %d is not supported", d->cell, d->H); return -1; }
    return 4 * n + 64;
}

int odpd_backbone_fwd(const OdpdDims *d, const float *x, const float *target, const float *params, float *out, double *loss,
                      double loss_scale, void *saved, int64_t *stats, void *stream) {
    if (check_dims(d)) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    if (loss && (d->flags & ODPD_F_ZERO_LOSS)) {
        cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(double), st);
        ODPD_CHECK(e == cudaSuccess, "cudaMemsetAsync(loss): %s", cudaGetErrorString(e));
    }
    if (d->B == 0 || d->T == 0) return 0;       // an empty batch (e.g. a data-parallel rank's empty shard): nothing to read or write
    ODPD_CHECK(params && out, "params/out must not be NULL");
    ODPD_CHECK(((uintptr_t)params & 15) == 0, "params must be 16-byte aligned");
    const bool save = (d->flags & ODPD_F_SAVE) != 0;
    ODPD_CHECK(!save || saved, "ODPD_F_SAVE set but saved==NULL");
    ODPD_CHECK(x != nullptr, "x must not be NULL");
    if (is_wide(d->cell, d->H, d->K)) {
        GruArgs a = wide_args(d);
        a.x = x; a.target = target; a.params = params; a.out = out; a.loss = loss;
        a.loss_scale = (float)loss_scale; a.saved = (float *)saved; a.save = save;
        return wide_run(d->cell, a, wide_layers(d->cell, d->K), 0, false, st, nullptr);
    }
    if (d->cell == ODPD_CELL_RVTDCNN) {
        GruArgs a = wide_args(d);
        a.x = x; a.target = target; a.params = params; a.out = out; a.loss = loss; a.loss_scale = (float)loss_scale;
        return rvtdcnn_run(a, 0, false, st, nullptr);
    }
    if (d->cell == ODPD_CELL_BOJANET) {
        GruArgs a = wide_args(d);
        a.x = x; a.target = target; a.params = params; a.out = out; a.loss = loss; a.loss_scale = (float)loss_scale; a.saved = (float *)saved;
        return bojanet_run(a, 0, false, st, nullptr);
    }
    if (is_tcn(d->cell)) {
        GruArgs a = wide_args(d);
        a.x = x; a.target = target; a.params = params; a.out = out; a.loss = loss; a.loss_scale = (float)loss_scale; a.saved = (float *)saved;
        a.save = save;
        return tcnn_run(d->cell, a, 0, false, st, nullptr);
    }
    if (d->cell == ODPD_CELL_APNRRU) {
        GruArgs a = wide_args(d);
        a.x = x; a.target = target; a.params = params; a.out = out; a.loss = loss; a.loss_scale = (float)loss_scale; a.saved = (float *)saved;
        return apnrru_run(a, 0, false, st, nullptr);
    }
    if (d->cell == ODPD_CELL_MCLDNN) {
        GruArgs a = wide_args(d);
        a.x = x; a.target = target; a.params = params; a.out = out; a.loss = loss; a.loss_scale = (float)loss_scale; a.saved = (float *)saved;
        return mcldnn_run(a, 0, false, st, nullptr);
    }
    if (d->cell == ODPD_CELL_DELTAJANET) {
        GruArgs a = wide_args(d);
        a.x = x; a.target = target; a.params = params; a.out = out; a.loss = loss; a.loss_scale = (float)loss_scale; a.saved = (float *)saved;
        return deltajanet_run(a, 0, false, st, nullptr);
    }
    if (d->cell == ODPD_CELL_TRES_QAT) {
        GruArgs a = wide_args(d);
        a.x = x; a.target = target; a.params = params; a.out = out; a.loss = loss; a.loss_scale = (float)loss_scale; a.saved = (float *)saved;
        a.save = save; a.K = d->K; a.thx = d->thx; a.thh = d->thh; a.stats = stats;
        return tresq_run(a, 0, false, st, nullptr);
    }
    if (is_gru_family(d->cell)) {
        GruArgs a{};
        a.B = d->B; a.T = d->T; a.H = d->H; a.x = x; a.target = target; a.params = params; a.out = out; a.loss = loss;
        a.loss_scale = (float)loss_scale; a.saved = (float *)saved; a.save = save;
        a.tchunks_req = d->tchunks; a.twarm_req = d->twarm;
        a.x_bf16 = (d->flags & ODPD_F_X_BF16) != 0; a.target_bf16 = (d->flags & ODPD_F_TARGET_BF16) != 0;
        a.x_starts = d->x_starts; a.target_starts = d->target_starts;
        return gru_family_run(d->cell, a, 0, false, st, nullptr);
    }
    return other_fwd(d, x, target, params, out, loss, loss_scale, saved, stats, st);
}

int odpd_backbone_bwd(const OdpdDims *d, const float *x, const float *params, const void *saved, const float *gout, const float *out,
                      const float *target, double gscale, const float *gscale_dev, float *gx, float *gparams, void *workspace,
                      void *stream) {
    if (check_dims(d)) return -1;
    const bool dx = (d->flags & ODPD_F_NEED_DX) != 0, dw = (d->flags & ODPD_F_NEED_DW) != 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t P = odpd_n_params(d->cell, d->H, d->K);
    ODPD_CHECK(!dw || gparams, "ODPD_F_NEED_DW set but gparams==NULL");
    if (d->B == 0 || d->T == 0) {
        // an empty shard (data-parallel rank whose share of the last partial batch is empty) contributes a ZERO gradient: with
        // OVERWRITE_DW the caller's buffer must not keep the previous step's values
        if (dw && (d->flags & ODPD_F_OVERWRITE_DW) && P > 0) {
            DpPushArgs push{};
            AdamFuseArgs adam{};
            const bool ap = dp_take_armed_push(push), aa = take_armed_adam(adam);
            if (ap || aa) return reduce_partials(nullptr, 0, P, gparams, 1, st, ap ? &push : nullptr, aa ? &adam : nullptr);   // zero gradient, still published / applied
            cudaError_t e = cudaMemsetAsync(gparams, 0, (size_t)P * sizeof(float), st);
            ODPD_CHECK(e == cudaSuccess, "cudaMemsetAsync(gparams): %s", cudaGetErrorString(e));
        }
        return 0;
    }
    ODPD_CHECK(params != nullptr, "params must not be NULL");
    ODPD_CHECK(((uintptr_t)params & 15) == 0, "params must be 16-byte aligned");
    ODPD_CHECK(!dx || gx, "ODPD_F_NEED_DX set but gx==NULL");
    ODPD_CHECK(!dw || workspace, "ODPD_F_NEED_DW set but workspace==NULL");
    ODPD_CHECK(gout || (out && target), "need gout, or out+target for the fused MSE gradient");
    if (!dx && !dw) return 0;
    ODPD_CHECK(x && (saved || d->cell == ODPD_CELL_GMP || d->cell == ODPD_CELL_RVTDCNN), "x/saved must not be NULL");
    int rc, rows = d->B;
    if (is_wide(d->cell, d->H, d->K)) {
        GruArgs a = wide_args(d);
        a.x = x; a.params = params; a.saved = (float *)saved; a.gout = gout; a.out_in = out; a.target = target;
        a.gscale = (float)gscale; a.gscale_dev = gscale_dev; a.gx = gx; a.partials = (float *)workspace; a.need_dx = dx;
        ODPD_CHECK(workspace != nullptr, "the layered RNN backward needs the workspace (odpd_bwd_workspace_bytes)");
        rc = wide_run(d->cell, a, wide_layers(d->cell, d->K), 1, dw, st, &rows);
    } else if (d->cell == ODPD_CELL_RVTDCNN) {
        GruArgs a = wide_args(d);
        a.x = x; a.params = params; a.gout = gout; a.out_in = out; a.target = target;
        a.gscale = (float)gscale; a.gscale_dev = gscale_dev; a.gx = gx; a.partials = (float *)workspace; a.need_dx = dx;
        rc = rvtdcnn_run(a, 1, dw, st, &rows);
    } else if (d->cell == ODPD_CELL_BOJANET) {
        GruArgs a = wide_args(d);
        a.x = x; a.params = params; a.saved = (float *)saved; a.gout = gout; a.out_in = out; a.target = target;
        a.gscale = (float)gscale; a.gscale_dev = gscale_dev; a.gx = gx; a.partials = (float *)workspace; a.need_dx = dx;
        ODPD_CHECK(workspace != nullptr, "the BOJANET backward needs the workspace (odpd_bwd_workspace_bytes)");
        rc = bojanet_run(a, 1, dw, st, &rows);
    } else if (is_tcn(d->cell)) {
        GruArgs a = wide_args(d);
        a.x = x; a.params = params; a.saved = (float *)saved; a.gout = gout; a.out_in = out; a.target = target;
        a.gscale = (float)gscale; a.gscale_dev = gscale_dev; a.gx = gx; a.partials = (float *)workspace; a.need_dx = dx;
        rc = tcnn_run(d->cell, a, 1, dw, st, &rows);
    } else if (d->cell == ODPD_CELL_APNRRU) {
        GruArgs a = wide_args(d);
        a.x = x; a.params = params; a.saved = (float *)saved; a.gout = gout; a.out_in = out; a.target = target;
        a.gscale = (float)gscale; a.gscale_dev = gscale_dev; a.gx = gx; a.partials = (float *)workspace; a.need_dx = dx;
        ODPD_CHECK(workspace != nullptr, "the APNRRU backward needs the workspace (odpd_bwd_workspace_bytes)");
        rc = apnrru_run(a, 1, dw, st, &rows);
    } else if (d->cell == ODPD_CELL_MCLDNN) {
        GruArgs a = wide_args(d);
        a.x = x; a.params = params; a.saved = (float *)saved; a.gout = gout; a.out_in = out; a.target = target;
        a.gscale = (float)gscale; a.gscale_dev = gscale_dev; a.gx = gx; a.partials = (float *)workspace; a.need_dx = dx;
        ODPD_CHECK(workspace != nullptr, "the MCLDNN backward needs the workspace (odpd_bwd_workspace_bytes)");
        rc = mcldnn_run(a, 1, dw, st, &rows);
    } else if (d->cell == ODPD_CELL_DELTAJANET) {
        GruArgs a = wide_args(d);
        a.x = x; a.params = params; a.saved = (float *)saved; a.gout = gout; a.out_in = out; a.target = target;
        a.gscale = (float)gscale; a.gscale_dev = gscale_dev; a.gx = gx; a.partials = (float *)workspace; a.need_dx = dx;
        ODPD_CHECK(workspace != nullptr, "the DeltaJANET backward needs the workspace (odpd_bwd_workspace_bytes)");
        rc = deltajanet_run(a, 1, dw, st, &rows);
    } else if (d->cell == ODPD_CELL_TRES_QAT) {
        GruArgs a = wide_args(d);
        a.x = x; a.params = params; a.saved = (float *)saved; a.gout = gout; a.out_in = out; a.target = target;
        a.gscale = (float)gscale; a.gscale_dev = gscale_dev; a.gx = gx; a.partials = (float *)workspace; a.need_dx = dx;
        a.K = d->K; a.thx = d->thx; a.thh = d->thh;
        ODPD_CHECK(workspace != nullptr, "the fake-quantised TRes-DeltaGRU backward needs the workspace (odpd_bwd_workspace_bytes)");
        rc = tresq_run(a, 1, dw, st, &rows);
    } else if (is_gru_family(d->cell)) {
        GruArgs a{};
        a.B = d->B; a.T = d->T; a.H = d->H; a.x = x; a.params = params; a.saved = (float *)saved; a.gout = gout; a.out_in = out;
        a.target = target; a.gscale = (float)gscale; a.gscale_dev = gscale_dev; a.gx = gx; a.partials = (float *)workspace;
        a.need_dx = dx;
        a.tchunks_req = d->tchunks; a.twarm_req = d->twarm;
        a.x_bf16 = (d->flags & ODPD_F_X_BF16) != 0; a.target_bf16 = (d->flags & ODPD_F_TARGET_BF16) != 0;
        a.x_starts = d->x_starts; a.target_starts = d->target_starts;
        rc = gru_family_run(d->cell, a, 1, dw, st, &rows);
    } else {
        rc = other_bwd(d, x, params, saved, gout, out, target, gscale, gscale_dev, gx, (float *)workspace, st, &rows);
    }
    if (rc) return rc;
    if (dw) {
        DpPushArgs push{};
        AdamFuseArgs adam{};
        const bool armed = dp_take_armed_push(push), aa = take_armed_adam(adam);
        return reduce_partials((const float *)workspace, rows, P, gparams, (d->flags & ODPD_F_OVERWRITE_DW) != 0, st, armed ? &push : nullptr,
                               aa ? &adam : nullptr);
    }
    return 0;
}

int odpd_chunk_plan_model(int32_t B, int32_t T, int32_t tchunks, int32_t twarm, int32_t slots, int32_t default_warm, int32_t out[3]) {
    ODPD_CHECK(out != nullptr && B >= 0 && T >= 0, "odpd_chunk_plan_model: bad arguments");
    GruArgs a{};
    a.B = B; a.T = T; a.tchunks_req = tchunks; a.twarm_req = twarm; a.twarm_default = default_warm;
    chunk_make_plan(a, slots, true);
    out[0] = a.C; out[1] = a.Lc; out[2] = a.Wu;
    return 0;
}

int odpd_chunk_plan(const OdpdDims *d, int32_t backward, int32_t out[4]) {
    if (check_dims(d)) return -1;
    ODPD_CHECK(out != nullptr, "out is NULL");
    out[0] = 1; out[1] = d->T; out[2] = 0; out[3] = -1;
    if (d->B == 0 || d->T == 0 || d->cell == ODPD_CELL_GMP || d->cell == ODPD_CELL_RVTDCNN || d->cell == ODPD_CELL_BOJANET || d->cell == ODPD_CELL_APNRRU || d->cell == ODPD_CELL_MCLDNN || d->cell == ODPD_CELL_DELTAJANET || d->cell == ODPD_CELL_TRES_QAT || is_tcn(d->cell) || is_wide(d->cell, d->H, d->K)) return 0;
    int info[4];
    if (!is_gru_family(d->cell)) {
        const int rc = other_plan(d, backward, info);
        if (rc) return rc;
        for (int i = 0; i < 4; ++i) out[i] = info[i];
        if (d->tchunks == 1) out[3] = -1;
        return 0;
    }
    const int rc = gru_family_plan(d->cell, d->B, d->T, d->H, d->tchunks, d->twarm, backward ? 1 : 0, (d->flags & ODPD_F_NEED_DW) != 0,
                                   (d->flags & ODPD_F_SAVE) != 0, info);
    if (rc) return rc;
    for (int i = 0; i < 4; ++i) out[i] = info[i];
    if (d->tchunks == 1) out[3] = -1;
    return 0;
}

int odpd_fuse_next_bwd_with_adamw(float *param, float *exp_avg, float *exp_avg_sq, const float *lr_dev, float beta1, float beta2, float eps,
                                  float weight_decay, float max_norm, int64_t *step_dev, float *gnorm_out, int32_t *ticket_dev) {
    ODPD_CHECK(param && exp_avg && exp_avg_sq && lr_dev && step_dev && ticket_dev, "odpd_fuse_next_bwd_with_adamw: NULL buffer");
    AdamFuseArgs a{};
    a.p = param; a.m = exp_avg; a.v = exp_avg_sq; a.lr_dev = lr_dev; a.step_dev = step_dev; a.gnorm_out = gnorm_out; a.ticket = ticket_dev;
    a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.wd = weight_decay; a.max_norm = max_norm;
    g_armed_adam = a;
    g_adam_armed = true;
    return 0;
}

int odpd_clip_adamw(float *param, float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, const float *lr_dev, float beta1,
                    float beta2, float eps, float weight_decay, float max_norm, int64_t *step_dev, float *gnorm_out, int zero_grad,
                    void *stream) {
    ODPD_CHECK(param && grad && exp_avg && exp_avg_sq && lr_dev && step_dev, "odpd_clip_adamw: NULL buffer");
    ODPD_CHECK(n >= 0, "odpd_clip_adamw: negative n");
    if (n == 0) return 0;
    launch_pdl(clip_adamw_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n, lr_dev, beta1, beta2, eps,
               weight_decay, max_norm, step_dev, gnorm_out, zero_grad);
    return check_launch("clip_adamw_kernel");
}

}  // extern "C"

// rvtdcnn.cu — RVTDCNN backbone: real-valued time-delay CNN, forward / backward (+ fused I/Q MSE).
//
// Replaces (reference, file:line): backbones/rvtdcnn.py:9-62 —
//   features (I, Q, |x|, |x|^2, |x|^3) per sample (:41-46); a 4-sample memory window per timestep whose first three taps WRAP to the end
//   of the frame (`pad = x[:, -(window_size-1):, :]`, :51-53: tap r of window t is sample (t-3+r) mod T); Conv2d(1 -> 3 channels, 3x3,
//   padding (1,0)) over the 4x5 window image -> tanh -> 36 values (channel, row, column) (:57-58); fc_hid(36 -> H) -> tanh (:59);
//   fc_out(H -> 2) (:60).  H = fc_hid_size = the CLI's hidden size (models.py:80-81).
// No recurrence: one thread per timestep, weights staged once per CTA in shared memory (broadcast reads); nothing is saved for the
// backward — it recomputes the 36 + H activations of its timestep (cheaper than 600 bytes/sample of HBM traffic).  Weight gradients:
// each tile of 64 timesteps parks its per-step factors in shared memory and the CTA's threads then own one parameter each and sum the
// outer products over the tile (K = time), accumulating into the CTA's gradient-partial row; rows are reduced in order by
// reduce_partials_kernel.  dL/dx: one (dI,dQ) contribution per (timestep, tap), summed per sample by a gather pass in tap order
// (bit-reproducible, no atomics).
//
// Flat parameter layout (named_parameters() order):  Conv2d.weight(3,1,3,3) Conv2d.bias(3) fc_hid.weight(H,36) fc_hid.bias(H)
// fc_out.weight(2,H) fc_out.bias(2)   = 32 + 39 H.
#include <mutex>
#include "cells.h"
#include "chunking.cuh"

namespace odpd {

static constexpr int RV_TT = 64;        // timesteps per tile = threads per CTA
static constexpr int RV_NZ = 36;        // conv outputs per timestep: 3 channels x 4 rows x 3 columns
static constexpr int RV_HMAX = 64;

struct RvLayout {
    int H, oWc, obc, oWh, obh, oWo, obo, P;
    __host__ __device__ explicit RvLayout(int h) {
        H = h; oWc = 0; obc = 27; oWh = 30; obh = oWh + RV_NZ * h; oWo = obh + h; obo = oWo + 2 * h; P = obo + 2;
    }
};
// shared staging of the parameters with 16-byte aligned blocks:  Wc[27] bc[3] pad2 | Wh[H][36] | bh[H] Wo[2][H] bo[2]
__host__ __device__ inline int rv_param_floats(int H) { return 32 + RV_NZ * H + ((3 * H + 2 + 3) & ~3); }

__device__ __forceinline__ void rv_stage(float *sp, const float *P, const RvLayout &L, int tid, int nthreads) {
    for (int i = tid; i < 30; i += nthreads) sp[i] = __ldg(P + i);
    for (int i = tid; i < RV_NZ * L.H; i += nthreads) sp[32 + i] = __ldg(P + L.oWh + i);
    for (int i = tid; i < 3 * L.H + 2; i += nthreads) sp[32 + RV_NZ * L.H + i] = __ldg(P + L.obh + i);
}

// window of timestep t: w[r][c], r = tap (sample (t-3+r) mod T), c = feature (I, Q, |x|, |x|^2, |x|^3)
__device__ __forceinline__ void rv_window(const IqRow &x2, int t, int T, float (&w)[4][5], float (&iq)[4][2]) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int s = (t - 3 + r) % T;
        if (s < 0) s += T;
        const float2 v = x2.ld(s);
        const float a2 = __fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y));
        const float a = __fsqrt_rn(a2);
        w[r][0] = v.x; w[r][1] = v.y; w[r][2] = a; w[r][3] = a2; w[r][4] = __fmul_rn(__fmul_rn(a, a), a);
        iq[r][0] = v.x; iq[r][1] = v.y;
    }
}
// z[ch*12 + r*3 + cc] = tanh(bc[ch] + sum_{dr,dc} Wc[ch][dr][dc] * w[r+dr-1][cc+dc])   (rows outside 0..3 are the zero padding)
__device__ __forceinline__ void rv_conv(const float *sp, const float (&w)[4][5], float (&z)[RV_NZ]) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float bias = sp[27 + ch];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                float acc = bias;
#pragma unroll
                for (int dr = 0; dr < 3; ++dr) {
                    const int rr = r + dr - 1;
                    if (rr < 0 || rr > 3) continue;
#pragma unroll
                    for (int dc = 0; dc < 3; ++dc) acc = fmaf(sp[ch * 9 + dr * 3 + dc], w[rr][cc + dc], acc);
                }
                z[ch * 12 + r * 3 + cc] = tanhf_(acc);
            }
    }
}
__device__ __forceinline__ float rv_hid_pre(const float *sp, int H, int k, const float (&z)[RV_NZ]) {
    const float4 *wr = reinterpret_cast<const float4 *>(sp + 32 + k * RV_NZ);
    float a0 = sp[32 + RV_NZ * H + k], a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int q = 0; q < RV_NZ / 4; ++q) {
        const float4 v = wr[q];
        a0 = fmaf(v.x, z[4 * q], a0); a1 = fmaf(v.y, z[4 * q + 1], a1); a2 = fmaf(v.z, z[4 * q + 2], a2); a3 = fmaf(v.w, z[4 * q + 3], a3);
    }
    return (a0 + a1) + (a2 + a3);
}

// ================================================================ forward
__global__ void __launch_bounds__(RV_TT) rvtdcnn_fwd_kernel(GruArgs a, int nts, int ntiles) {
    pdl_enter();
    const RvLayout L(a.H);
    const int H = a.H, T = a.T, tid = threadIdx.x;
    extern __shared__ __align__(16) float rsm[];
    float *sp = rsm;
    __shared__ float sred[RV_TT / 32];
    rv_stage(sp, a.params, L, tid, RV_TT);
    __syncthreads();
    const float *sWo = sp + 32 + RV_NZ * H + H;
    const float bo0 = sWo[2 * H], bo1 = sWo[2 * H + 1];
    float lsum = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / nts, t = (tile - b * nts) * RV_TT + tid;
        if (t >= T) continue;
        const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
        float w[4][5], iq[4][2], z[RV_NZ];
        rv_window(x2, t, T, w, iq);
        rv_conv(sp, w, z);
        float o0 = bo0, o1 = bo1;
        for (int k = 0; k < H; ++k) {
            const float hk = tanhf_(rv_hid_pre(sp, H, k, z));
            o0 = fmaf(sWo[k], hk, o0);
            o1 = fmaf(sWo[H + k], hk, o1);
        }
        reinterpret_cast<float2 *>(a.out)[(size_t)b * T + t] = make_float2(o0, o1);
        if (a.target) {
            const float2 y = iq_row(a.target, a.target_bf16, a.target_starts, b, T).ld(t);
            const float d0 = o0 - y.x, d1 = o1 - y.y;
            lsum = fmaf(d0, d0, fmaf(d1, d1, lsum));
        }
    }
    if (a.loss && a.target) {
        lsum = warp_sum(lsum);
        if ((tid & 31) == 0) sred[tid >> 5] = lsum;
        __syncthreads();
        if (tid == 0) atomicAdd(a.loss, (double)(sred[0] + sred[1]) * (double)a.loss_scale);
    }
}

// ================================================================ backward
// per-tile shared factors (odd pitches: thread t writes row t):  da[t][H] | hk[t][H] | z[t][36] | dcp[t][36] | w[t][20] | go[t][2]
template <bool DW>
__global__ void __launch_bounds__(RV_TT) rvtdcnn_bwd_kernel(GruArgs a, int nts, int ntiles, float2 *contrib) {
    pdl_enter();
    const RvLayout L(a.H);
    const int H = a.H, T = a.T, tid = threadIdx.x;
    const int HPi = H | 1;
    extern __shared__ __align__(16) float rsm[];
    float *sp = rsm;
    float *sda = sp + rv_param_floats(H);
    float *shk = sda + RV_TT * HPi;
    float *sz = shk + RV_TT * HPi;
    float *sdc = sz + RV_TT * 37;
    float *sw = sdc + RV_TT * 37;
    float *sgo = sw + RV_TT * 21;
    rv_stage(sp, a.params, L, tid, RV_TT);
    __syncthreads();
    const float *sWo = sp + 32 + RV_NZ * H + H;
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
    float *prt = (DW && a.partials) ? a.partials + (size_t)blockIdx.x * L.P : nullptr;
    bool first = true;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / nts, t = (tile - b * nts) * RV_TT + tid;
        const bool valid = t < T;
        float w[4][5], iq[4][2], z[RV_NZ], dz[RV_NZ];
        float2 go = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < RV_NZ; ++m) { z[m] = 0.f; dz[m] = 0.f; }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int c = 0; c < 5; ++c) w[r][c] = 0.f;
            iq[r][0] = iq[r][1] = 0.f;
        }
        if (valid) {
            const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
            rv_window(x2, t, T, w, iq);
            rv_conv(sp, w, z);
            if (a.gout) go = __ldg(reinterpret_cast<const float2 *>(a.gout) + (size_t)b * T + t);
            else {
                const float2 o = __ldg(reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * T + t);
                const float2 y = iq_row(a.target, a.target_bf16, a.target_starts, b, T).ld(t);
                go = make_float2(gs * (o.x - y.x), gs * (o.y - y.y));
            }
        }
        for (int k = 0; k < H; ++k) {
            float hk = 0.f, da = 0.f;
            if (valid) {
                hk = tanhf_(rv_hid_pre(sp, H, k, z));
                da = fmaf(go.x, sWo[k], go.y * sWo[H + k]) * (1.f - hk * hk);
                const float4 *wr = reinterpret_cast<const float4 *>(sp + 32 + k * RV_NZ);
#pragma unroll
                for (int q = 0; q < RV_NZ / 4; ++q) {
                    const float4 v = wr[q];
                    dz[4 * q] = fmaf(da, v.x, dz[4 * q]); dz[4 * q + 1] = fmaf(da, v.y, dz[4 * q + 1]);
                    dz[4 * q + 2] = fmaf(da, v.z, dz[4 * q + 2]); dz[4 * q + 3] = fmaf(da, v.w, dz[4 * q + 3]);
                }
            }
            if constexpr (DW) { sda[tid * HPi + k] = da; shk[tid * HPi + k] = hk; }
        }
        // through the tanh of the conv outputs
#pragma unroll
        for (int m = 0; m < RV_NZ; ++m) dz[m] *= 1.f - z[m] * z[m];
        if constexpr (DW) {
#pragma unroll
            for (int m = 0; m < RV_NZ; ++m) { sz[tid * 37 + m] = z[m]; sdc[tid * 37 + m] = dz[m]; }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) sw[tid * 21 + r * 5 + c] = w[r][c];
            sgo[tid * 3] = go.x; sgo[tid * 3 + 1] = go.y;
        }
        if (a.need_dx && contrib && valid) {
            // transposed convolution onto the window, then the feature Jacobian of each tap
            float dw[4][5];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) dw[r][c] = 0.f;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        const float d = dz[ch * 12 + r * 3 + cc];
#pragma unroll
                        for (int dr = 0; dr < 3; ++dr) {
                            const int rr = r + dr - 1;
                            if (rr < 0 || rr > 3) continue;
#pragma unroll
                            for (int dc = 0; dc < 3; ++dc) dw[rr][cc + dc] = fmaf(d, sp[ch * 9 + dr * 3 + dc], dw[rr][cc + dc]);
                        }
                    }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float i = iq[r][0], q = iq[r][1], am = w[r][2], a2 = w[r][3];
                // |x|^2 = I^2+Q^2, |x| = sqrt(|x|^2), |x|^3:   d/dI = g_I + 2 I g_a2 + (g_a + 3 |x|^2 g_a3) I/|x|
                const float ga = fmaf(3.f * a2, dw[r][4], dw[r][2]);
                const float sc = fmaf(2.f, dw[r][3], ga / am);
                contrib[((size_t)b * T + t) * 4 + r] = make_float2(fmaf(i, sc, dw[r][0]), fmaf(q, sc, dw[r][1]));
            }
        }
        if constexpr (DW) {
            __syncthreads();
            if (prt) {
                for (int o = tid; o < L.P; o += RV_TT) {
                    float s = 0.f;
                    if (o < 27) {                                   // Conv2d.weight[ch][0][dr][dc]
                        const int ch = o / 9, dr = (o - ch * 9) / 3, dc = o - ch * 9 - dr * 3;
                        for (int tt = 0; tt < RV_TT; ++tt)
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const int rr = r + dr - 1;
                                if (rr < 0 || rr > 3) continue;
#pragma unroll
                                for (int cc = 0; cc < 3; ++cc) s = fmaf(sdc[tt * 37 + ch * 12 + r * 3 + cc], sw[tt * 21 + rr * 5 + cc + dc], s);
                            }
                    } else if (o < 30) {                            // Conv2d.bias[ch]
                        const int ch = o - 27;
                        for (int tt = 0; tt < RV_TT; ++tt)
#pragma unroll
                            for (int m = 0; m < 12; ++m) s += sdc[tt * 37 + ch * 12 + m];
                    } else if (o < L.obh) {                         // fc_hid.weight[k][m]
                        const int k = (o - L.oWh) / RV_NZ, m = (o - L.oWh) - k * RV_NZ;
                        for (int tt = 0; tt < RV_TT; ++tt) s = fmaf(sda[tt * HPi + k], sz[tt * 37 + m], s);
                    } else if (o < L.oWo) {                         // fc_hid.bias[k]
                        const int k = o - L.obh;
                        for (int tt = 0; tt < RV_TT; ++tt) s += sda[tt * HPi + k];
                    } else if (o < L.obo) {                         // fc_out.weight[c][k]
                        const int c = (o - L.oWo) / H, k = (o - L.oWo) - c * H;
                        for (int tt = 0; tt < RV_TT; ++tt) s = fmaf(sgo[tt * 3 + c], shk[tt * HPi + k], s);
                    } else {                                        // fc_out.bias[c]
                        const int c = o - L.obo;
                        for (int tt = 0; tt < RV_TT; ++tt) s += sgo[tt * 3 + c];
                    }
                    prt[o] = first ? s : prt[o] + s;
                }
            }
            first = false;
            __syncthreads();
        }
    }
    if constexpr (DW) {
        if (prt && first)                                          // a CTA without tiles still owns a (zero) row
            for (int o = tid; o < L.P; o += RV_TT) prt[o] = 0.f;
    }
}

// dL/dx[s] = sum over the four window taps that touch sample s: timestep (s+3-r) mod T, tap r   (fixed order: bit-reproducible)
__global__ void rvtdcnn_gather_kernel(const float2 *__restrict__ ctr, float2 *__restrict__ gx, int B, int T) {
    pdl_enter();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * T) return;
    const int b = (int)(i / T), s = (int)(i - (int64_t)b * T);
    float gi = 0.f, gq = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int t = (s + 3 - r) % T;
        const float2 v = ctr[((size_t)b * T + t) * 4 + r];
        gi += v.x; gq += v.y;
    }
    gx[i] = make_float2(gi, gq);
}

// ================================================================ host
static int rv_grid(int B, int T) {
    const int64_t tiles = (int64_t)B * ((T + RV_TT - 1) / RV_TT);
    const int64_t cap = 8 * (int64_t)num_sms();
    return (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
}
int64_t rvtdcnn_nparams(int H) { return RvLayout(H).P; }
// workspace = gradient partials [rows][P] (4-aligned) | per-(timestep, tap) dL/dx contributions [B][T][4] float2
int64_t rvtdcnn_workspace_floats(int B, int T, int H) {
    const int64_t bt = (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1);
    return (((int64_t)rv_grid(B, T) * RvLayout(H).P + 3) & ~(int64_t)3) + bt * 8 + 4;
}

static void rv_ensure_smem(const void *k, size_t bytes) {
    static std::mutex mu;
    static size_t have[3][ODPD_MAX_DEV] = {};
    static const void *ks[3] = {nullptr, nullptr, nullptr};
    std::lock_guard<std::mutex> lock(mu);
    int slot = -1;
    for (int i = 0; i < 3; ++i) {
        if (ks[i] == k) { slot = i; break; }
        if (!ks[i]) { ks[i] = k; slot = i; break; }
    }
    const int dev = cur_dev_slot();
    if (slot < 0 || bytes > have[slot][dev]) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (slot >= 0) have[slot][dev] = bytes;
    }
}

int rvtdcnn_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out) {
    if (a.H < 1 || a.H > RV_HMAX) { set_error("RVTDCNN: fc_hid_size %d outside 1..%d", a.H, RV_HMAX); return -1; }
    if (a.T < 3) { set_error("RVTDCNN needs frame_length >= 3 (the reference's wrap-around window, rvtdcnn.py:51-53; got %d)", a.T); return -1; }
    const int nts = (a.T + RV_TT - 1) / RV_TT, ntiles = a.B * nts, grid = rv_grid(a.B, a.T);
    const size_t psm = (size_t)rv_param_floats(a.H) * sizeof(float);
    if (dir == 0) {
        rv_ensure_smem((const void *)rvtdcnn_fwd_kernel, psm);
        launch_pdl(rvtdcnn_fwd_kernel, dim3(grid), dim3(RV_TT), psm, st, a, nts, ntiles);
        return check_launch("rvtdcnn_fwd_kernel");
    }
    const RvLayout L(a.H);
    float2 *contrib = nullptr;
    if (a.need_dx) {
        if (!a.partials || !a.gx) { set_error("RVTDCNN backward: dX needs the workspace and gx"); return -1; }
        contrib = reinterpret_cast<float2 *>(a.partials + (((int64_t)grid * L.P + 3) & ~(int64_t)3));
    }
    const size_t bsm = psm + (size_t)RV_TT * (2 * (a.H | 1) + 37 + 37 + 21 + 3) * sizeof(float);
    if (dw) {
        rv_ensure_smem((const void *)rvtdcnn_bwd_kernel<true>, bsm);
        launch_pdl(rvtdcnn_bwd_kernel<true>, dim3(grid), dim3(RV_TT), bsm, st, a, nts, ntiles, contrib);
    } else {
        rv_ensure_smem((const void *)rvtdcnn_bwd_kernel<false>, bsm);
        launch_pdl(rvtdcnn_bwd_kernel<false>, dim3(grid), dim3(RV_TT), bsm, st, a, nts, ntiles, contrib);
    }
    if (a.need_dx) {
        const int64_t n = (int64_t)a.B * a.T;
        launch_pdl(rvtdcnn_gather_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, (const float2 *)contrib, reinterpret_cast<float2 *>(a.gx), a.B, a.T);
    }
    if (rows_out) *rows_out = grid;
    return check_launch("rvtdcnn backward");
}

}  // namespace odpd

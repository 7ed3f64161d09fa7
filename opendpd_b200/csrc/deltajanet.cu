// deltajanet.cu — DeltaJANET backbone, forward / backward (+ fused I/Q MSE).
//
// Replaces (reference, file:line): backbones/deltajanet.py:10-274 —
//   features (I,Q,|x|,|x|^3,sin,cos) (:49-57); the layer is constructed with thx = thh = 0 whatever the caller passes (:22-26), so every delta
//   passes its threshold and the "previous" states follow every step:  dx_t = f_t - f_{t-1},  dh_t = h_{t-1} - h_{t-2}  (zero before the frame);
//   M += W_ih dx + W_hh dh, M_0 = b_ih + b_hh (:160-167, :194-202);  f = sigmoid(M_f), g = sigmoid(M_g) (:246-247);  h = (1-f) g + f h (:248);
//   out = fc_out(h) (:59).  The accumulator keeps the reference's summation order (W_ih dx + M) + W_hh dh: the cell is a JANET computed
//   incrementally, and its fp32 rounding history is part of what "same result as the reference" means here.
// hidden_size <= 16 (lane = (gate, unit) of one warp), one layer.
//
//   forward   front (one thread per timestep: features of steps t and t-1, XD = W_ih dx) -> chain (one warp per sequence; M in a register per
//             lane, dh broadcast by shuffles) -> head (one thread per timestep)
//   backward  head_bwd (dL/dh) -> chain_bwd (reverse; the adjoint of the accumulator M is a RUNNING sum over later steps — stored per step as GM;
//             dh_t = h_{t-1} - h_{t-2} feeds two earlier states, handled with one pending term) -> post (one thread per timestep:
//             dL/df_t = W_ih^T (GM_t - GM_{t+1}) -> dL/dx; weight gradients as tile outer products GM_t (x) dx_t, GM_t (x) dh_t, biases = GM_0)
//
// Flat parameter layout: rnn.weight_ih_l0(2H,6) weight_hh_l0(2H,H) bias_ih_l0(2H) bias_hh_l0(2H) fc_out.weight(2,H) fc_out.bias(2) = 2H^2 + 18H + 2.
#include "cells.h"
#include "chunking.cuh"

namespace odpd {

static constexpr int DJ_TT = 64;
static constexpr int DJ_HMAX = 16;

struct DjLayout {
    int H, G, oWih, oWhh, obih, obhh, oWo, obo, P;
    __host__ __device__ explicit DjLayout(int h) {
        H = h; G = 2 * h; oWih = 0; oWhh = G * 6; obih = oWhh + G * h; obhh = obih + G; oWo = obhh + G; obo = oWo + 2 * h; P = obo + 2;
    }
};
// saved: XD [B][T][2H] | ACT [B][T][3H] = f | g | h          workspace: partials | DH [B][T][H] | GM [B][T][2H]
struct DjBufs { float *xd, *act, *dh, *gm, *partials; };

__device__ __forceinline__ void dj_tile(int tile, int nts, int tid, int &b, int &t) {
    b = tile / nts;
    t = (tile - b * nts) * DJ_TT + tid;
}
__device__ __forceinline__ void dj_feat(const IqRow &x2, int t, float *f) {
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = 0.f;
    if (t < 0) return;
    const float2 v = x2.ld(t);
    features_fwd<FM_DGRU6>(v.x, v.y, 0.f, 0.f, f);
}

// ================================================================ forward: XD = W_ih (f_t - f_{t-1})
__global__ void __launch_bounds__(DJ_TT) dj_front_kernel(GruArgs a, DjBufs u, int nts, int ntiles) {
    pdl_enter();
    const DjLayout L(a.H);
    const int G = L.G, T = a.T, tid = threadIdx.x;
    __shared__ float sW[2 * DJ_HMAX * 6];
    for (int i = tid; i < G * 6; i += DJ_TT) sW[i] = __ldg(a.params + i);
    __syncthreads();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        dj_tile(tile, nts, tid, b, t);
        if (t >= T) continue;
        const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
        float f1[8], f0[8], dx[6];
        dj_feat(x2, t, f1);
        dj_feat(x2, t - 1, f0);
#pragma unroll
        for (int k = 0; k < 6; ++k) dx[k] = f1[k] - f0[k];
        float *xd = u.xd + ((size_t)b * T + t) * G;
        for (int r = 0; r < G; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 6; ++k) acc = fmaf(sW[r * 6 + k], dx[k], acc);
            xd[r] = acc;
        }
    }
}

// ================================================================ forward: the cell, one warp per sequence, lane = (gate g = lane >> 4, unit j = lane & 15)
__global__ void __launch_bounds__(128) dj_chain_fwd_kernel(GruArgs a, DjBufs u) {
    pdl_enter();
    const DjLayout L(a.H);
    const int H = a.H, G = L.G, T = a.T, lane = threadIdx.x & 31, g = lane >> 4, j = lane & 15;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= a.B) return;
    const bool act = j < H;
    const int r = g * H + (act ? j : 0);
    float wh[DJ_HMAX];
#pragma unroll
    for (int k = 0; k < DJ_HMAX; ++k) wh[k] = (act && k < H) ? __ldg(a.params + L.oWhh + r * H + k) : 0.f;
    float M = act ? __ldg(a.params + L.obih + r) + __ldg(a.params + L.obhh + r) : 0.f;
    const float *xd = u.xd + (size_t)b * T * G + r;
    float *arow = u.act + (size_t)b * T * 3 * H;
    float h1 = 0.f, h2 = 0.f;               // h_{t-1}, h_{t-2} of unit j (both half-warps carry them)
    float cx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) cx[i] = (i < T && act) ? __ldg(xd + (size_t)i * G) : 0.f;
    for (int t0 = 0; t0 < T; t0 += 4) {
        float nx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) nx[i] = (t0 + 4 + i < T && act) ? __ldg(xd + (size_t)(t0 + 4 + i) * G) : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = t0 + i;
            if (t < T) {
                const float dh = h1 - h2;
                float d0 = 0.f, d1 = 0.f;
#pragma unroll
                for (int k = 0; k < DJ_HMAX; k += 2) {
                    d0 = fmaf(wh[k], __shfl_sync(ODPD_FULL, dh, k), d0);
                    d1 = fmaf(wh[k + 1], __shfl_sync(ODPD_FULL, dh, k + 1), d1);
                }
                M = (cx[i] + M) + (d0 + d1);
                const float gate = sigmoidf_(M);
                const float fg = __shfl_sync(ODPD_FULL, gate, j), gg = __shfl_sync(ODPD_FULL, gate, 16 + j);
                const float hn = act ? fmaf(fg, h1 - gg, gg) : 0.f;         // (1-f) g + f h
                h2 = h1; h1 = hn;
                if (act) {
                    float *row = arow + (size_t)t * 3 * H;
                    row[g * H + j] = gate;
                    if (g == 0) row[2 * H + j] = hn;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) cx[i] = nx[i];
    }
}

// ================================================================ forward: out = fc_out(h) + squared error
__global__ void __launch_bounds__(DJ_TT) dj_head_fwd_kernel(GruArgs a, DjBufs u, int nts, int ntiles) {
    pdl_enter();
    const DjLayout L(a.H);
    const int H = a.H, T = a.T, tid = threadIdx.x;
    __shared__ float sWo[2 * DJ_HMAX + 2], sred[DJ_TT / 32];
    for (int i = tid; i < 2 * H + 2; i += DJ_TT) sWo[i] = __ldg(a.params + L.oWo + i);
    __syncthreads();
    float lsum = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        dj_tile(tile, nts, tid, b, t);
        if (t >= T) continue;
        const float *hp = u.act + ((size_t)b * T + t) * 3 * H + 2 * H;
        float o0 = sWo[2 * H], o1 = sWo[2 * H + 1];
#pragma unroll
        for (int k = 0; k < DJ_HMAX; ++k)
            if (k < H) { const float hv = hp[k]; o0 = fmaf(sWo[k], hv, o0); o1 = fmaf(sWo[H + k], hv, o1); }
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

__device__ __forceinline__ float2 dj_go(const GruArgs &a, int b, int t, float gs) {
    if (a.gout) return __ldg(reinterpret_cast<const float2 *>(a.gout) + (size_t)b * a.T + t);
    const float2 o = __ldg(reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * a.T + t);
    const float2 y = iq_row(a.target, a.target_bf16, a.target_starts, b, a.T).ld(t);
    return make_float2(gs * (o.x - y.x), gs * (o.y - y.y));
}

// ================================================================ backward: dL/dh of the head
__global__ void __launch_bounds__(DJ_TT) dj_head_bwd_kernel(GruArgs a, DjBufs u, int nts, int ntiles) {
    pdl_enter();
    const DjLayout L(a.H);
    const int H = a.H, T = a.T, tid = threadIdx.x;
    __shared__ float sWo[2 * DJ_HMAX];
    for (int i = tid; i < 2 * H; i += DJ_TT) sWo[i] = __ldg(a.params + L.oWo + i);
    __syncthreads();
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        dj_tile(tile, nts, tid, b, t);
        if (t >= T) continue;
        const float2 go = dj_go(a, b, t, gs);
        float *d = u.dh + ((size_t)b * T + t) * H;
        for (int k = 0; k < H; ++k) d[k] = fmaf(go.x, sWo[k], go.y * sWo[H + k]);
    }
}

// ================================================================ backward: the cell in reverse, one warp per sequence.  GM[b][t][2H] = running adjoint of M
__global__ void __launch_bounds__(128) dj_chain_bwd_kernel(GruArgs a, DjBufs u) {
    pdl_enter();
    const DjLayout L(a.H);
    const int H = a.H, G = L.G, T = a.T, lane = threadIdx.x & 31, g = lane >> 4, j = lane & 15;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= a.B) return;
    const bool act = j < H;
    const int jj = act ? j : 0;
    float wc[2 * DJ_HMAX];                  // column j of W_hh: wc[q] for row r(q) = (q >> 4) * H + (q & 15)
#pragma unroll
    for (int q = 0; q < 2 * DJ_HMAX; ++q) wc[q] = (act && (q & 15) < H) ? __ldg(a.params + L.oWhh + ((q >> 4) * H + (q & 15)) * H + jj) : 0.f;
    const float *arow = u.act + (size_t)b * T * 3 * H;
    const float *dhp = u.dh + (size_t)b * T * H + jj;
    float *gmp = u.gm + (size_t)b * T * G + g * H + jj;
    float gM = 0.f, gH = 0.f, pend = 0.f;   // gM: lane (g, j);  gH, pend: unit j (both half-warps)
    struct In { float dh, f, g, hp; };
    auto fetch = [&](int t) {
        In q{};
        if (t < 0 || !act) return q;
        const float *row = arow + (size_t)t * 3 * H;
        q.dh = __ldg(dhp + (size_t)t * H); q.f = __ldg(row + jj); q.g = __ldg(row + H + jj);
        q.hp = t > 0 ? __ldg(row - 3 * H + 2 * H + jj) : 0.f;
        return q;
    };
    auto step = [&](int t, const In &q) {
        const float gh = q.dh + gH;
        gM += g == 0 ? gh * (q.hp - q.g) * q.f * (1.f - q.f) : gh * (1.f - q.f) * q.g * (1.f - q.g);
        if (act) gmp[(size_t)t * G] = gM;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int p = 0; p < 2 * DJ_HMAX; p += 2) {
            a0 = fmaf(wc[p], __shfl_sync(ODPD_FULL, gM, p), a0);
            a1 = fmaf(wc[p + 1], __shfl_sync(ODPD_FULL, gM, p + 1), a1);
        }
        const float at = a0 + a1;                                   // dL/d(dh_t)[j]
        gH = act ? fmaf(gh, q.f, at + pend) : 0.f;                  // adjoint of h_{t-1}: direct + dh_t - dh_{t+1} (the head's share is q.dh of step t-1)
        pend = -at;
    };
    In A = fetch(T - 1), Bq = fetch(T - 2);
    for (int t = T - 1; t >= 0; t -= 2) {
        const In An = fetch(t - 2);
        step(t, A);
        const In Bn = fetch(t - 3);
        if (t - 1 >= 0) step(t - 1, Bq);
        A = An; Bq = Bn;
    }
}

// ================================================================ backward: post.  dL/dx and the weight gradients
// per-tile shared factors (odd pitches):  gm[t][2H] | dx[t][7] | dhv[t][H] | h[t][H] | go[t][3]
template <bool DW>
__global__ void __launch_bounds__(DJ_TT) dj_post_kernel(GruArgs a, DjBufs u, int nts, int ntiles) {
    pdl_enter();
    const DjLayout L(a.H);
    const int H = a.H, G = L.G, T = a.T, tid = threadIdx.x;
    const int GP = G | 1, HPi = H | 1;
    __shared__ float sW[2 * DJ_HMAX * 6];
    extern __shared__ __align__(16) float dsm[];
    float *sGm = dsm;                      // [64][GP]
    float *sDx = sGm + DJ_TT * GP;         // [64][7]
    float *sDh = sDx + DJ_TT * 7;          // [64][HPi]
    float *sHh = sDh + DJ_TT * HPi;        // [64][HPi]
    float *sGo = sHh + DJ_TT * HPi;        // [64][3]
    float *sB0 = sGo + DJ_TT * 3;          // [GP]  GM of step 0 when the tile holds it (bias gradient)
    for (int i = tid; i < G * 6; i += DJ_TT) sW[i] = __ldg(a.params + i);
    __syncthreads();
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
    float *prt = (DW && u.partials) ? u.partials + (size_t)blockIdx.x * L.P : nullptr;
    bool first = true;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        dj_tile(tile, nts, tid, b, t);
        const bool valid = t < T;
        const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
        float f1[8], f0[8];
        dj_feat(x2, valid ? t : -1, f1);
        dj_feat(x2, valid ? t - 1 : -1, f0);
        if (valid && a.need_dx && a.gx) {
            const float *g0 = u.gm + ((size_t)b * T + t) * G;
            float gf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int r = 0; r < G; ++r) {
                const float d = g0[r] - (t + 1 < T ? g0[G + r] : 0.f);
#pragma unroll
                for (int k = 0; k < 6; ++k) gf[k] = fmaf(d, sW[r * 6 + k], gf[k]);
            }
            const float2 v = x2.ld(t);
            float gi, gq;
            features_bwd<FM_DGRU6>(v.x, v.y, gf, gi, gq);
            reinterpret_cast<float2 *>(a.gx)[(size_t)b * T + t] = make_float2(gi, gq);
        }
        if constexpr (DW) {
            const float *row = u.act + ((size_t)b * T + t) * 3 * H;
            for (int r = 0; r < G; ++r) sGm[tid * GP + r] = valid ? u.gm[((size_t)b * T + t) * G + r] : 0.f;
#pragma unroll
            for (int k = 0; k < 6; ++k) sDx[tid * 7 + k] = f1[k] - f0[k];
            for (int k = 0; k < H; ++k) {
                const float h1 = (valid && t > 0) ? row[-3 * H + 2 * H + k] : 0.f, h2 = (valid && t > 1) ? row[-6 * H + 2 * H + k] : 0.f;
                sDh[tid * HPi + k] = h1 - h2;
                sHh[tid * HPi + k] = valid ? row[2 * H + k] : 0.f;
            }
            float2 go = make_float2(0.f, 0.f);
            if (valid) go = dj_go(a, b, t, gs);
            sGo[tid * 3] = go.x; sGo[tid * 3 + 1] = go.y;
            const bool has0 = (t - tid) == 0;             // the tile that starts the sequence carries the bias gradient GM_0
            __syncthreads();
            if (prt) {
                for (int o = tid; o < L.P; o += DJ_TT) {
                    float s = 0.f;
                    if (o < L.oWhh) {                     // W_ih[r][k]
                        const int r = o / 6, k = o - r * 6;
                        for (int tt = 0; tt < DJ_TT; ++tt) s = fmaf(sGm[tt * GP + r], sDx[tt * 7 + k], s);
                    } else if (o < L.obih) {              // W_hh[r][k]
                        const int r = (o - L.oWhh) / H, k = (o - L.oWhh) - r * H;
                        for (int tt = 0; tt < DJ_TT; ++tt) s = fmaf(sGm[tt * GP + r], sDh[tt * HPi + k], s);
                    } else if (o < L.oWo) {               // b_ih[r], b_hh[r]: the adjoint of M_0
                        const int r = (o - L.obih) % G;
                        s = has0 ? sGm[r] : 0.f;
                    } else if (o < L.obo) {               // fc_out.weight[c][k]
                        const int c = (o - L.oWo) / H, k = (o - L.oWo) - c * H;
                        for (int tt = 0; tt < DJ_TT; ++tt) s = fmaf(sGo[tt * 3 + c], sHh[tt * HPi + k], s);
                    } else {                              // fc_out.bias[c]
                        for (int tt = 0; tt < DJ_TT; ++tt) s += sGo[tt * 3 + o - L.obo];
                    }
                    prt[o] = first ? s : prt[o] + s;
                }
            }
            first = false;
            __syncthreads();
        }
    }
    (void)sB0;
    if constexpr (DW) {
        if (prt && first)
            for (int o = tid; o < L.P; o += DJ_TT) prt[o] = 0.f;
    }
}

// ================================================================ host
static int dj_grid(int B, int T) {
    const int64_t tiles = (int64_t)B * ((T + DJ_TT - 1) / DJ_TT);
    const int64_t cap = 8 * (int64_t)num_sms();
    return (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
}
int64_t deltajanet_nparams(int H) { return DjLayout(H).P; }
int64_t deltajanet_saved_floats(int B, int T, int H) { return (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1) * 5 * H + 4; }
int64_t deltajanet_workspace_floats(int B, int T, int H) {
    const int64_t bt = (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1);
    return (((int64_t)dj_grid(B, T) * DjLayout(H).P + 3) & ~(int64_t)3) + bt * 3 * H + 4;
}

int deltajanet_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out) {
    if (a.H < 1 || a.H > DJ_HMAX) { set_error("DeltaJANET: hidden_size %d outside 1..%d", a.H, DJ_HMAX); return -1; }
    if (!a.saved) { set_error("DeltaJANET needs the `saved` buffer (odpd_saved_bytes), also without ODPD_F_SAVE"); return -1; }
    const DjLayout L(a.H);
    const int H = a.H, nts = (a.T + DJ_TT - 1) / DJ_TT, ntiles = a.B * nts, grid = dj_grid(a.B, a.T), wpc = a.B <= 2 * num_sms() ? 1 : 4 /* few sequences: one chain warp per CTA spreads them over the SMs */, cgrid = (a.B + wpc - 1) / wpc;
    const int64_t bt = (int64_t)a.B * a.T;
    DjBufs u{};
    u.xd = a.saved; u.act = a.saved + bt * 2 * H;
    if (dir == 0) {
        launch_pdl(dj_front_kernel, dim3(grid), dim3(DJ_TT), 0, st, a, u, nts, ntiles);
        launch_pdl(dj_chain_fwd_kernel, dim3(cgrid), dim3(32 * wpc), 0, st, a, u);
        launch_pdl(dj_head_fwd_kernel, dim3(grid), dim3(DJ_TT), 0, st, a, u, nts, ntiles);
        return check_launch("deltajanet forward");
    }
    if (!a.partials) { set_error("DeltaJANET backward needs the workspace (odpd_bwd_workspace_bytes)"); return -1; }
    if (a.need_dx && !a.gx) { set_error("DeltaJANET backward: ODPD_F_NEED_DX without gx"); return -1; }
    const int64_t poff = ((int64_t)grid * L.P + 3) & ~(int64_t)3;
    u.partials = a.partials; u.dh = a.partials + poff; u.gm = u.dh + bt * H;
    launch_pdl(dj_head_bwd_kernel, dim3(grid), dim3(DJ_TT), 0, st, a, u, nts, ntiles);
    launch_pdl(dj_chain_bwd_kernel, dim3(cgrid), dim3(32 * wpc), 0, st, a, u);
    const size_t psm = (size_t)(DJ_TT * (((2 * H) | 1) + 7 + 2 * (H | 1) + 3) + ((2 * H) | 1)) * sizeof(float);
    if (dw) launch_pdl(dj_post_kernel<true>, dim3(grid), dim3(DJ_TT), psm, st, a, u, nts, ntiles);
    else if (a.need_dx) launch_pdl(dj_post_kernel<false>, dim3(grid), dim3(DJ_TT), psm, st, a, u, nts, ntiles);
    if (rows_out) *rows_out = grid;
    return check_launch("deltajanet backward");
}

}  // namespace odpd

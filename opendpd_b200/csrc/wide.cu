// wide.cu — layered forward / backward for the nn.GRU / nn.LSTM based backbones outside the fused tiers:
// hidden_size 33..64 and/or num_layers > 1 (GRU, LSTM, DGRU, QGRU, QGRU_AMP1).
//
// Replaces (reference, file:line): backbones/gru.py:17-24,45-48, lstm.py:17-24,45-48, dgru.py:22-33,59-74, qgru.py:22-31,59-71,
// qgru_amp1.py:59-76 with `num_layers` passed through to torch.nn.GRU/LSTM (weight_ih_l{k}, weight_hh_l{k}, bias_ih_l{k},
// bias_hh_l{k}; layer k > 0 reads the hidden sequence of layer k-1) and hidden sizes the fused lane-per-unit kernels
// (gru_family.cu, lstm.cu: H <= 32, one layer) do not cover.  arguments.py:51,60 offer both on the command line.
//
// Structure.  Only the recurrence is serial, so each layer is split into time-parallel GEMM-shaped kernels around one chain kernel:
//   forward   xproj(l)      XP = W_ih in + b            (all timesteps at once; in = features(IQ) for l = 0, else h of layer l-1)
//             chain_fwd(l)  one CTA per sequence, G*64 threads: thread (gate g, unit j) keeps row g*H+j of W_hh in registers (64 floats),
//                           h in shared memory (double-buffered), two CTA barriers per timestep
//             head_fwd      fc_out (GRU/LSTM/QGRU) or relu(fc_hid) -> fc_out([g; features]) (DGRU) + squared error
//   backward  head_bwd      dL/dout -> dL/dh of the top layer, head weight gradients
//             chain_bwd(l)  reverse recurrence: thread (g, k) keeps column k of W_hh gate g in registers; writes the per-step gate
//                           gradients G
//             wgrad(l)      dW_hh = G^T h_prev, dW_ih = G^T in, biases   (register-tiled, K = time)
//             dx(l)         dL/din = G W_ih  -> dL/dh of layer l-1, or (l = 0) through the feature Jacobian to dL/dx
// Gradient partials are written one row per CTA and reduced in row order by reduce_partials_kernel (api.cu), as everywhere else.
// No time chunking here (odpd_chunk_plan reports 1 chunk): this path is about coverage of the reference's command line; the
// sizes its scripts of record use (H <= 32, one layer) take the fused kernels.
#include <mutex>
#include "cells.h"
#include "chunking.cuh"

namespace odpd {

static constexpr int WH = 64;       // widest hidden size
static constexpr int WLMAX = 8;     // most layers
static constexpr int TT = 32;       // timesteps per tile of the time-parallel kernels
static constexpr int TP = 36;       // row pitch of transposed [k][t] tiles (16-byte aligned rows, spreads banks)
static constexpr int VS = 72;       // row pitch of the head's input rows  [g or h (<=64) ; features (<=6)]

struct WideLayout {
    int G, H, F, L, head, O, P;
    int oWih[WLMAX], oWhh[WLMAX], obih[WLMAX], obhh[WLMAX];
    int oWo, obo, oWh, obh;
};

struct WideArgs {
    WideLayout L;
    int B, T, nts, ntiles, NS, layer, save, need_dx;
    const float *params;
    const void *x, *target;
    int x_bf16, target_bf16;
    const int *x_starts, *target_starts;
    const float *gout, *out_in, *gscale_dev;
    float gscale, loss_scale;
    float *out, *gx;
    double *loss;
    float *act, *xp, *dh, *gb, *partials;
};

__host__ __device__ inline int al4(int n) { return (n + 3) & ~3; }

static bool wide_cell_info(int cell, int &G, int &FM, int &head) {
    switch (cell) {
    case ODPD_CELL_GRU: G = 3; FM = FM_RAW2; head = 0; return true;
    case ODPD_CELL_LSTM: G = 4; FM = FM_RAW2; head = 0; return true;
    case ODPD_CELL_DGRU: G = 3; FM = FM_DGRU6; head = 1; return true;
    case ODPD_CELL_QGRU: G = 3; FM = FM_QGRU4; head = 0; return true;
    case ODPD_CELL_QGRU_AMP1: G = 3; FM = FM_AMP4; head = 0; return true;
    }
    return false;
}

static WideLayout make_layout(int cell, int H, int layers) {
    WideLayout L{};
    int FM = 0;
    wide_cell_info(cell, L.G, FM, L.head);
    L.H = H; L.L = layers;
    L.F = FM == FM_RAW2 ? 2 : ((FM == FM_QGRU4 || FM == FM_AMP4) ? 4 : 6);
    int off = 0;
    for (int l = 0; l < layers; ++l) {
        const int fin = l == 0 ? L.F : H;
        L.oWih[l] = off; off += L.G * H * fin;
        L.oWhh[l] = off; off += L.G * H * H;
        L.obih[l] = off; off += L.G * H;
        L.obhh[l] = off; off += L.G * H;
    }
    L.O = L.head ? H + L.F : H;
    L.oWo = off; off += 2 * L.O;
    L.obo = off; off += 2;
    L.oWh = off; L.obh = off;
    if (L.head) { off += H * H; L.obh = off; off += H; }
    L.P = off;
    return L;
}

__device__ __forceinline__ void tile_of(const WideArgs &w, int tile, int &b, int &t0, int &tv) {
    b = tile / w.nts;
    t0 = (tile - b * w.nts) * TT;
    tv = w.T - t0 < TT ? w.T - t0 : TT;
}
__device__ __forceinline__ const float *act_of(const WideArgs &w, int layer, int b) {
    return w.act + ((size_t)layer * w.B + b) * (size_t)w.T * (w.NS * w.L.H);
}

// ================================================================ forward: input projection of one layer, all timesteps
template <int FM>
__global__ void __launch_bounds__(256) wide_xproj_kernel(WideArgs w) {
    pdl_enter();
    constexpr int F = FeatN<FM>::value;
    const WideLayout &L = w.L;
    const int l = w.layer, H = L.H, GH = L.G * H, Fin = l == 0 ? F : H, NSH = w.NS * H, tid = threadIdx.x;
    extern __shared__ __align__(16) float wsm[];
    float *sWt = wsm;                       // [Fin][GH]  W_ih, k-major
    float *sb = sWt + al4(Fin * GH);        // [GH]       b_ih (+ b_hh where it is not inside r*(.))
    float *sinT = sb + al4(GH);             // [Fin][TP]  inputs of the tile, transposed
    const float *P = w.params;
    for (int i = tid; i < Fin * GH; i += 256) { const int o = i / Fin, k = i - o * Fin; sWt[k * GH + o] = __ldg(P + L.oWih[l] + i); }
    for (int o = tid; o < GH; o += 256) sb[o] = __ldg(P + L.obih[l] + o) + ((L.G == 3 && o >= 2 * H) ? 0.f : __ldg(P + L.obhh[l] + o));
    __syncthreads();
    for (int tile = blockIdx.x; tile < w.ntiles; tile += gridDim.x) {
        int b, t0, tv;
        tile_of(w, tile, b, t0, tv);
        if (l == 0) {
            if (tid < TT) {
                float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (tid < tv) {
                    const float2 v = iq_row(w.x, w.x_bf16, w.x_starts, b, w.T).ld(t0 + tid);
                    features_fwd<FM>(v.x, v.y, 0.f, 0.f, f);
                }
#pragma unroll
                for (int m = 0; m < F; ++m) sinT[m * TP + tid] = f[m];
            }
        } else {
            const float *hrow = act_of(w, l - 1, b) + (size_t)t0 * NSH;
            for (int i = tid; i < TT * H; i += 256) {
                const int t = i / H, k = i - t * H;
                sinT[k * TP + t] = t < tv ? __ldg(hrow + (size_t)t * NSH + k) : 0.f;
            }
        }
        __syncthreads();
        if (tid < GH) {
            float acc[TT];
            const float bias = sb[tid];
#pragma unroll
            for (int t = 0; t < TT; ++t) acc[t] = bias;
            for (int k = 0; k < Fin; ++k) {
                const float wv = sWt[k * GH + tid];
                const float4 *r = reinterpret_cast<const float4 *>(sinT + k * TP);
#pragma unroll
                for (int q = 0; q < TT / 4; ++q) {
                    const float4 v = r[q];
                    acc[4 * q] = fmaf(wv, v.x, acc[4 * q]); acc[4 * q + 1] = fmaf(wv, v.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(wv, v.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(wv, v.w, acc[4 * q + 3]);
                }
            }
            float *dst = w.xp + ((size_t)b * w.T + t0) * GH + tid;
#pragma unroll
            for (int t = 0; t < TT; ++t)
                if (t < tv) dst[(size_t)t * GH] = acc[t];
        }
        __syncthreads();
    }
}

// ================================================================ forward: the recurrence of one layer, one CTA per sequence
// GRU (G=3): ATen cell, gate order r,z,n;  n = tanh(W_in x + b_in + r*(W_hn h + b_hn));  h' = (h-n)*z + n.
// LSTM (G=4): gate order i,f,g,o;  c' = f*c + i*g;  h' = o*tanh(c').
// activation row per step (NS*H floats):  GRU  h | r | z | n | hgn      LSTM  h | i | f | g | o | c      (only h without ODPD_F_SAVE)
template <int G>
__global__ void __launch_bounds__(G * 64) wide_chain_fwd_kernel(WideArgs w) {
    pdl_enter();
    const WideLayout &L = w.L;
    const int l = w.layer, H = L.H, T = w.T, GH = G * H, NSH = w.NS * H;
    const int tid = threadIdx.x, g = tid >> 6, j = tid & 63, b = blockIdx.x;
    const bool act = j < H;
    __shared__ __align__(16) float sh[2][WH];
    __shared__ float sgate[G][WH];
    float W[WH];
    {
        const float *Whh = w.params + L.oWhh[l] + (size_t)(g * H + (act ? j : 0)) * H;
#pragma unroll
        for (int k = 0; k < WH; ++k) W[k] = (act && k < H) ? __ldg(Whh + k) : 0.f;
    }
    const float bhn = (G == 3 && g == 2 && act) ? __ldg(w.params + L.obhh[l] + 2 * H + j) : 0.f;
    if (tid < 2 * WH) (&sh[0][0])[tid] = 0.f;
    __syncthreads();
    const float *xp = w.xp + (size_t)b * T * GH + g * H + (act ? j : 0);
    float *arow = w.act + ((size_t)l * w.B + b) * (size_t)T * NSH;
    float h = 0.f, c = 0.f;
    int cur = 0;
    float xq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xq[i] = (i < T && act) ? __ldg(xp + (size_t)i * GH) : 0.f;
    for (int t0 = 0; t0 < T; t0 += 8) {
        float xn[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int tt = t0 + 8 + i;
            xn[i] = (tt < T && act) ? __ldg(xp + (size_t)tt * GH) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = t0 + i;
            if (t < T) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                const float4 *hv = reinterpret_cast<const float4 *>(sh[cur]);
#pragma unroll
                for (int q = 0; q < WH / 4; ++q) {
                    const float4 v = hv[q];
                    a0 = fmaf(W[4 * q], v.x, a0); a1 = fmaf(W[4 * q + 1], v.y, a1);
                    a2 = fmaf(W[4 * q + 2], v.z, a2); a3 = fmaf(W[4 * q + 3], v.w, a3);
                }
                const float dot = (a0 + a1) + (a2 + a3);
                float hgn = 0.f;
                if constexpr (G == 3) {
                    if (g < 2) sgate[g][j] = sigmoidf_(xq[i] + dot);
                    else hgn = dot + bhn;
                } else {
                    sgate[g][j] = g == 2 ? tanhf_(xq[i] + dot) : sigmoidf_(xq[i] + dot);
                }
                __syncthreads();
                if constexpr (G == 3) {
                    if (g == 2) {
                        const float r = sgate[0][j], z = sgate[1][j];
                        const float n = tanhf_(fmaf(r, hgn, xq[i]));
                        h = act ? fmaf(h - n, z, n) : 0.f;
                        sh[cur ^ 1][j] = h;
                        if (act) {
                            float *row = arow + (size_t)t * NSH;
                            row[j] = h;
                            if (w.save) { row[H + j] = r; row[2 * H + j] = z; row[3 * H + j] = n; row[4 * H + j] = hgn; }
                        }
                    }
                } else {
                    if (g == 0) {
                        const float ig = sgate[0][j], fg = sgate[1][j], gg = sgate[2][j], og = sgate[3][j];
                        c = act ? fmaf(fg, c, ig * gg) : 0.f;
                        h = act ? og * tanhf_(c) : 0.f;
                        sh[cur ^ 1][j] = h;
                        if (act) {
                            float *row = arow + (size_t)t * NSH;
                            row[j] = h;
                            if (w.save) { row[H + j] = ig; row[2 * H + j] = fg; row[3 * H + j] = gg; row[4 * H + j] = og; row[5 * H + j] = c; }
                        }
                    }
                }
                __syncthreads();
                cur ^= 1;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) xq[i] = xn[i];
    }
}

// ================================================================ forward: output head + squared error, all timesteps
// HEAD 0: out = fc_out(h)   (gru.py:47, lstm.py:47, qgru.py:70)        HEAD 1: out = fc_out([relu(fc_hid(h)); features])   (dgru.py:71-73)
template <int HEAD>
__device__ __forceinline__ void head_fc_hid(const float *sWhT, const float *sbh, const float *shT, int H, int tid, float (&acc)[8]) {
    const int j = tid & 63, tg = tid >> 6;
    const float bias = sbh[j];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bias;
    for (int k = 0; k < H; ++k) {
        const float wv = sWhT[k * WH + j];
        const float4 a = *reinterpret_cast<const float4 *>(shT + k * TP + tg * 8), c = *reinterpret_cast<const float4 *>(shT + k * TP + tg * 8 + 4);
        acc[0] = fmaf(wv, a.x, acc[0]); acc[1] = fmaf(wv, a.y, acc[1]); acc[2] = fmaf(wv, a.z, acc[2]); acc[3] = fmaf(wv, a.w, acc[3]);
        acc[4] = fmaf(wv, c.x, acc[4]); acc[5] = fmaf(wv, c.y, acc[5]); acc[6] = fmaf(wv, c.z, acc[6]); acc[7] = fmaf(wv, c.w, acc[7]);
    }
}

// stage the head's weights:  sWo [2][VS],  HEAD: sWhT [H][64] (k-major), sWhr [H][64] (row-major, backward only), sbh [64]
template <int HEAD>
__device__ __forceinline__ void head_stage(const WideArgs &w, float *sWo, float *sWhT, float *sWhr, float *sbh, int tid) {
    const WideLayout &L = w.L;
    const int H = L.H;
    for (int i = tid; i < 2 * VS; i += 256) { const int c = i / VS, m = i - c * VS; sWo[i] = m < L.O ? __ldg(w.params + L.oWo + c * L.O + m) : 0.f; }
    if constexpr (HEAD == 1) {
        for (int i = tid; i < WH * WH; i += 256) {
            const int jj = i >> 6, k = i & 63;     // row jj, column k of fc_hid.weight
            const float v = (jj < H && k < H) ? __ldg(w.params + L.oWh + jj * H + k) : 0.f;
            sWhT[k * WH + jj] = v;
            if (sWhr) sWhr[jj * WH + k] = v;
        }
        for (int i = tid; i < WH; i += 256) sbh[i] = i < H ? __ldg(w.params + L.obh + i) : 0.f;
    }
}

template <int FM, int HEAD>
__global__ void __launch_bounds__(256) wide_head_fwd_kernel(WideArgs w) {
    pdl_enter();
    constexpr int F = FeatN<FM>::value;
    const WideLayout &L = w.L;
    const int H = L.H, NSH = w.NS * H, tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
    extern __shared__ __align__(16) float wsm[];
    float *shT = wsm;                   // [64][TP]   h tile, transposed (HEAD 1)
    float *sv = shT + WH * TP;          // [TT][VS]   head input rows
    float *sWo = sv + TT * VS;          // [2][VS]
    float *sbh = sWo + 2 * VS;          // [64]
    float *sWhT = sbh + WH;             // [64][64]   (HEAD 1)
    __shared__ float sred[8];
    head_stage<HEAD>(w, sWo, sWhT, nullptr, sbh, tid);
    const float bo0 = __ldg(w.params + L.obo), bo1 = __ldg(w.params + L.obo + 1);
    float lsum = 0.f;
    __syncthreads();
    for (int tile = blockIdx.x; tile < w.ntiles; tile += gridDim.x) {
        int b, t0, tv;
        tile_of(w, tile, b, t0, tv);
        const float *hrow = act_of(w, L.L - 1, b) + (size_t)t0 * NSH;
        for (int i = tid; i < TT * H; i += 256) {
            const int t = i / H, k = i - t * H;
            const float v = t < tv ? __ldg(hrow + (size_t)t * NSH + k) : 0.f;
            if constexpr (HEAD == 1) shT[k * TP + t] = v; else sv[t * VS + k] = v;
        }
        if constexpr (HEAD == 1) {
            if (tid < TT) {
                float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (tid < tv) {
                    const float2 v = iq_row(w.x, w.x_bf16, w.x_starts, b, w.T).ld(t0 + tid);
                    features_fwd<FM>(v.x, v.y, 0.f, 0.f, f);
                }
#pragma unroll
                for (int m = 0; m < F; ++m) sv[tid * VS + H + m] = f[m];
            }
        }
        __syncthreads();
        if constexpr (HEAD == 1) {
            float acc[8];
            head_fc_hid<HEAD>(sWhT, sbh, shT, H, tid, acc);
            const int j = tid & 63, tg = tid >> 6;
            if (j < H) {
#pragma unroll
                for (int i = 0; i < 8; ++i) sv[(tg * 8 + i) * VS + j] = fmaxf(acc[i], 0.f);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = wi * 4 + i;
            float p0 = 0.f, p1 = 0.f;
            for (int m = lane; m < L.O; m += 32) {
                const float v = sv[t * VS + m];
                p0 = fmaf(sWo[m], v, p0);
                p1 = fmaf(sWo[VS + m], v, p1);
            }
            p0 = warp_sum(p0); p1 = warp_sum(p1);
            if (lane == 0 && t < tv) {
                const float o0 = p0 + bo0, o1 = p1 + bo1;
                reinterpret_cast<float2 *>(w.out)[(size_t)b * w.T + t0 + t] = make_float2(o0, o1);
                if (w.target) {
                    const float2 y = iq_row(w.target, w.target_bf16, w.target_starts, b, w.T).ld(t0 + t);
                    const float d0 = o0 - y.x, d1 = o1 - y.y;
                    lsum = fmaf(d0, d0, fmaf(d1, d1, lsum));
                }
            }
        }
        __syncthreads();
    }
    if (w.loss && w.target) {
        if (lane == 0) sred[wi] = lsum;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += sred[i];
            atomicAdd(w.loss, (double)s * (double)w.loss_scale);
        }
    }
}

// ================================================================ backward: head.  dL/dout -> dL/dh (top layer) + head weight gradients
// go[t] = gout[t]  or  gscale * (out - target)   (the fused nn.MSELoss gradient, project.py:262-272)
__device__ __forceinline__ float2 wide_go(const WideArgs &w, int b, int t, float gs) {
    if (w.gout) return __ldg(reinterpret_cast<const float2 *>(w.gout) + (size_t)b * w.T + t);
    const float2 o = __ldg(reinterpret_cast<const float2 *>(w.out_in) + (size_t)b * w.T + t);
    const float2 y = iq_row(w.target, w.target_bf16, w.target_starts, b, w.T).ld(t);
    return make_float2(gs * (o.x - y.x), gs * (o.y - y.y));
}

template <int FM, int HEAD, bool DW>
__global__ void __launch_bounds__(256) wide_head_bwd_kernel(WideArgs w) {
    pdl_enter();
    constexpr int F = FeatN<FM>::value;
    const WideLayout &L = w.L;
    const int H = L.H, NSH = w.NS * H, tid = threadIdx.x;
    extern __shared__ __align__(16) float wsm[];
    float *sh = wsm;                    // [TT][64]   h tile, row-major
    float *sgo = sh + TT * WH;          // [TT][2]
    float *sWo = sgo + TT * 2;          // [2][VS]
    float *sbh = sWo + 2 * VS;          // [64]
    float *shT = sbh + WH;              // HEAD 1 from here:  [64][TP]
    float *sv = shT + WH * TP;          // [TT][VS]
    float *sdp = sv + TT * VS;          // [TT][64]   dL/d(fc_hid pre-activation)
    float *sdpT = sdp + TT * WH;        // [64][TP]
    float *sWhT = sdpT + WH * TP;       // [64][64]
    float *sWhr = sWhT + WH * WH;       // [64][64]
    head_stage<HEAD>(w, sWo, sWhT, sWhr, sbh, tid);
    const float gs = w.gscale * (w.gscale_dev ? __ldg(w.gscale_dev) : 1.0f);
    float accWo = 0.f;                  // HEAD 0: tid < 2H -> dWo[c][k], tid 2H+c -> dbo[c];  HEAD 1: tid < 2O -> dWo[c][m], tid 2O+c -> dbo[c]
    float accWh[16];                    // HEAD 1: dWh[jg*16+q][k]
    float accbh = 0.f;                  // HEAD 1: tid < 64 -> dbh[tid]
#pragma unroll
    for (int q = 0; q < 16; ++q) accWh[q] = 0.f;
    __syncthreads();
    for (int tile = blockIdx.x; tile < w.ntiles; tile += gridDim.x) {
        int b, t0, tv;
        tile_of(w, tile, b, t0, tv);
        const float *hrow = act_of(w, L.L - 1, b) + (size_t)t0 * NSH;
        for (int i = tid; i < TT * WH; i += 256) {
            const int t = i >> 6, k = i & 63;
            const float v = (t < tv && k < H) ? __ldg(hrow + (size_t)t * NSH + k) : 0.f;
            sh[i] = v;
            if constexpr (HEAD == 1) shT[k * TP + t] = v;
        }
        if (tid < TT) {
            float2 go = make_float2(0.f, 0.f);
            if (tid < tv) go = wide_go(w, b, t0 + tid, gs);
            sgo[2 * tid] = go.x; sgo[2 * tid + 1] = go.y;
            if constexpr (HEAD == 1) {
                float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (tid < tv) {
                    const float2 v = iq_row(w.x, w.x_bf16, w.x_starts, b, w.T).ld(t0 + tid);
                    features_fwd<FM>(v.x, v.y, 0.f, 0.f, f);
                }
#pragma unroll
                for (int m = 0; m < F; ++m) sv[tid * VS + H + m] = f[m];
            }
        }
        __syncthreads();
        float *dhrow = w.dh + ((size_t)b * w.T + t0) * H;
        if constexpr (HEAD == 0) {
            for (int i = tid; i < TT * H; i += 256) {
                const int t = i / H, k = i - t * H;
                if (t < tv) dhrow[(size_t)t * H + k] = fmaf(sgo[2 * t], sWo[k], sgo[2 * t + 1] * sWo[VS + k]);
            }
            if constexpr (DW) {
                if (tid < 2 * H) {
                    const int c = tid / H, k = tid - c * H;
#pragma unroll 8
                    for (int t = 0; t < TT; ++t) accWo = fmaf(sgo[2 * t + c], sh[t * WH + k], accWo);
                } else if (tid < 2 * H + 2) {
                    const int c = tid - 2 * H;
                    for (int t = 0; t < TT; ++t) accWo += sgo[2 * t + c];
                }
            }
        } else {
            {   // g = relu(fc_hid h), dL/dpre = (g > 0) * (go . W_o[:, j])
                float acc[8];
                head_fc_hid<HEAD>(sWhT, sbh, shT, H, tid, acc);
                const int j = tid & 63, tg = tid >> 6;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int t = tg * 8 + i;
                    const float gv = j < H ? fmaxf(acc[i], 0.f) : 0.f;
                    const float dg = fmaf(sgo[2 * t], sWo[j], sgo[2 * t + 1] * sWo[VS + j]);
                    const float dp = gv > 0.f ? dg : 0.f;
                    if (j < H) sv[t * VS + j] = gv;
                    sdp[t * WH + j] = dp;
                    sdpT[j * TP + t] = dp;
                }
            }
            __syncthreads();
            {   // dL/dh[t][k] = sum_j dpre[t][j] * Wh[j][k]
                const int k = tid & 63, tg = tid >> 6;
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                for (int j = 0; j < H; ++j) {
                    const float wv = sWhr[j * WH + k];
                    const float4 a = *reinterpret_cast<const float4 *>(sdpT + j * TP + tg * 8), c = *reinterpret_cast<const float4 *>(sdpT + j * TP + tg * 8 + 4);
                    acc[0] = fmaf(wv, a.x, acc[0]); acc[1] = fmaf(wv, a.y, acc[1]); acc[2] = fmaf(wv, a.z, acc[2]); acc[3] = fmaf(wv, a.w, acc[3]);
                    acc[4] = fmaf(wv, c.x, acc[4]); acc[5] = fmaf(wv, c.y, acc[5]); acc[6] = fmaf(wv, c.z, acc[6]); acc[7] = fmaf(wv, c.w, acc[7]);
                }
                if (k < H) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (tg * 8 + i < tv) dhrow[(size_t)(tg * 8 + i) * H + k] = acc[i];
                }
            }
            if constexpr (DW) {
                const int k = tid & 63, jg = tid >> 6;
                for (int t = 0; t < TT; ++t) {
                    const float hv = sh[t * WH + k];
                    const float4 *dp = reinterpret_cast<const float4 *>(sdp + t * WH + jg * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 d = dp[q];
                        accWh[4 * q] = fmaf(d.x, hv, accWh[4 * q]); accWh[4 * q + 1] = fmaf(d.y, hv, accWh[4 * q + 1]);
                        accWh[4 * q + 2] = fmaf(d.z, hv, accWh[4 * q + 2]); accWh[4 * q + 3] = fmaf(d.w, hv, accWh[4 * q + 3]);
                    }
                }
                if (tid < WH) {
                    for (int t = 0; t < TT; ++t) accbh += sdp[t * WH + tid];
                }
                if (tid < 2 * L.O) {
                    const int c = tid / L.O, m = tid - c * L.O;
                    for (int t = 0; t < TT; ++t) accWo = fmaf(sgo[2 * t + c], sv[t * VS + m], accWo);
                } else if (tid < 2 * L.O + 2) {
                    const int c = tid - 2 * L.O;
                    for (int t = 0; t < TT; ++t) accWo += sgo[2 * t + c];
                }
            }
        }
        __syncthreads();
    }
    if constexpr (DW) {
        float *prt = w.partials + (size_t)blockIdx.x * L.P;
        const int O = L.O;
        if (tid < 2 * O) prt[L.oWo + tid] = accWo;
        else if (tid < 2 * O + 2) prt[L.obo + tid - 2 * O] = accWo;
        if constexpr (HEAD == 1) {
            const int k = tid & 63, jg = tid >> 6;
            if (k < H) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int jj = jg * 16 + q;
                    if (jj < H) prt[L.oWh + jj * H + k] = accWh[q];
                }
            }
            if (tid < H) prt[L.obh + tid] = accbh;
        }
    }
}

// ================================================================ backward: reverse recurrence of one layer, one CTA per sequence
// Writes the gate-gradient row per step  GB[b][t][4H]:   GRU  ar | az | an | an*r      LSTM  ai | af | ag | ao
//   GRU   dn = dh (1-z), dz = dh (h_prev - n), an = dn (1-n^2), ar = an hgn r (1-r), az = dz z (1-z);
//         dh_prev = dh z + W_hr^T ar + W_hz^T az + W_hn^T (an r)
//   LSTM  do = dh tanh(c), dc += dh o (1 - tanh(c)^2), di = dc g, dg = dc i, df = dc c_prev, dc_prev = dc f;
//         dh_prev = W_hh^T [ai af ag ao]
template <int G>
__global__ void __launch_bounds__(G * 64) wide_chain_bwd_kernel(WideArgs w) {
    pdl_enter();
    const WideLayout &L = w.L;
    const int l = w.layer, H = L.H, T = w.T, NSH = w.NS * H;
    const int tid = threadIdx.x, g = tid >> 6, k = tid & 63, b = blockIdx.x;
    const bool act = k < H;
    __shared__ __align__(16) float sG[G][WH];
    __shared__ float spart[G][WH];
    float Wc[WH];
    {
        const float *Whh = w.params + L.oWhh[l] + (size_t)g * H * H + (act ? k : 0);
#pragma unroll
        for (int jj = 0; jj < WH; ++jj) Wc[jj] = (act && jj < H) ? __ldg(Whh + (size_t)jj * H) : 0.f;
    }
    sG[g][k] = 0.f;
    spart[g][k] = 0.f;
    __syncthreads();
    const float *arow = act_of(w, l, b);
    const float *dhrow = w.dh + (size_t)b * T * H;
    float *gbrow = w.gb + (size_t)b * T * 4 * H;
    constexpr int NV = G == 3 ? 6 : 7;
    const bool loader = g == 0 && act;
    auto load = [&](int t, float *d) {
#pragma unroll
        for (int q = 0; q < NV; ++q) d[q] = 0.f;
        if (loader && t >= 0) {
            const float *row = arow + (size_t)t * NSH;
            d[0] = __ldg(dhrow + (size_t)t * H + k);
            d[1] = __ldg(row + H + k); d[2] = __ldg(row + 2 * H + k); d[3] = __ldg(row + 3 * H + k); d[4] = __ldg(row + 4 * H + k);
            if constexpr (G == 3) {
                d[5] = t > 0 ? __ldg(row - NSH + k) : 0.f;                       // h_{t-1}
            } else {
                d[5] = __ldg(row + 5 * H + k);                                   // c_t
                d[6] = t > 0 ? __ldg(row - NSH + 5 * H + k) : 0.f;               // c_{t-1}
            }
        }
    };
    float rec = 0.f, dcr = 0.f;
    float cq[4][NV];
#pragma unroll
    for (int i = 0; i < 4; ++i) load(T - 1 - i, cq[i]);
    for (int t0 = T - 1; t0 >= 0; t0 -= 4) {
        float nq[4][NV];
#pragma unroll
        for (int i = 0; i < 4; ++i) load(t0 - 4 - i, nq[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = t0 - i;
            if (t >= 0) {
                float carry = 0.f;
                if (g == 0) {
                    const float *d = cq[i];
                    const float dh = d[0] + rec;
                    float g0, g1, g2, g3;
                    if constexpr (G == 3) {
                        const float r = d[1], z = d[2], n = d[3], hgn = d[4], hp = d[5];
                        const float dn = dh * (1.f - z), dz = dh * (hp - n);
                        const float an = dn * (1.f - n * n);
                        g0 = an * hgn * r * (1.f - r);     // ar
                        g1 = dz * z * (1.f - z);           // az
                        g2 = an;                           // an
                        g3 = an * r;                       // an*r
                        carry = dh * z;
                        sG[0][k] = g0; sG[1][k] = g1; sG[2][k] = g3;
                    } else {
                        const float ig = d[1], fg = d[2], gg = d[3], og = d[4], cc = d[5], cp = d[6];
                        const float tc = tanhf_(cc);
                        const float dc = fmaf(dh * og, 1.f - tc * tc, dcr);
                        g0 = dc * gg * ig * (1.f - ig);            // ai
                        g1 = dc * cp * fg * (1.f - fg);            // af
                        g2 = dc * ig * (1.f - gg * gg);            // ag
                        g3 = dh * tc * og * (1.f - og);            // ao
                        dcr = dc * fg;
                        sG[0][k] = g0; sG[1][k] = g1; sG[2][k] = g2; sG[3][k] = g3;
                    }
                    if (act) {
                        float *gr = gbrow + (size_t)t * 4 * H;
                        gr[k] = g0; gr[H + k] = g1; gr[2 * H + k] = g2; gr[3 * H + k] = g3;
                    }
                }
                __syncthreads();
                {
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                    const float4 *gv = reinterpret_cast<const float4 *>(sG[g]);
#pragma unroll
                    for (int q = 0; q < WH / 4; ++q) {
                        const float4 v = gv[q];
                        a0 = fmaf(Wc[4 * q], v.x, a0); a1 = fmaf(Wc[4 * q + 1], v.y, a1);
                        a2 = fmaf(Wc[4 * q + 2], v.z, a2); a3 = fmaf(Wc[4 * q + 3], v.w, a3);
                    }
                    spart[g][k] = (a0 + a1) + (a2 + a3);
                }
                __syncthreads();
                if (g == 0) {
                    float s = carry + spart[0][k] + spart[1][k] + spart[2][k];
                    if constexpr (G == 4) s += spart[3][k];
                    rec = s;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < NV; ++q) cq[i][q] = nq[i][q];
    }
}

// ================================================================ backward: weight gradients of one layer (gate g = blockIdx.y)
// dW_hh[gH+j][k] = sum_t Gh[t][j] h_{t-1}[k],  dW_ih[gH+j][m] = sum_t Gx[t][j] in_t[m],  db_hh = sum Gh,  db_ih = sum Gx;
// GRU gate n:  Gh = an*r (the r*(W_hn h + b_hn) term), Gx = an;  every other gate: Gh = Gx.
template <int G, int FM>
__global__ void __launch_bounds__(256) wide_wgrad_kernel(WideArgs w) {
    pdl_enter();
    constexpr int F = FeatN<FM>::value;
    const WideLayout &L = w.L;
    const int l = w.layer, H = L.H, T = w.T, NSH = w.NS * H, Fin = l == 0 ? F : H, tid = threadIdx.x;
    const int g = blockIdx.y, comp_x = g, comp_h = (G == 3 && g == 2) ? 3 : g;
    __shared__ __align__(16) float sGh[TT * WH];
    __shared__ __align__(16) float sGx[TT * WH];
    __shared__ __align__(16) float sV[TT * 2 * WH];       // [t][0..63] = h_{t-1},  [t][64..127] = input of the layer
    float ah[4][4], ax[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { ah[r][c] = 0.f; ax[r][c] = 0.f; }
    float bsh = 0.f, bsx = 0.f;
    const int rg = tid >> 4, cg = tid & 15;
    for (int tile = blockIdx.x; tile < w.ntiles; tile += gridDim.x) {
        int b, t0, tv;
        tile_of(w, tile, b, t0, tv);
        const float *gbrow = w.gb + ((size_t)b * T + t0) * 4 * H;
        const float *arow = act_of(w, l, b);
        const float *irow = l > 0 ? act_of(w, l - 1, b) : nullptr;
        for (int i = tid; i < TT * WH; i += 256) {
            const int t = i >> 6, jj = i & 63;
            const bool ok = t < tv && jj < H;
            sGh[i] = ok ? __ldg(gbrow + (size_t)t * 4 * H + comp_h * H + jj) : 0.f;
            sGx[i] = ok ? __ldg(gbrow + (size_t)t * 4 * H + comp_x * H + jj) : 0.f;
            const int tp = t0 + t - 1;
            sV[t * 2 * WH + jj] = (ok && tp >= 0) ? __ldg(arow + (size_t)tp * NSH + jj) : 0.f;
            sV[t * 2 * WH + WH + jj] = (ok && irow) ? __ldg(irow + (size_t)(t0 + t) * NSH + jj) : 0.f;
        }
        if (l == 0) {
            __syncthreads();
            if (tid < tv) {
                float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                const float2 v = iq_row(w.x, w.x_bf16, w.x_starts, b, T).ld(t0 + tid);
                features_fwd<FM>(v.x, v.y, 0.f, 0.f, f);
#pragma unroll
                for (int m = 0; m < F; ++m) sV[tid * 2 * WH + WH + m] = f[m];
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < TT; ++t) {
            const float4 gh = *reinterpret_cast<const float4 *>(sGh + t * WH + rg * 4);
            const float4 gx = *reinterpret_cast<const float4 *>(sGx + t * WH + rg * 4);
            const float4 vh = *reinterpret_cast<const float4 *>(sV + t * 2 * WH + cg * 4);
            const float4 vx = *reinterpret_cast<const float4 *>(sV + t * 2 * WH + WH + cg * 4);
            const float a[4] = {gh.x, gh.y, gh.z, gh.w}, c[4] = {gx.x, gx.y, gx.z, gx.w};
            const float p[4] = {vh.x, vh.y, vh.z, vh.w}, q[4] = {vx.x, vx.y, vx.z, vx.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int s = 0; s < 4; ++s) { ah[r][s] = fmaf(a[r], p[s], ah[r][s]); ax[r][s] = fmaf(c[r], q[s], ax[r][s]); }
        }
        if (tid < WH) {
            for (int t = 0; t < TT; ++t) { bsh += sGh[t * WH + tid]; bsx += sGx[t * WH + tid]; }
        }
        __syncthreads();
    }
    float *prt = w.partials + (size_t)blockIdx.x * L.P;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int jj = rg * 4 + r;
        if (jj < H) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int col = cg * 4 + s;
                if (col < H) prt[L.oWhh[l] + (size_t)(g * H + jj) * H + col] = ah[r][s];
                if (col < Fin) prt[L.oWih[l] + (size_t)(g * H + jj) * Fin + col] = ax[r][s];
            }
        }
    }
    if (tid < H) { prt[L.obih[l] + g * H + tid] = bsx; prt[L.obhh[l] + g * H + tid] = bsh; }
}

// ================================================================ backward: dL/d(input of the layer) = Gx W_ih
// layer > 0: becomes dL/dh of layer l-1 (overwrites DH);  layer 0: through the feature Jacobian to dL/dx (+ the DGRU head's direct
// feature path fc_out[:, H:]).
template <int FM, int HEAD>
__global__ void __launch_bounds__(256) wide_dx_kernel(WideArgs w) {
    pdl_enter();
    constexpr int F = FeatN<FM>::value;
    const WideLayout &L = w.L;
    const int l = w.layer, H = L.H, T = w.T, GH = L.G * H, Fin = l == 0 ? F : H, tid = threadIdx.x;
    extern __shared__ __align__(16) float wsm[];
    float *sW = wsm;                     // [GH][Fin]  W_ih row-major
    float *sGT = sW + al4(GH * Fin);     // [GH][TP]   Gx of the tile, transposed
    float *sdf = sGT + GH * TP;          // [TT][8]    dL/dfeatures (layer 0)
    for (int i = tid; i < GH * Fin; i += 256) sW[i] = __ldg(w.params + L.oWih[l] + i);
    const float gs = w.gscale * (w.gscale_dev ? __ldg(w.gscale_dev) : 1.0f);
    __syncthreads();
    for (int tile = blockIdx.x; tile < w.ntiles; tile += gridDim.x) {
        int b, t0, tv;
        tile_of(w, tile, b, t0, tv);
        const float *gbrow = w.gb + ((size_t)b * T + t0) * 4 * H;
        for (int i = tid; i < TT * GH; i += 256) {
            const int t = i / GH, row = i - t * GH;
            sGT[row * TP + t] = t < tv ? __ldg(gbrow + (size_t)t * 4 * H + row) : 0.f;
        }
        __syncthreads();
        if (l > 0) {
            const int m = tid & 63, tg = tid >> 6;
            if (m < H) {
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                for (int row = 0; row < GH; ++row) {
                    const float wv = sW[row * Fin + m];
                    const float4 a = *reinterpret_cast<const float4 *>(sGT + row * TP + tg * 8), c = *reinterpret_cast<const float4 *>(sGT + row * TP + tg * 8 + 4);
                    acc[0] = fmaf(wv, a.x, acc[0]); acc[1] = fmaf(wv, a.y, acc[1]); acc[2] = fmaf(wv, a.z, acc[2]); acc[3] = fmaf(wv, a.w, acc[3]);
                    acc[4] = fmaf(wv, c.x, acc[4]); acc[5] = fmaf(wv, c.y, acc[5]); acc[6] = fmaf(wv, c.z, acc[6]); acc[7] = fmaf(wv, c.w, acc[7]);
                }
                float *dhrow = w.dh + ((size_t)b * T + t0) * H;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (tg * 8 + i < tv) dhrow[(size_t)(tg * 8 + i) * H + m] = acc[i];
            }
        } else {
            if (tid < TT * F) {
                const int m = tid / TT, t = tid - m * TT;
                float acc = 0.f;
                for (int row = 0; row < GH; ++row) acc = fmaf(sW[row * F + m], sGT[row * TP + t], acc);
                sdf[t * 8 + m] = acc;
            }
            __syncthreads();
            if (tid < tv) {
                float gf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int m = 0; m < F; ++m) gf[m] = sdf[tid * 8 + m];
                if constexpr (HEAD == 1) {
                    const float2 go = wide_go(w, b, t0 + tid, gs);
#pragma unroll
                    for (int m = 0; m < F; ++m)
                        gf[m] += fmaf(go.x, __ldg(w.params + L.oWo + H + m), go.y * __ldg(w.params + L.oWo + L.O + H + m));
                }
                const float2 v = iq_row(w.x, w.x_bf16, w.x_starts, b, T).ld(t0 + tid);
                float gi, gq;
                features_bwd<FM>(v.x, v.y, gf, gi, gq);
                reinterpret_cast<float2 *>(w.gx)[(size_t)b * T + t0 + tid] = make_float2(gi, gq);
            }
        }
        __syncthreads();
    }
}

// ================================================================ host
static int wide_grid(int B, int T) {
    const int64_t tiles = (int64_t)B * ((T + TT - 1) / TT);
    const int64_t cap = 2 * (int64_t)num_sms();
    return (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
}
static int act_ns(int G, bool save) { return save ? (G == 3 ? 5 : 6) : 1; }

bool wide_supported(int cell) { int G, FM, head; return wide_cell_info(cell, G, FM, head); }
int64_t wide_nparams(int cell, int H, int layers) { return make_layout(cell, H, layers).P; }
// `saved` = per layer [B][T][NS*H] activation rows | XP [B][T][G*H] (input projection of the layer in flight)
int64_t wide_saved_floats(int cell, int B, int T, int H, int layers, bool save) {
    int G, FM, head;
    if (!wide_cell_info(cell, G, FM, head)) return -1;
    const int64_t bt = (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1);
    return bt * ((int64_t)layers * act_ns(G, save) * H + (int64_t)G * H) + 4;
}
// workspace = gradient partials [rows][P] (4-aligned) | DH [B][T][H] | GB [B][T][4H]
int64_t wide_workspace_floats(int cell, int B, int T, int H, int layers) {
    const WideLayout L = make_layout(cell, H, layers);
    const int64_t bt = (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1);
    return (((int64_t)wide_grid(B, T) * L.P + 3) & ~(int64_t)3) + bt * 5 * H + 4;
}

// raise a kernel's dynamic shared-memory limit when a call needs more than any before it on this device (the attribute is per
// device and per kernel; different instantiations share one function-pointer type, hence the pointer-keyed table)
static void ensure_smem(const void *k, size_t bytes) {
    struct Entry { const void *k; int dev; size_t bytes; };
    static Entry tab[64];
    static int n = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    const int dev = cur_dev_slot();
    for (int i = 0; i < n; ++i)
        if (tab[i].k == k && tab[i].dev == dev) {
            if (bytes > tab[i].bytes) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); tab[i].bytes = bytes; }
            return;
        }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (n < 64) tab[n++] = Entry{k, dev, bytes};
}

template <int FM, int HEAD>
static int wide_run_t(WideArgs w, int dir, bool dw, cudaStream_t st, int *rows_out) {
    const WideLayout &L = w.L;
    const int H = L.H, GH = L.G * H, grid = wide_grid(w.B, w.T);
    auto xproj_smem = [&](int l) { const int Fin = l == 0 ? L.F : H; return (size_t)(al4(Fin * GH) + al4(GH) + Fin * TP) * sizeof(float); };
    const size_t head_f_smem = (size_t)(WH * TP + TT * VS + 2 * VS + WH + (HEAD ? WH * WH : 0)) * sizeof(float);
    const size_t head_b_smem = (size_t)(TT * WH + TT * 2 + 2 * VS + WH + (HEAD ? WH * TP + TT * VS + TT * WH + WH * TP + 2 * WH * WH : 0)) * sizeof(float);
    auto dx_smem = [&](int l) { const int Fin = l == 0 ? L.F : H; return (size_t)(al4(GH * Fin) + GH * TP + TT * 8) * sizeof(float); };
    if (dir == 0) {
        for (int l = 0; l < L.L; ++l) {
            w.layer = l;
            ensure_smem((const void *)wide_xproj_kernel<FM>, xproj_smem(l));
            launch_pdl(wide_xproj_kernel<FM>, dim3(grid), dim3(256), xproj_smem(l), st, w);
            if (L.G == 3) launch_pdl(wide_chain_fwd_kernel<3>, dim3(w.B), dim3(192), 0, st, w);
            else launch_pdl(wide_chain_fwd_kernel<4>, dim3(w.B), dim3(256), 0, st, w);
        }
        ensure_smem((const void *)wide_head_fwd_kernel<FM, HEAD>, head_f_smem);
        launch_pdl(wide_head_fwd_kernel<FM, HEAD>, dim3(grid), dim3(256), head_f_smem, st, w);
        return check_launch("wide forward");
    }
    if (dw) {
        ensure_smem((const void *)wide_head_bwd_kernel<FM, HEAD, true>, head_b_smem);
        launch_pdl(wide_head_bwd_kernel<FM, HEAD, true>, dim3(grid), dim3(256), head_b_smem, st, w);
    } else {
        ensure_smem((const void *)wide_head_bwd_kernel<FM, HEAD, false>, head_b_smem);
        launch_pdl(wide_head_bwd_kernel<FM, HEAD, false>, dim3(grid), dim3(256), head_b_smem, st, w);
    }
    for (int l = L.L - 1; l >= 0; --l) {
        w.layer = l;
        if (L.G == 3) launch_pdl(wide_chain_bwd_kernel<3>, dim3(w.B), dim3(192), 0, st, w);
        else launch_pdl(wide_chain_bwd_kernel<4>, dim3(w.B), dim3(256), 0, st, w);
        if (dw) {
            if (L.G == 3) launch_pdl(wide_wgrad_kernel<3, FM>, dim3(grid, 3), dim3(256), 0, st, w);
            else launch_pdl(wide_wgrad_kernel<4, FM>, dim3(grid, 4), dim3(256), 0, st, w);
        }
        if (l > 0 || w.need_dx) {
            ensure_smem((const void *)wide_dx_kernel<FM, HEAD>, dx_smem(l));
            launch_pdl(wide_dx_kernel<FM, HEAD>, dim3(grid), dim3(256), dx_smem(l), st, w);
        }
    }
    if (rows_out) *rows_out = grid;
    return check_launch("wide backward");
}

// dir 0 = forward, 1 = backward.  a.K carries nothing here; `layers` comes from OdpdDims.K (api.cu).
int wide_run(int cell, const GruArgs &a, int layers, int dir, bool dw, cudaStream_t st, int *rows_out) {
    int G, FM, head;
    if (!wide_cell_info(cell, G, FM, head)) { set_error("wide_run: cell %d has no layered path", cell); return -1; }
    if (a.H < 1 || a.H > WH || layers < 1 || layers > WLMAX) {
        set_error("layered RNN path: hidden_size %d / num_layers %d outside 1..%d / 1..%d", a.H, layers, WH, WLMAX);
        return -1;
    }
    WideArgs w{};
    w.L = make_layout(cell, a.H, layers);
    w.B = a.B; w.T = a.T; w.nts = (a.T + TT - 1) / TT; w.ntiles = a.B * w.nts;
    w.save = dir == 0 ? a.save : 1;
    w.NS = act_ns(G, w.save != 0);
    w.need_dx = a.need_dx;
    w.params = a.params; w.x = a.x; w.target = a.target; w.x_bf16 = a.x_bf16; w.target_bf16 = a.target_bf16;
    w.x_starts = a.x_starts; w.target_starts = a.target_starts;
    w.gout = a.gout; w.out_in = a.out_in; w.gscale_dev = a.gscale_dev; w.gscale = a.gscale; w.loss_scale = a.loss_scale;
    w.out = a.out; w.gx = a.gx; w.loss = a.loss;
    const int64_t bt = (int64_t)a.B * a.T;
    w.act = a.saved;
    w.xp = a.saved ? a.saved + bt * layers * w.NS * a.H : nullptr;
    if (dir == 1) {
        const int grid = wide_grid(a.B, a.T);
        const int64_t poff = ((int64_t)grid * w.L.P + 3) & ~(int64_t)3;
        w.partials = a.partials;
        w.dh = a.partials ? a.partials + poff : nullptr;
        w.gb = a.partials ? a.partials + poff + bt * a.H : nullptr;
        if (!a.partials) { set_error("layered RNN path: the backward needs the workspace (also for a dX-only call)"); return -1; }
    } else if (!a.saved) {
        set_error("layered RNN path: the forward needs the `saved` buffer (odpd_saved_bytes), also without ODPD_F_SAVE");
        return -1;
    }
#define ODPD_WIDE_CASE(FMV, HEADV) return wide_run_t<FMV, HEADV>(w, dir, dw, st, rows_out)
    switch (cell) {
    case ODPD_CELL_GRU: case ODPD_CELL_LSTM: ODPD_WIDE_CASE(FM_RAW2, 0);
    case ODPD_CELL_DGRU: ODPD_WIDE_CASE(FM_DGRU6, 1);
    case ODPD_CELL_QGRU: ODPD_WIDE_CASE(FM_QGRU4, 0);
    case ODPD_CELL_QGRU_AMP1: ODPD_WIDE_CASE(FM_AMP4, 0);
    }
#undef ODPD_WIDE_CASE
    return -1;
}

}  // namespace odpd

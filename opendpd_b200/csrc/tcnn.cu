// tcnn.cu — TCNN and NeuralTX backbones: the dilated depthwise temporal-convolution stack, forward / backward.
//
// Replaces (reference, file:line):
//   backbones/tcnn.py:5-32,83-97   features (I,Q,|x|,|x|^3,sin,cos) (:85-93) -> Conv1d(6->C,k=1,bias) hardswish -> 4 x [depthwise Conv1d(C,C,k=5,
//                                  dilation d = 1,2,4,8, padding 2d, no bias) hardswish] -> Conv1d(C->2,k=1,no bias);  out = stack(x) + (I,Q) (:97)
//   backbones/neuraltx.py:5-38,107-124   (the fft over a length-1 axis, :108, is the identity)  5-tap complex FIR with two real kernels, zero 'same'
//                                  padding (:113-114: I_f = conv_I(I) - conv_Q(Q), Q_f = conv_Q(I) + conv_I(Q)); features (I_f,Q_f,|.|,|.|^3) (:115-119);
//                                  the same stack with 4 inputs; out = stack + IQ_match(I_f,Q_f) + (I_f,Q_f) (:123)
// C = hidden_channels = the CLI's hidden size (models.py:130-135), 1..64.  Every convolution zero-pads per frame.
//
// No recurrence: one CTA per 64-step tile.  Forward: the tile plus a 30-step halo on both sides (the stack's receptive field: 2(1+2+4+8))
// is pushed through the five layers in shared memory ([C][124] ping-pong); with ODPD_F_SAVE the five pre-activation maps of the tile core
// go to HBM ([B][5][C][T], time-contiguous).  Backward: the gradient of the tile plus halo walks the layers in reverse in shared memory,
// hardswish' and the layer inputs come from the saved maps; weight gradients are sums over the tile core, one parameter per thread at a
// time, accumulated in the CTA's gradient-partial row (reduced in order by reduce_partials_kernel).  dL/dx needs no scatter: the halo
// makes every core step's input gradient complete inside its own tile (NeuralTX's FIR adds a second, 2-step gather pass).
//
// Flat parameter layouts (named_parameters() order):
//   TCNN      network.0.weight(C,6,1) network.0.bias(C) network.{2,4,6,8}.weight(C,1,5) network.10.weight(2,C,1)                      = 29 C
//   NeuralTX  conv_I.weight(1,1,5) conv_Q.weight(1,1,5) network.0.weight(C,4,1) network.0.bias(C) network.{2,4,6,8}.weight(C,1,5)
//             network.10.weight(2,C,1) IQ_match.weight(2,2)                                                                             = 27 C + 14
#include <mutex>
#include "cells.h"
#include "chunking.cuh"

namespace odpd {

static constexpr int TC_TT = 64;                 // tile core (timesteps)
static constexpr int TC_HALO = 30;               // 2 * (1 + 2 + 4 + 8)
static constexpr int TC_W = TC_TT + 2 * TC_HALO; // 124 window positions
static constexpr int TC_WP = TC_W + 1;           // row pitch
static constexpr int TC_CMAX = 64;
static constexpr int TC_NT = 128;                // threads per CTA

template <bool NTX>
struct TcLayout {
    static constexpr int F = NTX ? 4 : 6;
    int C, oFI, oFQ, oW0, ob0, oDw[4], oW10, oIQ, P;
    __host__ __device__ explicit TcLayout(int c) {
        C = c;
        int off = 0;
        oFI = 0; oFQ = 5;
        if (NTX) off = 10;
        oW0 = off; off += c * F;
        ob0 = off; off += c;
        for (int l = 0; l < 4; ++l) { oDw[l] = off; off += 5 * c; }
        oW10 = off; off += 2 * c;
        oIQ = off;
        if (NTX) off += 4;
        P = off;
    }
};

__device__ __forceinline__ float hsw(float v) { return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f); }
// d hardswish / dv:  0 for v < -3, 1 for v > 3, (2v+3)/6 between (PyTorch's hardswish_backward)
__device__ __forceinline__ float hsw_grad(float v) { return v < -3.f ? 0.f : (v > 3.f ? 1.f : fmaf(v, 1.f / 3.f, 0.5f)); }

// input of the stack at frame position s (0 <= s < T): TCNN features of sample s; NeuralTX: FIR over samples s-2..s+2, then features
template <bool NTX>
__device__ __forceinline__ void tc_input(const IqRow &x2, const float *sFir, int s, int T, float *f, float &fi, float &fq) {
    if constexpr (!NTX) {
        const float2 v = x2.ld(s);
        features_fwd<FM_DGRU6>(v.x, v.y, 0.f, 0.f, f);
        fi = v.x; fq = v.y;
    } else {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int q = s + k - 2;
            if (q < 0 || q >= T) continue;
            const float2 v = x2.ld(q);
            a = fmaf(sFir[k], v.x, fmaf(-sFir[5 + k], v.y, a));
            b = fmaf(sFir[5 + k], v.x, fmaf(sFir[k], v.y, b));
        }
        const float a2 = fmaf(a, a, b * b), am = sqrtf(a2);
        f[0] = a; f[1] = b; f[2] = am; f[3] = am * am * am;
        fi = a; fq = b;
    }
}

// ================================================================ forward
template <bool NTX>
__global__ void __launch_bounds__(TC_NT) tcnn_fwd_kernel(GruArgs a, int nts, int ntiles) {
    pdl_enter();
    using LT = TcLayout<NTX>;
    constexpr int F = LT::F;
    const LT L(a.H);
    const int C = a.H, T = a.T, tid = threadIdx.x;
    extern __shared__ __align__(16) float tsm[];
    float *sp = tsm;                         // parameters (flat)
    float *bufA = sp + ((L.P + 3) & ~3);     // [C][TC_WP]
    float *bufB = bufA + C * TC_WP;          // [C][TC_WP]
    float *sfe = bufB + C * TC_WP;           // [TC_W][8]  stack inputs (features) | [6],[7] = residual (I,Q) or (I_f,Q_f)
    __shared__ float sred[TC_NT / 32];
    for (int i = tid; i < L.P; i += TC_NT) sp[i] = __ldg(a.params + i);
    __syncthreads();
    float lsum = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / nts, t0 = (tile - b * nts) * TC_TT;
        const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
        // stack inputs over the window
        for (int p = tid; p < TC_W; p += TC_NT) {
            const int s = t0 - TC_HALO + p;
            float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            float fi = 0.f, fq = 0.f;
            if (s >= 0 && s < T) tc_input<NTX>(x2, sp, s, T, f, fi, fq);
#pragma unroll
            for (int m = 0; m < F; ++m) sfe[p * 8 + m] = f[m];
            sfe[p * 8 + 6] = fi; sfe[p * 8 + 7] = fq;
        }
        __syncthreads();
        // layer 0: 1x1 conv + bias  (positions outside the frame stay exactly zero: they are the next layer's padding)
        float *cur = bufA, *nxt = bufB;
        for (int i = tid; i < C * TC_W; i += TC_NT) {
            const int c = i / TC_W, p = i - c * TC_W, s = t0 - TC_HALO + p;
            float pre = 0.f, v = 0.f;
            if (s >= 0 && s < T) {
                pre = sp[L.ob0 + c];
#pragma unroll
                for (int m = 0; m < F; ++m) pre = fmaf(sp[L.oW0 + c * F + m], sfe[p * 8 + m], pre);
                v = hsw(pre);
                if (a.save && p >= TC_HALO && p < TC_HALO + TC_TT) a.saved[(((size_t)b * 5 + 0) * C + c) * T + s] = pre;
            }
            cur[c * TC_WP + p] = v;
        }
        __syncthreads();
        // layers 1..4: depthwise k=5, dilation d
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
            const int d = 1 << l;
            for (int i = tid; i < C * TC_W; i += TC_NT) {
                const int c = i / TC_W, p = i - c * TC_W, s = t0 - TC_HALO + p;
                float pre = 0.f, v = 0.f;
                if (s >= 0 && s < T) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const int q = p + (k - 2) * d;
                        if (q >= 0 && q < TC_W) pre = fmaf(sp[L.oDw[l] + c * 5 + k], cur[c * TC_WP + q], pre);
                    }
                    v = hsw(pre);
                    if (a.save && p >= TC_HALO && p < TC_HALO + TC_TT) a.saved[(((size_t)b * 5 + l + 1) * C + c) * T + s] = pre;
                }
                nxt[c * TC_WP + p] = v;
            }
            __syncthreads();
            float *tmp = cur; cur = nxt; nxt = tmp;
        }
        // output 1x1 conv + residual (+ IQ_match)
        if (tid < TC_TT) {
            const int p = TC_HALO + tid, s = t0 + tid;
            if (s < T) {
                float o0 = 0.f, o1 = 0.f;
                for (int c = 0; c < C; ++c) {
                    const float v = cur[c * TC_WP + p];
                    o0 = fmaf(sp[L.oW10 + c], v, o0);
                    o1 = fmaf(sp[L.oW10 + C + c], v, o1);
                }
                const float ri = sfe[p * 8 + 6], rq = sfe[p * 8 + 7];
                if constexpr (NTX) {
                    o0 += fmaf(sp[L.oIQ], ri, sp[L.oIQ + 1] * rq) + ri;
                    o1 += fmaf(sp[L.oIQ + 2], ri, sp[L.oIQ + 3] * rq) + rq;
                } else {
                    o0 += ri; o1 += rq;
                }
                reinterpret_cast<float2 *>(a.out)[(size_t)b * T + s] = make_float2(o0, o1);
                if (a.target) {
                    const float2 y = iq_row(a.target, a.target_bf16, a.target_starts, b, T).ld(s);
                    const float d0 = o0 - y.x, d1 = o1 - y.y;
                    lsum = fmaf(d0, d0, fmaf(d1, d1, lsum));
                }
            }
        }
        __syncthreads();
    }
    if (a.loss && a.target) {
        lsum = warp_sum(lsum);
        if ((tid & 31) == 0) sred[tid >> 5] = lsum;
        __syncthreads();
        if (tid == 0) atomicAdd(a.loss, (double)(sred[0] + sred[1] + sred[2] + sred[3]) * (double)a.loss_scale);
    }
}

// ================================================================ backward
// saved pre-activation of layer l, channel c at frame position s (0 outside the frame -> hardswish(0) = 0: the zero padding)
__device__ __forceinline__ float tc_pre(const float *saved, int b, int l, int C, int c, int T, int s) {
    return (s >= 0 && s < T) ? __ldg(saved + (((size_t)b * 5 + l) * C + c) * T + s) : 0.f;
}

template <bool NTX, bool DW>
__global__ void __launch_bounds__(TC_NT) tcnn_bwd_kernel(GruArgs a, int nts, int ntiles, float2 *dfir) {
    pdl_enter();
    using LT = TcLayout<NTX>;
    constexpr int F = LT::F;
    const LT L(a.H);
    const int C = a.H, T = a.T, tid = threadIdx.x;
    extern __shared__ __align__(16) float tsm[];
    float *sp = tsm;
    float *bufA = sp + ((L.P + 3) & ~3);     // [C][TC_WP]  gradient w.r.t. a layer's pre-activation
    float *bufB = bufA + C * TC_WP;          // [C][TC_WP]
    float *sgo = bufB + C * TC_WP;           // [TC_W][2]   dL/dout
    float *sfe = sgo + TC_W * 2;             // [TC_TT][8]  stack inputs of the core (+ residual values)
    for (int i = tid; i < L.P; i += TC_NT) sp[i] = __ldg(a.params + i);
    __syncthreads();
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
    float *prt = (DW && a.partials) ? a.partials + (size_t)blockIdx.x * L.P : nullptr;
    bool first = true;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / nts, t0 = (tile - b * nts) * TC_TT;
        const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
        for (int p = tid; p < TC_W; p += TC_NT) {
            const int s = t0 - TC_HALO + p;
            float2 go = make_float2(0.f, 0.f);
            if (s >= 0 && s < T) {
                if (a.gout) go = __ldg(reinterpret_cast<const float2 *>(a.gout) + (size_t)b * T + s);
                else {
                    const float2 o = __ldg(reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * T + s);
                    const float2 y = iq_row(a.target, a.target_bf16, a.target_starts, b, T).ld(s);
                    go = make_float2(gs * (o.x - y.x), gs * (o.y - y.y));
                }
            }
            sgo[2 * p] = go.x; sgo[2 * p + 1] = go.y;
        }
        if (tid < TC_TT) {
            const int s = t0 + tid;
            float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            float fi = 0.f, fq = 0.f;
            if (s < T) tc_input<NTX>(x2, sp, s, T, f, fi, fq);
#pragma unroll
            for (int m = 0; m < F; ++m) sfe[tid * 8 + m] = f[m];
            sfe[tid * 8 + 6] = fi; sfe[tid * 8 + 7] = fq;
        }
        __syncthreads();
        // gradient w.r.t. the pre-activation of layer 4 over the window
        float *cur = bufA, *nxt = bufB;
        for (int i = tid; i < C * TC_W; i += TC_NT) {
            const int c = i / TC_W, p = i - c * TC_W, s = t0 - TC_HALO + p;
            float g = 0.f;
            if (s >= 0 && s < T) g = fmaf(sgo[2 * p], sp[L.oW10 + c], sgo[2 * p + 1] * sp[L.oW10 + C + c]) * hsw_grad(tc_pre(a.saved, b, 4, C, c, T, s));
            cur[c * TC_WP + p] = g;
        }
        __syncthreads();
        if constexpr (DW) {
            if (prt) {
                // network.10.weight[o][c] = sum_core dout[o] * act4[c];  IQ_match.weight[o][m] = sum_core dout[o] * (I_f,Q_f)[m]
                for (int o = tid; o < 2 * C + (NTX ? 4 : 0); o += TC_NT) {
                    float sacc = 0.f;
                    if (o < 2 * C) {
                        const int oo = o / C, c = o - oo * C;
                        for (int tt = 0; tt < TC_TT; ++tt) {
                            const int s = t0 + tt;
                            if (s < T) sacc = fmaf(sgo[2 * (TC_HALO + tt) + oo], hsw(tc_pre(a.saved, b, 4, C, c, T, s)), sacc);
                        }
                        const int idx = L.oW10 + o;
                        prt[idx] = first ? sacc : prt[idx] + sacc;
                    } else {
                        const int q = o - 2 * C, oo = q >> 1, m = q & 1;
                        for (int tt = 0; tt < TC_TT; ++tt) sacc = fmaf(sgo[2 * (TC_HALO + tt) + oo], sfe[tt * 8 + 6 + m], sacc);
                        const int idx = L.oIQ + q;
                        prt[idx] = first ? sacc : prt[idx] + sacc;
                    }
                }
            }
        }
        // layers 4..1: weight gradient of the depthwise kernel, then the transposed convolution and hardswish' of the layer below
#pragma unroll 1
        for (int l = 3; l >= 0; --l) {
            const int d = 1 << l;
            if constexpr (DW) {
                if (prt) {
                    for (int o = tid; o < 5 * C; o += TC_NT) {
                        const int c = o / 5, k = o - c * 5;
                        float sacc = 0.f;
                        for (int tt = 0; tt < TC_TT; ++tt) {
                            const int s = t0 + tt;
                            if (s < T) sacc = fmaf(cur[c * TC_WP + TC_HALO + tt], hsw(tc_pre(a.saved, b, l, C, c, T, s + (k - 2) * d)), sacc);
                        }
                        const int idx = L.oDw[l] + o;
                        prt[idx] = first ? sacc : prt[idx] + sacc;
                    }
                }
            }
            for (int i = tid; i < C * TC_W; i += TC_NT) {
                const int c = i / TC_W, p = i - c * TC_W, s = t0 - TC_HALO + p;
                float g = 0.f;
                if (s >= 0 && s < T) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const int q = p - (k - 2) * d;          // y[q] = sum_k w[k] x[q + (k-2) d]  ->  dx[p] += w[k] dy[p - (k-2) d]
                        if (q >= 0 && q < TC_W) g = fmaf(sp[L.oDw[l] + c * 5 + k], cur[c * TC_WP + q], g);
                    }
                    g *= hsw_grad(tc_pre(a.saved, b, l, C, c, T, s));
                }
                nxt[c * TC_WP + p] = g;
            }
            __syncthreads();
            float *tmp = cur; cur = nxt; nxt = tmp;
        }
        // cur = dL/dpre0.  network.0.weight / bias
        if constexpr (DW) {
            if (prt) {
                for (int o = tid; o < C * (F + 1); o += TC_NT) {
                    float sacc = 0.f;
                    int idx;
                    if (o < C * F) {
                        const int c = o / F, m = o - c * F;
                        for (int tt = 0; tt < TC_TT; ++tt) sacc = fmaf(cur[c * TC_WP + TC_HALO + tt], sfe[tt * 8 + m], sacc);
                        idx = L.oW0 + o;
                    } else {
                        const int c = o - C * F;
                        for (int tt = 0; tt < TC_TT; ++tt) sacc += cur[c * TC_WP + TC_HALO + tt];
                        idx = L.ob0 + c;
                    }
                    prt[idx] = first ? sacc : prt[idx] + sacc;
                }
            }
        }
        // input gradient of the core steps
        if (tid < TC_TT && (a.need_dx || (NTX && DW))) {
            const int p = TC_HALO + tid, s = t0 + tid;
            if (s < T) {
                float gf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                for (int c = 0; c < C; ++c) {
                    const float g = cur[c * TC_WP + p];
#pragma unroll
                    for (int m = 0; m < F; ++m) gf[m] = fmaf(g, sp[L.oW0 + c * F + m], gf[m]);
                }
                const float go0 = sgo[2 * p], go1 = sgo[2 * p + 1];
                if constexpr (!NTX) {
                    const float2 v = x2.ld(s);
                    float gi, gq;
                    features_bwd<FM_DGRU6>(v.x, v.y, gf, gi, gq);
                    reinterpret_cast<float2 *>(a.gx)[(size_t)b * T + s] = make_float2(gi + go0, gq + go1);
                } else {
                    // features (I_f, Q_f, |.|, |.|^3) + residual + IQ_match  ->  dL/d(I_f, Q_f); the FIR is transposed by tcnn_fir_bwd_kernel
                    const float fi = sfe[tid * 8 + 6], fq = sfe[tid * 8 + 7], am = sfe[tid * 8 + 2];
                    const float ga = fmaf(3.f * am * am, gf[3], gf[2]);
                    float di = gf[0] + go0 + fmaf(go0, sp[L.oIQ], go1 * sp[L.oIQ + 2]);
                    float dq = gf[1] + go1 + fmaf(go0, sp[L.oIQ + 1], go1 * sp[L.oIQ + 3]);
                    di = fmaf(ga, fi / am, di);
                    dq = fmaf(ga, fq / am, dq);
                    dfir[(size_t)b * T + s] = make_float2(di, dq);
                }
            }
        }
        first = false;
        __syncthreads();
    }
    if constexpr (DW) {
        if (prt && first)
            for (int o = tid; o < L.P; o += TC_NT) prt[o] = 0.f;
    }
}

// NeuralTX: transposed 5-tap complex FIR.  dL/dx[s] gathers dL/d(I_f,Q_f)[t], t = s-2..s+2;  conv_I / conv_Q kernel gradients.
template <bool DW>
__global__ void __launch_bounds__(TC_NT) tcnn_fir_bwd_kernel(GruArgs a, int nts2, int ntiles2, const float2 *__restrict__ dfir, int first_rows) {
    pdl_enter();
    const int T = a.T, tid = threadIdx.x;
    __shared__ float sF[10];
    __shared__ float sred[TC_NT / 32][10];
    if (tid < 10) sF[tid] = __ldg(a.params + tid);
    __syncthreads();
    float gw[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) gw[k] = 0.f;
    for (int tile = blockIdx.x; tile < ntiles2; tile += gridDim.x) {
        const int b = tile / nts2, s = (tile - b * nts2) * TC_NT + tid;
        if (s >= T) continue;
        const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
        const float2 *df = dfir + (size_t)b * T;
        // I_f[t] = sum_k wI[k] I[t+k-2] - wQ[k] Q[t+k-2],  Q_f[t] = sum_k wQ[k] I[t+k-2] + wI[k] Q[t+k-2]
        if (a.need_dx) {
            float gi = 0.f, gq = 0.f;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int t = s - (k - 2);
                if (t < 0 || t >= T) continue;
                const float2 d = df[t];
                gi = fmaf(d.x, sF[k], fmaf(d.y, sF[5 + k], gi));
                gq = fmaf(d.y, sF[k], fmaf(-d.x, sF[5 + k], gq));
            }
            reinterpret_cast<float2 *>(a.gx)[(size_t)b * T + s] = make_float2(gi, gq);
        }
        if constexpr (DW) {
            const float2 d = df[s];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int q = s + k - 2;
                if (q < 0 || q >= T) continue;
                const float2 v = x2.ld(q);
                gw[k] = fmaf(d.x, v.x, fmaf(d.y, v.y, gw[k]));
                gw[5 + k] = fmaf(d.y, v.x, fmaf(-d.x, v.y, gw[5 + k]));
            }
        }
    }
    if constexpr (DW) {
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            const float sres = warp_sum(gw[k]);
            if ((tid & 31) == 0) sred[tid >> 5][k] = sres;
        }
        __syncthreads();
        // rows [0, first_rows) of the partials belong to the stack kernel's CTAs: the FIR gradient of CTA r goes to row r as well
        // (the stack kernel leaves those 10 slots untouched), rows beyond the grid of this kernel get zeros from CTA 0
        if (a.partials && tid < 10) {
            const TcLayout<true> L(a.H);
            a.partials[(size_t)blockIdx.x * L.P + tid] = sred[0][tid] + sred[1][tid] + sred[2][tid] + sred[3][tid];
            if (blockIdx.x == 0)
                for (int r = gridDim.x; r < first_rows; ++r) a.partials[(size_t)r * L.P + tid] = 0.f;
        }
    }
}

// ================================================================ host
static int tc_grid(int B, int T) {
    const int64_t tiles = (int64_t)B * ((T + TC_TT - 1) / TC_TT);
    const int64_t cap = 8 * (int64_t)num_sms();
    return (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
}
int64_t tcnn_nparams(int cell, int C) { return cell == ODPD_CELL_NEURALTX ? TcLayout<true>(C).P : TcLayout<false>(C).P; }
// saved = the five pre-activation maps [B][5][C][T]
int64_t tcnn_saved_floats(int B, int T, int C) { return (int64_t)(B > 0 ? B : 1) * 5 * C * (T > 0 ? T : 1) + 4; }
// workspace = gradient partials [rows][P] (4-aligned) | NeuralTX: dL/d(I_f,Q_f) [B][T] float2
int64_t tcnn_workspace_floats(int cell, int B, int T, int C) {
    const int64_t bt = (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1);
    return (((int64_t)tc_grid(B, T) * tcnn_nparams(cell, C) + 3) & ~(int64_t)3) + 2 * bt + 4;
}

static void tc_ensure_smem(const void *k, size_t bytes) {
    if (bytes <= 48 * 1024) return;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <bool NTX>
static int tcnn_run_t(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out) {
    const TcLayout<NTX> L(a.H);
    const int C = a.H, nts = (a.T + TC_TT - 1) / TC_TT, ntiles = a.B * nts, grid = tc_grid(a.B, a.T);
    const size_t base = (size_t)(((L.P + 3) & ~3) + 2 * C * TC_WP) * sizeof(float);
    if (dir == 0) {
        const size_t sm = base + (size_t)TC_W * 8 * sizeof(float);
        tc_ensure_smem((const void *)tcnn_fwd_kernel<NTX>, sm);
        launch_pdl(tcnn_fwd_kernel<NTX>, dim3(grid), dim3(TC_NT), sm, st, a, nts, ntiles);
        return check_launch("tcnn_fwd_kernel");
    }
    if (!a.saved) { set_error("TCNN / NeuralTX backward needs the activations saved by the forward (ODPD_F_SAVE)"); return -1; }
    if ((dw || NTX) && !a.partials) { set_error("TCNN / NeuralTX backward needs the workspace (odpd_bwd_workspace_bytes)"); return -1; }
    if (a.need_dx && !a.gx) { set_error("TCNN / NeuralTX backward: ODPD_F_NEED_DX without gx"); return -1; }
    float2 *dfir = NTX ? reinterpret_cast<float2 *>(a.partials + (((int64_t)grid * L.P + 3) & ~(int64_t)3)) : nullptr;
    const size_t sm = base + (size_t)(TC_W * 2 + TC_TT * 8) * sizeof(float);
    if (dw) {
        tc_ensure_smem((const void *)tcnn_bwd_kernel<NTX, true>, sm);
        launch_pdl(tcnn_bwd_kernel<NTX, true>, dim3(grid), dim3(TC_NT), sm, st, a, nts, ntiles, dfir);
    } else {
        tc_ensure_smem((const void *)tcnn_bwd_kernel<NTX, false>, sm);
        launch_pdl(tcnn_bwd_kernel<NTX, false>, dim3(grid), dim3(TC_NT), sm, st, a, nts, ntiles, dfir);
    }
    if constexpr (NTX) {
        const int nts2 = (a.T + TC_NT - 1) / TC_NT, ntiles2 = a.B * nts2;
        int g2 = ntiles2 < grid ? ntiles2 : grid;
        if (g2 < 1) g2 = 1;
        if (dw) launch_pdl(tcnn_fir_bwd_kernel<true>, dim3(g2), dim3(TC_NT), 0, st, a, nts2, ntiles2, (const float2 *)dfir, grid);
        else if (a.need_dx) launch_pdl(tcnn_fir_bwd_kernel<false>, dim3(g2), dim3(TC_NT), 0, st, a, nts2, ntiles2, (const float2 *)dfir, grid);
    }
    if (rows_out) *rows_out = grid;
    return check_launch("tcnn backward");
}

int tcnn_run(int cell, const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out) {
    if (a.H < 1 || a.H > TC_CMAX) { set_error("TCNN / NeuralTX: hidden_channels %d outside 1..%d", a.H, TC_CMAX); return -1; }
    return cell == ODPD_CELL_NEURALTX ? tcnn_run_t<true>(a, dir, dw, st, rows_out) : tcnn_run_t<false>(a, dir, dw, st, rows_out);
}

}  // namespace odpd

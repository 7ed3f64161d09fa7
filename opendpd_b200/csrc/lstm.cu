// lstm.cu — fused forward / backward of the LSTM backbones: LSTM and VDLSTM.
// Replaces reference backbones/lstm.py:45-48 (nn.LSTM(2->H) with (h0,c0)=(0,0) + fc_out) and the ATen LSTM cell
// (gate order i,f,g,o; c' = f*c + i*g; h' = o*tanh(c')), plus nn.MSELoss.  Same 3-warp chunk pipeline as gru_family.cu.
// VDLSTM (backbones/vdlstm.py:58-82): the LSTM input is a 4-tap window of the amplitude |x| with WRAP-AROUND
// padding (the last 3 samples of the frame precede sample 0, :65-73); the head is lambda1 = W1 h + b1, lambda2 = W2 h + b2 (H -> 4 each),
// out = W_o [lambda1 * cos_w ; lambda2 * sin_w] + b_o with cos_w = I_w / |x|_w, sin_w = Q_w / |x|_w over the same window.  The recurrence,
// the chunking and the weight-gradient machinery are the LSTM's; the window makes dL/dx a 4-tap scatter with wrap-around, written as
// per-(step, tap) contributions and summed by a gather kernel (ordered, no float atomics).
#include "cells.h"
#include "pipeline.cuh"
#include "chunking.cuh"

namespace odpd {

static constexpr int VW = 4;   // VDLSTM window length (vdlstm.py:11 window_length=4)
template <bool VD>
struct LstmLayoutT {
    static constexpr int F = VD ? VW : 2;
    int H, oWih, oWhh, obih, obhh, oW1, ob1, oW2, ob2, oWo, obo, P;
    __host__ __device__ explicit LstmLayoutT(int h) {
        H = h; oWih = 0; oWhh = 4 * h * F; obih = oWhh + 4 * h * h; obhh = obih + 4 * h;
        if (VD) { oW1 = obhh + 4 * h; ob1 = oW1 + VW * h; oW2 = ob1 + VW; ob2 = oW2 + VW * h; oWo = ob2 + VW; obo = oWo + 2 * 2 * VW; P = obo + 2; }
        else { oW1 = ob1 = oW2 = ob2 = -1; oWo = obhh + 4 * h; obo = oWo + 2 * h; P = obo + 2; }
    }
};
using LstmLayout = LstmLayoutT<false>;
// saved row per step: i | f | g | o | c_t | h_t | tanh(c_t), each HP floats
template <int HT> struct LRow { static constexpr int value = 7 * Pad4<HT>::value; };

template <int HT>
struct LFwdSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = LRow<HT>::value;
    static constexpr int XP = CH * 4 * HP, ACT = CH * ROW, PO = CH * 33, WIN = 3 * 36;     // WIN: |x|, I, Q of a block + 3 halo samples (VDLSTM)
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + ROW + 2 * XP + 2 * ACT + 2 * PO + 3 * WIN; }
};
template <int HT, bool VD = false>
struct LBwdSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = LRow<HT>::value;
    // PRE per step: LSTM  I Q go0 go1;  VDLSTM  |x|_w[4] dlam1[4] dlam2[4] gcos[4] gsin[4] I_w[4] Q_w[4] go[2] -[2]
    static constexpr int ACT = (CH + 1) * ROW, PRE = CH * (VD ? 32 : 4), DH = CH * HP, G = CH * 4 * HP, DF = CH * (VD ? 4 : 2);
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + 3 * ACT + 3 * PRE + 2 * DH + 2 * G + DF; }
};

// the three window samples that precede sample t0 (wrap-around: vdlstm.py:65-73 prepends the LAST W-1 samples of the frame)
__device__ __forceinline__ int vd_wrap(int t, int T) { return t < 0 ? t + T : t; }

template <int HT, bool VD>
__global__ void __launch_bounds__(96, 1) lstm_fwd_kernel(GruArgs a) {
    pdl_enter();   // launched through chunk_launch with the programmatic-serialization attribute (common.cuh)
    constexpr int HP = Pad4<HT>::value, ROW = LRow<HT>::value, F = VD ? VW : 2;
    using SM = LFwdSmem<HT>;
    const LstmLayoutT<VD> L(a.H);
    const int H = a.H, T = a.T;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *zero = sp + Ppad;                 // [ROW] zero row (h_{-1}, c_{-1})
    float *sxp = zero + ROW;                 // [2][CH][4*HP]
    float *sact = sxp + 2 * SM::XP;          // [2][CH][ROW]
    float *spo = sact + 2 * SM::ACT;
    float *swin = spo + 2 * SM::PO;          // [3][3][36]  VDLSTM: |x|, I, Q of samples t0-3 .. t0+31 of a block
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const FwdRange R = fwd_range(a);          // chunking.cuh; recurrent state per chunk = (c, h), 2*HP floats
    const bool spec = R.spec;
    const int b = R.b, t_emit = R.t_emit, t_hi = R.t_hi;
    if (a.mode == 2 && fwd_verify_pass(a, b, 2 * HP, HP, H)) return;
    stage_params(sp, a.params, L.P, bars);
    for (int i = threadIdx.x; i < ROW; i += blockDim.x) zero[i] = 0.f;
    __syncthreads();
    const bool act = lane < H;
    const int j = act ? lane : 0, lp = lane < HP ? lane : 0;
    const int cb = R.t_lo / CH, nchunks = (t_hi + CH - 1) / CH - cb;
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);

    if (warp == 1) {
        float wx[4][F], bb[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int k = 0; k < F; ++k) wx[g][k] = act ? sp[L.oWih + (g * H + j) * F + k] : 0.f;
            bb[g] = act ? sp[L.obih + g * H + j] + sp[L.obhh + g * H + j] : 0.f;
        }
        for (int s = 0; s < nchunks + 2; ++s) {
            if (s < nchunks) {
                const int t0 = (cb + s) * CH, nt = min(CH, t_hi - t0);
                float *xp = sxp + (s & 1) * SM::XP;
                float2 v = make_float2(0.f, 0.f);
                if (lane < nt) v = __ldg(x2 + t0 + lane);
                if constexpr (VD) {
                    // |x| (torch.pow(i,2)+torch.pow(q,2), sqrt: separately rounded like the reference's ATen ops), I, Q of the block
                    // at [3 + lane], of the 3 samples before it (wrapping to the end of the frame) at [lane]
                    float *wa = swin + (s % 3) * SM::WIN, *wi_ = wa + 36, *wq_ = wa + 72;
                    if (lane < nt) { wa[3 + lane] = __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y))); wi_[3 + lane] = v.x; wq_[3 + lane] = v.y; }
                    if (lane < 3) {
                        const float2 h2 = __ldg(x2 + vd_wrap(t0 - 3 + lane, T));
                        wa[lane] = __fsqrt_rn(__fadd_rn(__fmul_rn(h2.x, h2.x), __fmul_rn(h2.y, h2.y))); wi_[lane] = h2.x; wq_[lane] = h2.y;
                    }
                    __syncwarp();
                    if (lane < HP) {
#pragma unroll 4
                        for (int tl = 0; tl < nt; ++tl) {
                            const float aw[4] = {wa[tl], wa[tl + 1], wa[tl + 2], wa[tl + 3]};      // window of step t: samples t-3 .. t
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                xp[tl * 4 * HP + g * HP + lane] = fmaf(wx[g][3], aw[3], fmaf(wx[g][2], aw[2], fmaf(wx[g][1], aw[1], fmaf(wx[g][0], aw[0], bb[g]))));
                        }
                    }
                } else {
                    for (int tl = 0; tl < nt; ++tl) {
                        const float xi = __shfl_sync(ODPD_FULL, v.x, tl), xq = __shfl_sync(ODPD_FULL, v.y, tl);
                        if (lane < HP) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) xp[tl * 4 * HP + g * HP + lane] = fmaf(wx[g][1], xq, fmaf(wx[g][0], xi, bb[g]));
                        }
                    }
                }
            }
            __syncthreads();
        }
    } else if (warp == 0) {
        float wi[HT], wf[HT], wg[HT], wo[HT];
#pragma unroll
        for (int k = 0; k < HT; ++k) {
            const bool ok = act && k < H;
            wi[k] = ok ? sp[L.oWhh + (0 * H + j) * H + k] : 0.f;
            wf[k] = ok ? sp[L.oWhh + (1 * H + j) * H + k] : 0.f;
            wg[k] = ok ? sp[L.oWhh + (2 * H + j) * H + k] : 0.f;
            wo[k] = ok ? sp[L.oWhh + (3 * H + j) * H + k] : 0.f;
        }
        float c = 0.f, hl = 0.f;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int ck = s - 1;
            if (ck >= 0 && ck < nchunks) {
                const int t0 = (cb + ck) * CH, nt = min(CH, t_hi - t0);
                if (spec && R.cc > 0 && t0 == t_emit && lane < HP) {
                    a.sc_guess[(size_t)blockIdx.x * 2 * HP + lane] = c;
                    a.sc_guess[(size_t)blockIdx.x * 2 * HP + HP + lane] = hl;
                }
                const float *xp = sxp + (ck & 1) * SM::XP + lp;
                float *ac = sact + (ck & 1) * SM::ACT;
                const float *hrow = (ck == 0) ? zero : sact + ((ck - 1) & 1) * SM::ACT + (CH - 1) * ROW + 5 * HP;
                float x0 = xp[0], x1 = xp[HP], x2v = xp[2 * HP], x3 = xp[3 * HP];
                for (int tl = 0; tl < nt; ++tl) {
                    const int tn = (tl + 1 < nt) ? tl + 1 : tl;
                    const float n0 = xp[tn * 4 * HP], n1 = xp[tn * 4 * HP + HP], n2 = xp[tn * 4 * HP + 2 * HP], n3 = xp[tn * 4 * HP + 3 * HP];
                    float ai0 = x0, ai1 = 0.f, af0 = x1, af1 = 0.f, ag0 = x2v, ag1 = 0.f, ao0 = x3, ao1 = 0.f;
                    bcast_dot<HT>(hrow, wi, ai0, ai1);
                    bcast_dot<HT>(hrow, wf, af0, af1);
                    bcast_dot<HT>(hrow, wg, ag0, ag1);
                    bcast_dot<HT>(hrow, wo, ao0, ao1);
                    const float ig = sigmoidf_(ai0 + ai1), fg = sigmoidf_(af0 + af1), gg = tanhf_(ag0 + ag1), og = sigmoidf_(ao0 + ao1);
                    c = fmaf(fg, c, ig * gg);
                    const float tc = tanhf_(c);
                    const float h = og * tc;
                    hl = h;
                    float *row = ac + tl * ROW;
                    if (lane < HP) {
                        row[5 * HP + lane] = h;
                        row[lane] = ig; row[HP + lane] = fg; row[2 * HP + lane] = gg; row[3 * HP + lane] = og; row[4 * HP + lane] = c;
                        row[6 * HP + lane] = tc;
                    }
                    hrow = row + 5 * HP;
                    x0 = n0; x1 = n1; x2v = n2; x3 = n3;
                    __syncwarp();
                }
                fence_async_smem();
            }
            __syncthreads();
        }
        if (spec && lane < HP) {
            a.sc_end[(size_t)blockIdx.x * 2 * HP + lane] = c;
            a.sc_end[(size_t)blockIdx.x * 2 * HP + HP + lane] = hl;
        }
    } else {
        const float wo0 = (!VD && act) ? sp[L.oWo + j] : 0.f, wo1 = (!VD && act) ? sp[L.oWo + H + j] : 0.f;
        const float bo0 = sp[L.obo], bo1 = sp[L.obo + 1];
        const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
        float2 *o2 = reinterpret_cast<float2 *>(a.out) + (size_t)b * T;
        float *svg = a.save ? a.saved + (size_t)b * T * ROW : nullptr;
        float lsum = 0.f;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int ck = s - 2;
            if (ck >= 0 && (cb + ck) * CH >= t_emit) {     // warm-up blocks emit nothing
                const int t0 = (cb + ck) * CH, nt = min(CH, t_hi - t0);
                float *ac = sact + (ck & 1) * SM::ACT;
                if (svg && lane == 0) tma_store_1d(svg + (size_t)t0 * ROW, ac, (uint32_t)(nt * ROW * 4));
                if constexpr (VD) {
                    // VDLSTM head, one timestep per lane: lambda1/2 = W1/2 h + b1/2; out = W_o [lambda1*cos_w ; lambda2*sin_w] + b_o
                    if (lane < nt) {
                        const float *wa = swin + (ck % 3) * SM::WIN, *wi_ = wa + 36, *wq_ = wa + 72;
                        const float *hrow = ac + lane * ROW + 5 * HP;
                        float l1[VW], l2[VW];
#pragma unroll
                        for (int k = 0; k < VW; ++k) { l1[k] = sp[L.ob1 + k]; l2[k] = sp[L.ob2 + k]; }
                        for (int q = 0; q < H; ++q) {
                            const float hv = hrow[q];
#pragma unroll
                            for (int k = 0; k < VW; ++k) { l1[k] = fmaf(sp[L.oW1 + k * H + q], hv, l1[k]); l2[k] = fmaf(sp[L.oW2 + k * H + q], hv, l2[k]); }
                        }
                        float o0 = bo0, o1 = bo1;
#pragma unroll
                        for (int k = 0; k < VW; ++k) {
                            const float am = wa[lane + k];
                            const float fc = l1[k] * __fdiv_rn(wi_[lane + k], am), fs = l2[k] * __fdiv_rn(wq_[lane + k], am);
                            o0 = fmaf(sp[L.oWo + k], fc, o0); o1 = fmaf(sp[L.oWo + 2 * VW + k], fc, o1);
                            o0 = fmaf(sp[L.oWo + VW + k], fs, o0); o1 = fmaf(sp[L.oWo + 3 * VW + k], fs, o1);
                        }
                        o2[t0 + lane] = make_float2(o0, o1);
                        if (y2) {
                            const float2 y = __ldg(y2 + t0 + lane);
                            const float d0 = o0 - y.x, d1 = o1 - y.y;
                            lsum = fmaf(d0, d0, fmaf(d1, d1, lsum));
                        }
                    }
                } else {
                    linear_head_chunk(ac, ROW, 5 * HP, HP, H, nt, lane, wo0, wo1, bo0, bo1, spo, nullptr, o2 + t0, y2 ? y2 + t0 : iq_none(), lsum);
                }
                if (svg && lane == 0) tma_store_wait_read();
                __syncwarp();
            }
            __syncthreads();
        }
        if (y2) {
            lsum = warp_sum(lsum);
            if (lane == 0) chunk_store_loss(a, spec, lsum);
        }
    }
}

template <int HT, bool DW, bool VD>
__global__ void __launch_bounds__(96, 1) lstm_bwd_kernel(GruArgs a) {
    pdl_enter();   // launched through chunk_launch with the programmatic-serialization attribute (common.cuh)
    constexpr int HP = Pad4<HT>::value, ROW = LRow<HT>::value, F = VD ? VW : 2, PS = VD ? 32 : 4;
    using SM = LBwdSmem<HT, VD>;
    const LstmLayoutT<VD> L(a.H);
    const int H = a.H, T = a.T;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *sact = sp + Ppad;                 // [3][CH+1][ROW]
    float *spre = sact + 3 * SM::ACT;        // [3][CH][PS]: I Q go0 go1  (VDLSTM: see LBwdSmem)
    float *sdh = spre + 3 * SM::PRE;         // [2][CH][HP]
    float *sG = sdh + 2 * SM::DH;            // [2][CH][4HP]: di df dg do
    float *sdf = sG + 2 * SM::G;             // [CH][F]  dL/d(input features) of the block
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const BwdRange R = bwd_range(a);          // chunking.cuh; adjoint state per chunk = (dL/dh, dL/dc), 2*HP floats
    const bool spec = R.spec;
    const int b = R.b, t_ehi = R.t_ehi, t_hi = R.t_hi;
    if (a.mode == 2 && bwd_verify_pass(a, b, 2 * HP, HP, H)) return;
    if (threadIdx.x == 0) { mbar_init(bars + 1, 1); mbar_init(bars + 2, 1); mbar_init(bars + 3, 1); }
    stage_params(sp, a.params, L.P, bars);
    const bool act = lane < H;
    const int j = act ? lane : 0, lp = lane < HP ? lane : 0;
    // 32-step blocks [cb, ce) are processed last to first; blocks >= ce_emit are warm-up
    const int cb = R.t_elo / CH, ce = (t_hi + CH - 1) / CH, nchunks = ce - cb, ce_emit = (t_ehi + CH - 1) / CH;
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
    const float *svg = a.saved + (size_t)b * T * ROW;

    if (warp == 1) {
        const float wo0 = (!VD && act) ? sp[L.oWo + j] : 0.f, wo1 = (!VD && act) ? sp[L.oWo + H + j] : 0.f;
        const float2 *go2 = a.gout ? reinterpret_cast<const float2 *>(a.gout) + (size_t)b * T : nullptr;
        const float2 *oi2 = a.out_in ? reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * T : nullptr;
        const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
        const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
        float gwo[(VD && DW) ? 4 * VW : 1], gbo0 = 0.f, gbo1 = 0.f;      // VDLSTM: dL/dfc_out accumulates here (one timestep per lane)
        if constexpr (VD && DW) {
#pragma unroll
            for (int q = 0; q < 4 * VW; ++q) gwo[q] = 0.f;
        }
        for (int s = 0; s < nchunks + 2; ++s) {
            if (s < nchunks) {
                const int c = ce - 1 - s, t0 = c * CH, nt = min(CH, t_hi - t0), slot = s % 3;
                float *ac = sact + slot * SM::ACT, *pr = spre + slot * SM::PRE, *dh = sdh + (s & 1) * SM::DH;
                uint64_t *bar = bars + 1 + slot;
                load_rows_with_prev(ac, svg, ROW, t0, nt, lane, bar);
                if constexpr (VD) {
                    // one timestep per lane: window of the sample, dLoss/dout, lambda1/2 (recomputed from the saved h), and everything the
                    // chain (dL/dh) and the post warp (weight gradients, window gradients) need, parked in spre
                    float iw[VW], qw[VW], am[VW];
                    float2 g = make_float2(0.f, 0.f);
                    if (lane < nt) {
                        const int t = t0 + lane;
#pragma unroll
                        for (int k = 0; k < VW; ++k) {
                            const float2 v = __ldg(x2 + vd_wrap(t - (VW - 1) + k, T));
                            iw[k] = v.x; qw[k] = v.y;
                            am[k] = __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));
                        }
                        g = load_gout(go2, oi2, y2, t, gs);
                    }
                    mbar_wait(bar, (uint32_t)((s / 3) & 1));
                    if (lane < nt) {
                        const float *hrow = ac + (lane + 1) * ROW + 5 * HP;
                        float l1[VW], l2[VW];
#pragma unroll
                        for (int k = 0; k < VW; ++k) { l1[k] = sp[L.ob1 + k]; l2[k] = sp[L.ob2 + k]; }
                        for (int q = 0; q < H; ++q) {
                            const float hv = hrow[q];
#pragma unroll
                            for (int k = 0; k < VW; ++k) { l1[k] = fmaf(sp[L.oW1 + k * H + q], hv, l1[k]); l2[k] = fmaf(sp[L.oW2 + k * H + q], hv, l2[k]); }
                        }
                        float dl1[VW], dl2[VW], gc[VW], gsn[VW];
#pragma unroll
                        for (int k = 0; k < VW; ++k) {
                            const float cs = __fdiv_rn(iw[k], am[k]), sn = __fdiv_rn(qw[k], am[k]);
                            const float dzc = fmaf(sp[L.oWo + k], g.x, sp[L.oWo + 2 * VW + k] * g.y);
                            const float dzs = fmaf(sp[L.oWo + VW + k], g.x, sp[L.oWo + 3 * VW + k] * g.y);
                            dl1[k] = dzc * cs; dl2[k] = dzs * sn; gc[k] = dzc * l1[k]; gsn[k] = dzs * l2[k];
                            if (DW && c < ce_emit) {            // warm-up blocks emit nothing
                                const float fc = l1[k] * cs, fs = l2[k] * sn;
                                gwo[k] = fmaf(g.x, fc, gwo[k]); gwo[VW + k] = fmaf(g.x, fs, gwo[VW + k]);
                                gwo[2 * VW + k] = fmaf(g.y, fc, gwo[2 * VW + k]); gwo[3 * VW + k] = fmaf(g.y, fs, gwo[3 * VW + k]);
                            }
                        }
                        if (DW && c < ce_emit) { gbo0 += g.x; gbo1 += g.y; }
                        float4 *p4 = reinterpret_cast<float4 *>(pr + lane * PS);
                        p4[0] = make_float4(am[0], am[1], am[2], am[3]);
                        p4[1] = make_float4(dl1[0], dl1[1], dl1[2], dl1[3]);
                        p4[2] = make_float4(dl2[0], dl2[1], dl2[2], dl2[3]);
                        p4[3] = make_float4(gc[0], gc[1], gc[2], gc[3]);
                        p4[4] = make_float4(gsn[0], gsn[1], gsn[2], gsn[3]);
                        p4[5] = make_float4(iw[0], iw[1], iw[2], iw[3]);
                        p4[6] = make_float4(qw[0], qw[1], qw[2], qw[3]);
                        p4[7] = make_float4(g.x, g.y, 0.f, 0.f);
                        for (int q = 0; q < H; ++q) {       // dL/dh_t from the head: W1^T dlam1 + W2^T dlam2
                            float d = 0.f;
#pragma unroll
                            for (int k = 0; k < VW; ++k) d = fmaf(sp[L.oW2 + k * H + q], dl2[k], fmaf(sp[L.oW1 + k * H + q], dl1[k], d));
                            dh[lane * HP + q] = d;
                        }
                    }
                } else {
                    if (lane < nt) {
                        const float2 v = __ldg(x2 + t0 + lane);
                        const float2 g = load_gout(go2, oi2, y2, t0 + lane, gs);
                        *reinterpret_cast<float4 *>(pr + lane * 4) = make_float4(v.x, v.y, g.x, g.y);
                    }
                    __syncwarp();
                    if (lane < HP)
                        for (int tl = 0; tl < nt; ++tl) dh[tl * HP + lane] = fmaf(wo0, pr[tl * 4 + 2], wo1 * pr[tl * 4 + 3]);
                    mbar_wait(bar, (uint32_t)((s / 3) & 1));
                }
            }
            __syncthreads();
        }
        if constexpr (VD && DW) {
            if (a.partials) {
                float *prt = a.partials + (size_t)(spec ? blockIdx.x : b * a.C) * L.P;
#pragma unroll
                for (int q = 0; q < 4 * VW; ++q) {
                    const float sum = warp_sum(gwo[q]);       // q = o*8 + part*4 + k  -> fc_out.weight[o][part*4 + k]
                    if (lane == 0) prt[L.oWo + q] = sum;
                }
                gbo0 = warp_sum(gbo0); gbo1 = warp_sum(gbo1);
                if (lane == 0) { prt[L.obo] = gbo0; prt[L.obo + 1] = gbo1; }
            }
        }
    } else if (warp == 0) {
        float wc0[HT], wc1[HT], wc2[HT], wc3[HT];   // column j of W_hh, per gate
#pragma unroll
        for (int k = 0; k < HT; ++k) {
            const bool ok = act && k < H;
            wc0[k] = ok ? sp[L.oWhh + (0 * H + k) * H + j] : 0.f;
            wc1[k] = ok ? sp[L.oWhh + (1 * H + k) * H + j] : 0.f;
            wc2[k] = ok ? sp[L.oWhh + (2 * H + k) * H + j] : 0.f;
            wc3[k] = ok ? sp[L.oWhh + (3 * H + k) * H + j] : 0.f;
        }
        float gH = 0.f, gC = 0.f;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 1;
            if (sc >= 0 && sc < nchunks) {
                const int c = ce - 1 - sc, t0 = c * CH, nt = min(CH, t_hi - t0);
                if (spec && c == ce_emit - 1 && t_ehi < T && lane < HP) {
                    a.sc_guess[(size_t)blockIdx.x * 2 * HP + lane] = gH;
                    a.sc_guess[(size_t)blockIdx.x * 2 * HP + HP + lane] = gC;
                }
                const float *ac = sact + (sc % 3) * SM::ACT + lp;
                const float *dh = sdh + (sc & 1) * SM::DH + lp;
                float *Gb = sG + (sc & 1) * SM::G;
                const float *row = ac + nt * ROW;
                float ig = row[0], fg = row[HP], gg = row[2 * HP], og = row[3 * HP], cp = row[4 * HP - ROW], tc = row[6 * HP], dht = dh[(nt - 1) * HP];
                for (int tl = nt - 1; tl >= 0; --tl) {
                    const int tp = tl > 0 ? tl - 1 : 0;
                    const float *rn = ac + (tp + 1) * ROW;
                    const float ig_n = rn[0], fg_n = rn[HP], gg_n = rn[2 * HP], og_n = rn[3 * HP], cp_n = rn[4 * HP - ROW], tc_n = rn[6 * HP],
                                dh_n = dh[tp * HP];
                    gH += dht;
                    const float go_ = gH * tc;
                    const float gc = fmaf(gH * og, 1.f - tc * tc, gC);
                    const float di = gc * gg * ig * (1.f - ig);
                    const float df = gc * cp * fg * (1.f - fg);
                    const float dg = gc * ig * (1.f - gg * gg);
                    const float dO = go_ * og * (1.f - og);
                    gC = gc * fg;
                    float *G = Gb + tl * 4 * HP;
                    if (lane < HP) { G[lane] = di; G[HP + lane] = df; G[2 * HP + lane] = dg; G[3 * HP + lane] = dO; }
                    __syncwarp();
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                    bcast_dot<HT>(G, wc0, a0, a1);
                    bcast_dot<HT>(G + HP, wc1, a2, a3);
                    bcast_dot<HT>(G + 2 * HP, wc2, a0, a1);
                    bcast_dot<HT>(G + 3 * HP, wc3, a2, a3);
                    gH = (a0 + a1) + (a2 + a3);
                    ig = ig_n; fg = fg_n; gg = gg_n; og = og_n; cp = cp_n; tc = tc_n; dht = dh_n;
                }
            }
            __syncthreads();
        }
        if (spec && lane < HP) {
            a.sc_end[(size_t)blockIdx.x * 2 * HP + lane] = gH;
            a.sc_end[(size_t)blockIdx.x * 2 * HP + HP + lane] = gC;
        }
    } else {
        const int fl = lane - H;                       // lanes H .. H+F-1 serve dL/d(input features)   (lanes 0.. when they do not fit)
        const bool split = (H + F > 32);
        const int fidx = split ? lane : fl;
        const bool isf = fidx >= 0 && fidx < F;
        float wic[4 * HT];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int k = 0; k < HT; ++k) wic[g * HT + k] = (isf && k < H) ? sp[L.oWih + (g * H + k) * F + fidx] : 0.f;
        float gwhh[DW ? 4 * HT : 1], gwih[4 * F], gb[4], gwo0 = 0.f, gwo1 = 0.f, gbo0 = 0.f, gbo1 = 0.f;
        float gw1[(VD && DW) ? VW : 1], gw2[(VD && DW) ? VW : 1], gb1[(VD && DW) ? VW : 1], gb2[(VD && DW) ? VW : 1];
        if constexpr (DW) {
#pragma unroll
            for (int k = 0; k < 4 * HT; ++k) gwhh[k] = 0.f;
            if constexpr (VD) {
#pragma unroll
                for (int k = 0; k < VW; ++k) { gw1[k] = gw2[k] = gb1[k] = gb2[k] = 0.f; }
            }
        }
#pragma unroll
        for (int k = 0; k < 4 * F; ++k) gwih[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) gb[k] = 0.f;
        float2 *gx2 = (a.need_dx && a.gx) ? reinterpret_cast<float2 *>(a.gx) + (size_t)b * T : nullptr;
        float2 *ctr = (VD && a.need_dx && a.vd_contrib) ? reinterpret_cast<float2 *>(a.vd_contrib) + (size_t)b * T * VW : nullptr;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 2;
            if (sc >= 0 && ce - 1 - sc < ce_emit) {          // warm-up blocks emit nothing
                const int c = ce - 1 - sc, t0 = c * CH, nt = min(CH, t_hi - t0);
                const float *ac = sact + (sc % 3) * SM::ACT, *pr = spre + (sc % 3) * SM::PRE, *Gb = sG + (sc & 1) * SM::G;
                for (int tl = 0; tl < nt; ++tl) {
                    const float *G = Gb + tl * 4 * HP;
                    const float *row = ac + (tl + 1) * ROW;
                    const float hp = row[5 * HP - ROW + lp], ht = row[5 * HP + lp];
                    const float4 p = *reinterpret_cast<const float4 *>(pr + tl * PS);      // LSTM: I Q go0 go1;  VDLSTM: |x| window
                    if constexpr (DW) {
                        const float fin[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float own = G[g * HP + lp];
#pragma unroll
                            for (int k = 0; k < F; ++k) gwih[g * F + k] = fmaf(own, fin[k], gwih[g * F + k]);
                            gb[g] += own;
                        }
                        if constexpr (VD) {
                            const float4 d1 = *reinterpret_cast<const float4 *>(pr + tl * PS + 4), d2 = *reinterpret_cast<const float4 *>(pr + tl * PS + 8);
                            const float e1[4] = {d1.x, d1.y, d1.z, d1.w}, e2[4] = {d2.x, d2.y, d2.z, d2.w};
#pragma unroll
                            for (int k = 0; k < VW; ++k) {
                                gw1[k] = fmaf(e1[k], ht, gw1[k]); gw2[k] = fmaf(e2[k], ht, gw2[k]);
                                gb1[k] += e1[k]; gb2[k] += e2[k];
                            }
                        } else {
                            gwo0 = fmaf(p.z, ht, gwo0); gwo1 = fmaf(p.w, ht, gwo1);
                        }
                    }
                    float fa0 = 0.f, fa1 = 0.f;
                    const float4 *G4 = reinterpret_cast<const float4 *>(G);
#pragma unroll
                    for (int g = 0; g < 4; ++g)
#pragma unroll
                        for (int k4 = 0; k4 < HP / 4; ++k4) {
                            const float4 v = G4[g * (HP / 4) + k4];
                            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int k = k4 * 4 + i;
                                if (k < HT) {
                                    if (g & 1) fa1 = fmaf(wic[g * HT + k], e[i], fa1); else fa0 = fmaf(wic[g * HT + k], e[i], fa0);
                                    if constexpr (DW) gwhh[g * HT + k] = fmaf(e[i], hp, gwhh[g * HT + k]);
                                }
                            }
                        }
                    if (isf) sdf[tl * F + fidx] = fa0 + fa1;
                }
                __syncwarp();
                if (lane < nt) {
                    if constexpr (VD) {
                        if (ctr) {
                            // window quantities -> samples: |x| = sqrt(I^2+Q^2), cos = I/|x|, sin = Q/|x| at window position (t,k) <-> sample
                            // (t-3+k) mod T.  One contribution per (step, tap); vd_gather_kernel sums the four taps of a sample in tap order.
                            const float *q = pr + lane * PS;
#pragma unroll
                            for (int k = 0; k < VW; ++k) {
                                const float am = q[k], gc = q[12 + k], gsn = q[16 + k], I = q[20 + k], Q = q[24 + k], ga = sdf[lane * F + k];
                                const float ia = 1.f / am, ia3 = ia * ia * ia;
                                const float gI = ga * I * ia + gc * (ia - I * I * ia3) - gsn * (Q * I * ia3);
                                const float gQ = ga * Q * ia - gc * (I * Q * ia3) + gsn * (ia - Q * Q * ia3);
                                ctr[(size_t)(t0 + lane) * VW + k] = make_float2(gI, gQ);
                            }
                        }
                    } else {
                        if constexpr (DW) { gbo0 += pr[lane * 4 + 2]; gbo1 += pr[lane * 4 + 3]; }
                        if (gx2) gx2[t0 + lane] = *reinterpret_cast<const float2 *>(sdf + lane * 2);
                    }
                }
                __syncwarp();
            }
            __syncthreads();
        }
        if constexpr (DW) {
            if (a.partials) {
                float *prt = chunk_partial_row(a, spec, b, L.P, lane, 32);
                if (act) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
#pragma unroll
                        for (int k = 0; k < F; ++k) prt[L.oWih + (g * H + lane) * F + k] = gwih[g * F + k];
                        prt[L.obih + g * H + lane] = gb[g];
                        prt[L.obhh + g * H + lane] = gb[g];
#pragma unroll
                        for (int k = 0; k < HT; ++k)
                            if (k < H) prt[L.oWhh + (g * H + k) * H + lane] = gwhh[g * HT + k];
                    }
                    if constexpr (VD) {
#pragma unroll
                        for (int k = 0; k < VW; ++k) { prt[L.oW1 + k * H + lane] = gw1[k]; prt[L.oW2 + k * H + lane] = gw2[k]; }
                    } else {
                        prt[L.oWo + lane] = gwo0; prt[L.oWo + H + lane] = gwo1;
                    }
                }
                if constexpr (VD) {
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < VW; ++k) { prt[L.ob1 + k] = gb1[k]; prt[L.ob2 + k] = gb2[k]; }
                    }
                } else {
                    gbo0 = warp_sum(gbo0); gbo1 = warp_sum(gbo1);
                    if (lane == 0) { prt[L.obo] = gbo0; prt[L.obo + 1] = gbo1; }
                }
            }
        }
    }
}

// VDLSTM: dL/dx[s] = sum over the four window taps that touch sample s: step (s+3-k) mod T, tap k  (fixed order: bit-reproducible)
__global__ void vd_gather_kernel(const float2 *__restrict__ ctr, float2 *__restrict__ gx, int B, int T) {
    pdl_enter();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * T) return;
    const int b = (int)(i / T), sidx = (int)(i - (int64_t)b * T);
    float gi = 0.f, gq = 0.f;
#pragma unroll
    for (int k = 0; k < VW; ++k) {
        int t = sidx + (VW - 1) - k;
        if (t >= T) t -= T;
        const float2 v = ctr[((size_t)b * T + t) * VW + k];
        gi += v.x; gq += v.y;
    }
    gx[i] = make_float2(gi, gq);
}

#define ODPD_LSTM_TIERS(X) X(9) X(16) X(32)      // 9: the LSTM/VDLSTM size of train_all_pa.sh / train_all_dpd.sh
static int lstm_tier(int H) {
#define X(HTV) if (H <= HTV) return HTV;
    ODPD_LSTM_TIERS(X)
#undef X
    return -1;
}
// dir: 0 fwd, 1 bwd, +2 = plan only.   backward workspace (floats): [rows][P] partials (4-aligned) | chunk scratch | VDLSTM: [B][T][4] float2 window contributions
template <int HT, bool VD> static int lstm_launch(GruArgs a, int dir, bool dw, cudaStream_t st, int *info) {
    constexpr int HP = Pad4<HT>::value, ROW = LRow<HT>::value;
    const LstmLayoutT<VD> L(a.H);
    const int Ppad = (L.P + 3) & ~3;
    const bool plan_only = dir >= 2;
    if ((dir & 1) == 0) {
        const size_t smem = (size_t)LFwdSmem<HT>::total(Ppad) * 4;
        static OccCache occ{};
        const int64_t soff = a.save ? (int64_t)a.B * a.T * ROW : 0;
        return chunk_launch(lstm_fwd_kernel<HT, VD>, 96, smem, &occ, a, 0, a.saved ? a.saved + soff : nullptr, soff, 2 * HP, st, plan_only, info,
                            VD ? "vdlstm_fwd_kernel" : "lstm_fwd_kernel");
    }
    const size_t smem = (size_t)LBwdSmem<HT, VD>::total(Ppad) * 4;
    const int64_t rows = chunk_rows(a.B, a.tchunks_req);
    const int64_t woff = (rows * L.P + 3) & ~(int64_t)3;
    float *scr = a.partials ? a.partials + woff : nullptr;
    if (VD) a.vd_contrib = scr ? scr + chunk_bwd_scratch_floats(rows, 2 * HP) : nullptr;
    int rc;
    if (dw) {
        static OccCache occ{};
        rc = chunk_launch(lstm_bwd_kernel<HT, true, VD>, 96, smem, &occ, a, 1, scr, woff, 2 * HP, st, plan_only, info, "lstm_bwd_kernel");
    } else {
        static OccCache occ0{};
        rc = chunk_launch(lstm_bwd_kernel<HT, false, VD>, 96, smem, &occ0, a, 1, scr, woff, 2 * HP, st, plan_only, info, "lstm_bwd_kernel");
    }
    if (rc || plan_only || !VD || !a.need_dx || !a.gx) return rc;
    if (!a.vd_contrib) { set_error("VDLSTM backward with dL/dx needs the workspace"); return -1; }
    const int64_t n = (int64_t)a.B * a.T;
    launch_pdl(vd_gather_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, reinterpret_cast<const float2 *>(a.vd_contrib),
               reinterpret_cast<float2 *>(a.gx), a.B, a.T);
    return check_launch("vd_gather_kernel");
}
int64_t lstm_nparams(int H, bool vd) { return vd ? LstmLayoutT<true>(H).P : LstmLayoutT<false>(H).P; }
int64_t lstm_saved_floats(int B, int T, int H, bool save, int tchunks_req) {
    const int ht = lstm_tier(H);
    if (ht < 0) return -1;
    const int HP = (ht + 3) & ~3;
    return (save ? (int64_t)B * T * 7 * HP : 0) + chunk_fwd_scratch_floats(chunk_rows(B, tchunks_req), 2 * HP);
}
int64_t lstm_workspace_floats(int B, int T, int H, int64_t P, int tchunks_req, bool vd) {
    const int ht = lstm_tier(H);
    if (ht < 0) return -1;
    return chunk_workspace_floats(chunk_rows(B, tchunks_req), P, 2 * ((ht + 3) & ~3)) + (vd ? (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1) * 2 * VW : 0);
}
int lstm_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *info) {
    const bool vd = a.cell == ODPD_CELL_VDLSTM;
    if (vd && a.T < VW - 1) { set_error("VDLSTM needs frame_length >= %d (the window wraps over the last %d samples, vdlstm.py:65-73; got %d)", VW - 1, VW - 1, a.T); return -1; }
#define X(HTV) if (a.H <= HTV) return vd ? lstm_launch<HTV, true>(a, dir, dw, st, info) : lstm_launch<HTV, false>(a, dir, dw, st, info);
    ODPD_LSTM_TIERS(X)
#undef X
    set_error("LSTM kernels support hidden_size <= 32 (got %d)", a.H);
    return -1;
}

}  // namespace odpd

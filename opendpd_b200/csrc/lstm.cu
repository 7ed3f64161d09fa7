// lstm.cu — fused forward / backward of the LSTM backbone.
// Replaces reference backbones/lstm.py:45-48 (nn.LSTM(2->H) with (h0,c0)=(0,0) + fc_out) and the ATen LSTM cell
// (gate order i,f,g,o; c' = f*c + i*g; h' = o*tanh(c')), plus nn.MSELoss.  Same 3-warp chunk pipeline as gru_family.cu.
#include "cells.h"
#include "pipeline.cuh"
#include "chunking.cuh"

namespace odpd {

struct LstmLayout {
    int H, oWih, oWhh, obih, obhh, oWo, obo, P;
    __host__ __device__ explicit LstmLayout(int h) {
        H = h; oWih = 0; oWhh = 8 * h; obih = oWhh + 4 * h * h; obhh = obih + 4 * h; oWo = obhh + 4 * h; obo = oWo + 2 * h; P = obo + 2;
    }
};
// saved row per step: i | f | g | o | c_t | h_t | tanh(c_t), each HP floats
template <int HT> struct LRow { static constexpr int value = 7 * Pad4<HT>::value; };

template <int HT>
struct LFwdSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = LRow<HT>::value;
    static constexpr int XP = CH * 4 * HP, ACT = CH * ROW, PO = CH * 33;
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + ROW + 2 * XP + 2 * ACT + 2 * PO; }
};
template <int HT>
struct LBwdSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = LRow<HT>::value;
    static constexpr int ACT = (CH + 1) * ROW, PRE = CH * 4, DH = CH * HP, G = CH * 4 * HP, DF = CH * 2;
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + 3 * ACT + 3 * PRE + 2 * DH + 2 * G + DF; }
};

template <int HT>
__global__ void __launch_bounds__(96, 1) lstm_fwd_kernel(GruArgs a) {
    pdl_enter();   // launched through chunk_launch with the programmatic-serialization attribute (common.cuh)
    constexpr int HP = Pad4<HT>::value, ROW = LRow<HT>::value;
    using SM = LFwdSmem<HT>;
    const LstmLayout L(a.H);
    const int H = a.H, T = a.T;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *zero = sp + Ppad;                 // [ROW] zero row (h_{-1}, c_{-1})
    float *sxp = zero + ROW;                 // [2][CH][4*HP]
    float *sact = sxp + 2 * SM::XP;          // [2][CH][ROW]
    float *spo = sact + 2 * SM::ACT;
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
        float w0[4], w1[4], bb[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            w0[g] = act ? sp[L.oWih + (g * H + j) * 2] : 0.f;
            w1[g] = act ? sp[L.oWih + (g * H + j) * 2 + 1] : 0.f;
            bb[g] = act ? sp[L.obih + g * H + j] + sp[L.obhh + g * H + j] : 0.f;
        }
        for (int s = 0; s < nchunks + 2; ++s) {
            if (s < nchunks) {
                const int t0 = (cb + s) * CH, nt = min(CH, t_hi - t0);
                float *xp = sxp + (s & 1) * SM::XP;
                float2 v = make_float2(0.f, 0.f);
                if (lane < nt) v = __ldg(x2 + t0 + lane);
                for (int tl = 0; tl < nt; ++tl) {
                    const float xi = __shfl_sync(ODPD_FULL, v.x, tl), xq = __shfl_sync(ODPD_FULL, v.y, tl);
                    if (lane < HP) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) xp[tl * 4 * HP + g * HP + lane] = fmaf(w1[g], xq, fmaf(w0[g], xi, bb[g]));
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
        const float wo0 = act ? sp[L.oWo + j] : 0.f, wo1 = act ? sp[L.oWo + H + j] : 0.f;
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
                linear_head_chunk(ac, ROW, 5 * HP, HP, H, nt, lane, wo0, wo1, bo0, bo1, spo, nullptr, o2 + t0, y2 ? y2 + t0 : iq_none(), lsum);
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

template <int HT, bool DW>
__global__ void __launch_bounds__(96, 1) lstm_bwd_kernel(GruArgs a) {
    pdl_enter();   // launched through chunk_launch with the programmatic-serialization attribute (common.cuh)
    constexpr int HP = Pad4<HT>::value, ROW = LRow<HT>::value;
    using SM = LBwdSmem<HT>;
    const LstmLayout L(a.H);
    const int H = a.H, T = a.T;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *sact = sp + Ppad;                 // [3][CH+1][ROW]
    float *spre = sact + 3 * SM::ACT;        // [3][CH][4]: I Q go0 go1
    float *sdh = spre + 3 * SM::PRE;         // [2][CH][HP]
    float *sG = sdh + 2 * SM::DH;            // [2][CH][4HP]: di df dg do
    float *sdf = sG + 2 * SM::G;             // [CH][2]
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
        const float wo0 = act ? sp[L.oWo + j] : 0.f, wo1 = act ? sp[L.oWo + H + j] : 0.f;
        const float2 *go2 = a.gout ? reinterpret_cast<const float2 *>(a.gout) + (size_t)b * T : nullptr;
        const float2 *oi2 = a.out_in ? reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * T : nullptr;
        const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
        const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
        for (int s = 0; s < nchunks + 2; ++s) {
            if (s < nchunks) {
                const int c = ce - 1 - s, t0 = c * CH, nt = min(CH, t_hi - t0), slot = s % 3;
                float *ac = sact + slot * SM::ACT, *pr = spre + slot * SM::PRE, *dh = sdh + (s & 1) * SM::DH;
                uint64_t *bar = bars + 1 + slot;
                load_rows_with_prev(ac, svg, ROW, t0, nt, lane, bar);
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
            __syncthreads();
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
        const int fl = lane - H;                       // lanes H, H+1 serve dL/dI, dL/dQ   (H<=30; H=31,32 use lanes 0,1 below)
        const bool split = (H + 2 > 32);
        const int fidx = split ? lane : fl;
        const bool isf = fidx >= 0 && fidx < 2;
        float wic[4 * HT];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int k = 0; k < HT; ++k) wic[g * HT + k] = (isf && k < H) ? sp[L.oWih + (g * H + k) * 2 + fidx] : 0.f;
        float gwhh[DW ? 4 * HT : 1], gwih[8], gb[4], gwo0 = 0.f, gwo1 = 0.f, gbo0 = 0.f, gbo1 = 0.f;
        if constexpr (DW) {
#pragma unroll
            for (int k = 0; k < 4 * HT; ++k) gwhh[k] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) gwih[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) gb[k] = 0.f;
        float2 *gx2 = (a.need_dx && a.gx) ? reinterpret_cast<float2 *>(a.gx) + (size_t)b * T : nullptr;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 2;
            if (sc >= 0 && ce - 1 - sc < ce_emit) {          // warm-up blocks emit nothing
                const int c = ce - 1 - sc, t0 = c * CH, nt = min(CH, t_hi - t0);
                const float *ac = sact + (sc % 3) * SM::ACT, *pr = spre + (sc % 3) * SM::PRE, *Gb = sG + (sc & 1) * SM::G;
                for (int tl = 0; tl < nt; ++tl) {
                    const float *G = Gb + tl * 4 * HP;
                    const float *row = ac + (tl + 1) * ROW;
                    const float hp = row[5 * HP - ROW + lp], ht = row[5 * HP + lp];
                    const float4 p = *reinterpret_cast<const float4 *>(pr + tl * 4);
                    if constexpr (DW) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float own = G[g * HP + lp];
                            gwih[g * 2] = fmaf(own, p.x, gwih[g * 2]);
                            gwih[g * 2 + 1] = fmaf(own, p.y, gwih[g * 2 + 1]);
                            gb[g] += own;
                        }
                        gwo0 = fmaf(p.z, ht, gwo0); gwo1 = fmaf(p.w, ht, gwo1);
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
                    if (isf) sdf[tl * 2 + fidx] = fa0 + fa1;
                }
                __syncwarp();
                if (lane < nt) {
                    if constexpr (DW) { gbo0 += pr[lane * 4 + 2]; gbo1 += pr[lane * 4 + 3]; }
                    if (gx2) gx2[t0 + lane] = *reinterpret_cast<const float2 *>(sdf + lane * 2);
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
                        prt[L.oWih + (g * H + lane) * 2] = gwih[g * 2];
                        prt[L.oWih + (g * H + lane) * 2 + 1] = gwih[g * 2 + 1];
                        prt[L.obih + g * H + lane] = gb[g];
                        prt[L.obhh + g * H + lane] = gb[g];
#pragma unroll
                        for (int k = 0; k < HT; ++k)
                            if (k < H) prt[L.oWhh + (g * H + k) * H + lane] = gwhh[g * HT + k];
                    }
                    prt[L.oWo + lane] = gwo0; prt[L.oWo + H + lane] = gwo1;
                }
                gbo0 = warp_sum(gbo0); gbo1 = warp_sum(gbo1);
                if (lane == 0) { prt[L.obo] = gbo0; prt[L.obo + 1] = gbo1; }
            }
        }
    }
}

#define ODPD_LSTM_TIERS(X) X(9) X(16) X(32)
static int lstm_tier(int H) {
#define X(HTV) if (H <= HTV) return HTV;
    ODPD_LSTM_TIERS(X)
#undef X
    return -1;
}
// dir: 0 fwd, 1 bwd, +2 = plan only
template <int HT> static int lstm_launch(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *info) {
    constexpr int HP = Pad4<HT>::value, ROW = LRow<HT>::value;
    const LstmLayout L(a.H);
    const int Ppad = (L.P + 3) & ~3;
    const bool plan_only = dir >= 2;
    if ((dir & 1) == 0) {
        const size_t smem = (size_t)LFwdSmem<HT>::total(Ppad) * 4;
        static OccCache occ{};
        const int64_t soff = a.save ? (int64_t)a.B * a.T * ROW : 0;
        return chunk_launch(lstm_fwd_kernel<HT>, 96, smem, &occ, a, 0, a.saved ? a.saved + soff : nullptr, soff, 2 * HP, st, plan_only, info,
                            "lstm_fwd_kernel");
    }
    const size_t smem = (size_t)LBwdSmem<HT>::total(Ppad) * 4;
    const int64_t woff = (chunk_rows(a.B, a.tchunks_req) * L.P + 3) & ~(int64_t)3;
    float *scr = a.partials ? a.partials + woff : nullptr;
    if (dw) {
        static OccCache occ{};
        return chunk_launch(lstm_bwd_kernel<HT, true>, 96, smem, &occ, a, 1, scr, woff, 2 * HP, st, plan_only, info, "lstm_bwd_kernel");
    }
    static OccCache occ0{};
    return chunk_launch(lstm_bwd_kernel<HT, false>, 96, smem, &occ0, a, 1, scr, woff, 2 * HP, st, plan_only, info, "lstm_bwd_kernel");
}
int64_t lstm_saved_floats(int B, int T, int H, bool save, int tchunks_req) {
    const int ht = lstm_tier(H);
    if (ht < 0) return -1;
    const int HP = (ht + 3) & ~3;
    return (save ? (int64_t)B * T * 7 * HP : 0) + chunk_fwd_scratch_floats(chunk_rows(B, tchunks_req), 2 * HP);
}
int64_t lstm_workspace_floats(int B, int H, int64_t P, int tchunks_req) {
    const int ht = lstm_tier(H);
    if (ht < 0) return -1;
    return chunk_workspace_floats(chunk_rows(B, tchunks_req), P, 2 * ((ht + 3) & ~3));
}
int lstm_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *info) {
#define X(HTV) if (a.H <= HTV) return lstm_launch<HTV>(a, dir, dw, st, info);
    ODPD_LSTM_TIERS(X)
#undef X
    set_error("LSTM kernels support hidden_size <= 32 (got %d)", a.H);
    return -1;
}

}  // namespace odpd

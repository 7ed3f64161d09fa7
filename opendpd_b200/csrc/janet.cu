// janet.cu — fused forward / backward of the JANET-style backbones.
//
// Replaces (reference, file:line):
//   PGJANET   backbones/pgjanet.py:24-77   a=|x|, th=atan2(q,i); a_n,p1,p2 = tanh(W[h;a|cos|sin]); u = a_n p1 p2 (1-a_n)(1-p1)(1-p2);
//                                          f = sigmoid(W_f[h;u]); g = tanh(W_g[h;u]); h = f h + (1-f) g; y = W_o h + b
//   DVRJANET  backbones/dvrjanet.py:43-102 (dvr_block :32-41): s = h_I + h_Q; th~ = W_pth th + W_ph s;
//                                          a~ = sum_k c_k |W_ax a + W_ah s - k/K|; f = sigmoid(W_f s + b);
//                                          g_c = tanh(W_ccos[h_I; a~ cos th~]); g_s = tanh(W_csin[h_Q; a~ sin th~]);
//                                          h_I = f h_I + (1-f) g_c; h_Q = f h_Q + (1-f) g_s; y = (W_o1 h_I, W_o2 h_Q)
// Same 3-warp chunk pipeline as gru_family.cu; the chain warp needs two shared-memory broadcasts per step (h, then u / a~cos,a~sin).
// atan2f / sinf / cosf are the full-range libdevice versions (the learned phase th~ is unbounded, SURVEY §7 hard part 2).
#include "cells.h"
#include "pipeline.cuh"
#include "chunking.cuh"

namespace odpd {

// =====================================================================================================================
//                                                       PGJANET
// =====================================================================================================================
struct PgLayout {
    int H, oWa, oba, oWp1, obp1, oWp2, obp2, oWf, obf, oWg, obg, oWo, obo, P;
    __host__ __device__ explicit PgLayout(int h) {
        H = h; const int h1 = h * (h + 1), h2 = 2 * h * h;
        oWa = 0; oba = h1; oWp1 = oba + h; obp1 = oWp1 + h1; oWp2 = obp1 + h; obp2 = oWp2 + h1; oWf = obp2 + h; obf = oWf + h2;
        oWg = obf + h; obg = oWg + h2; oWo = obg + h; obo = oWo + 2 * h; P = obo + 2;
    }
};
// row: a_n | p1 | p2 | u | f | g | h_t
template <int HT> struct PgRow { static constexpr int value = 7 * Pad4<HT>::value; };
template <int HT> struct PgFwdSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = PgRow<HT>::value, XP = CH * 3 * HP, ACT = CH * ROW, PO = CH * 33;
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + HP + 2 * XP + 2 * ACT + 2 * PO + 2 * HP; }
};
template <int HT> struct PgBwdSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = PgRow<HT>::value, ACT = (CH + 1) * ROW, PRE = CH * 8, DH = CH * HP, G = CH * 5 * HP, DF = CH * 4;
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + 3 * ACT + 3 * PRE + 2 * DH + 2 * G + DF + 4 * HP; }
};

__device__ __forceinline__ void pg_features(float i, float q, float &a, float &c, float &s) {
    a = __fsqrt_rn(__fadd_rn(__fmul_rn(i, i), __fmul_rn(q, q)));
    const float th = atan2f(q, i);
    c = cosf(th);
    s = sinf(th);
}

template <int HT>
__global__ void __launch_bounds__(96, 1) pgjanet_fwd_kernel(GruArgs a) {
    pdl_enter();   // launched through chunk_launch with the programmatic-serialization attribute (common.cuh)
    constexpr int HP = Pad4<HT>::value, ROW = PgRow<HT>::value;
    using SM = PgFwdSmem<HT>;
    const PgLayout L(a.H);
    const int H = a.H, T = a.T, H1 = H + 1, H2 = 2 * H;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *zero = sp + Ppad;
    float *sxp = zero + HP;                  // [2][CH][3HP]
    float *sact = sxp + 2 * SM::XP;          // [2][CH][ROW]
    float *spo = sact + 2 * SM::ACT;
    float *sul = spo + 2 * SM::PO;           // [2][HP] u broadcast line
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const FwdRange R = fwd_range(a);          // chunking.cuh; recurrent state per chunk = h, HP floats
    const bool spec = R.spec;
    const int b = R.b, t_emit = R.t_emit, t_hi = R.t_hi;
    if (a.mode == 2 && fwd_verify_pass(a, b, HP, HP, H)) return;
    stage_params(sp, a.params, L.P, bars);
    if (threadIdx.x < HP) zero[threadIdx.x] = 0.f;
    __syncthreads();
    const bool act = lane < H;
    const int j = act ? lane : 0, lp = lane < HP ? lane : 0;
    const int cb = R.t_lo / CH, nchunks = (t_hi + CH - 1) / CH - cb;
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);

    if (warp == 1) {
        const float wa = act ? sp[L.oWa + j * H1 + H] : 0.f, w1 = act ? sp[L.oWp1 + j * H1 + H] : 0.f, w2 = act ? sp[L.oWp2 + j * H1 + H] : 0.f;
        const float ba = act ? sp[L.oba + j] : 0.f, b1 = act ? sp[L.obp1 + j] : 0.f, b2 = act ? sp[L.obp2 + j] : 0.f;
        for (int s = 0; s < nchunks + 2; ++s) {
            if (s < nchunks) {
                const int t0 = (cb + s) * CH, nt = min(CH, t_hi - t0);
                float *xp = sxp + (s & 1) * SM::XP;
                float fa = 0.f, fc = 0.f, fs = 0.f;
                if (lane < nt) { const float2 v = __ldg(x2 + t0 + lane); pg_features(v.x, v.y, fa, fc, fs); }
                for (int tl = 0; tl < nt; ++tl) {
                    const float av = __shfl_sync(ODPD_FULL, fa, tl), cv = __shfl_sync(ODPD_FULL, fc, tl), sv = __shfl_sync(ODPD_FULL, fs, tl);
                    if (lane < HP) { float *o = xp + tl * 3 * HP + lane; o[0] = fmaf(wa, av, ba); o[HP] = fmaf(w1, cv, b1); o[2 * HP] = fmaf(w2, sv, b2); }
                }
            }
            __syncthreads();
        }
    } else if (warp == 0) {
        float wa[HT], w1[HT], w2[HT], wfh[HT], wfu[HT], wgh[HT], wgu[HT];
#pragma unroll
        for (int k = 0; k < HT; ++k) {
            const bool ok = act && k < H;
            wa[k] = ok ? sp[L.oWa + j * H1 + k] : 0.f;  w1[k] = ok ? sp[L.oWp1 + j * H1 + k] : 0.f;  w2[k] = ok ? sp[L.oWp2 + j * H1 + k] : 0.f;
            wfh[k] = ok ? sp[L.oWf + j * H2 + k] : 0.f; wfu[k] = ok ? sp[L.oWf + j * H2 + H + k] : 0.f;
            wgh[k] = ok ? sp[L.oWg + j * H2 + k] : 0.f; wgu[k] = ok ? sp[L.oWg + j * H2 + H + k] : 0.f;
        }
        const float bf = act ? sp[L.obf + j] : 0.f, bg = act ? sp[L.obg + j] : 0.f;
        float h = 0.f;
        int cur = 0;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int c = s - 1;
            if (c >= 0 && c < nchunks) {
                const int t0 = (cb + c) * CH, nt = min(CH, t_hi - t0);
                if (spec && R.cc > 0 && t0 == t_emit && lane < HP) a.sc_guess[(size_t)blockIdx.x * HP + lane] = h;
                const float *xp = sxp + (c & 1) * SM::XP + lp;
                float *ac = sact + (c & 1) * SM::ACT;
                const float *hrow = (c == 0) ? zero : sact + ((c - 1) & 1) * SM::ACT + (CH - 1) * ROW + 6 * HP;
                float xa = xp[0], x1 = xp[HP], x2v = xp[2 * HP];
                for (int tl = 0; tl < nt; ++tl) {
                    const int tn = (tl + 1 < nt) ? tl + 1 : tl;
                    const float nxa = xp[tn * 3 * HP], nx1 = xp[tn * 3 * HP + HP], nx2 = xp[tn * 3 * HP + 2 * HP];
                    float a0 = xa, a1 = 0.f, p10 = x1, p11 = 0.f, p20 = x2v, p21 = 0.f, f0 = bf, f1 = 0.f, g0 = bg, g1 = 0.f;
                    bcast_dot<HT>(hrow, wa, a0, a1);
                    bcast_dot<HT>(hrow, w1, p10, p11);
                    bcast_dot<HT>(hrow, w2, p20, p21);
                    bcast_dot<HT>(hrow, wfh, f0, f1);
                    bcast_dot<HT>(hrow, wgh, g0, g1);
                    const float an = tanhf_(a0 + a1), p1 = tanhf_(p10 + p11), p2 = tanhf_(p20 + p21);
                    const float u = an * p1 * p2 * (1.f - an) * (1.f - p1) * (1.f - p2);
                    float *ul = sul + cur * HP;
                    if (lane < HP) ul[lane] = act ? u : 0.f;
                    __syncwarp();
                    bcast_dot<HT>(ul, wfu, f0, f1);
                    bcast_dot<HT>(ul, wgu, g0, g1);
                    const float f = sigmoidf_(f0 + f1), g = tanhf_(g0 + g1);
                    h = fmaf(f, h, (1.f - f) * g);
                    float *row = ac + tl * ROW;
                    if (lane < HP) {
                        row[6 * HP + lane] = h;
                        row[lane] = an; row[HP + lane] = p1; row[2 * HP + lane] = p2; row[3 * HP + lane] = u; row[4 * HP + lane] = f; row[5 * HP + lane] = g;
                    }
                    hrow = row + 6 * HP;
                    cur ^= 1;
                    xa = nxa; x1 = nx1; x2v = nx2;
                    __syncwarp();
                }
                fence_async_smem();
            }
            __syncthreads();
        }
        if (spec && lane < HP) a.sc_end[(size_t)blockIdx.x * HP + lane] = h;
    } else {
        const float wo0 = act ? sp[L.oWo + j] : 0.f, wo1 = act ? sp[L.oWo + H + j] : 0.f;
        const float bo0 = sp[L.obo], bo1 = sp[L.obo + 1];
        const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
        float2 *o2 = reinterpret_cast<float2 *>(a.out) + (size_t)b * T;
        float *svg = a.save ? a.saved + (size_t)b * T * ROW : nullptr;
        float lsum = 0.f;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int c = s - 2;
            if (c >= 0 && (cb + c) * CH >= t_emit) {       // warm-up blocks emit nothing
                const int t0 = (cb + c) * CH, nt = min(CH, t_hi - t0);
                float *ac = sact + (c & 1) * SM::ACT;
                if (svg && lane == 0) tma_store_1d(svg + (size_t)t0 * ROW, ac, (uint32_t)(nt * ROW * 4));
                linear_head_chunk(ac, ROW, 6 * HP, HP, H, nt, lane, wo0, wo1, bo0, bo1, spo, nullptr, o2 + t0, y2 ? y2 + t0 : iq_none(), lsum);
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
__global__ void __launch_bounds__(96, 1) pgjanet_bwd_kernel(GruArgs a) {
    pdl_enter();   // launched through chunk_launch with the programmatic-serialization attribute (common.cuh)
    constexpr int HP = Pad4<HT>::value, ROW = PgRow<HT>::value;
    using SM = PgBwdSmem<HT>;
    const PgLayout L(a.H);
    const int H = a.H, T = a.T, H1 = H + 1, H2 = 2 * H;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *sact = sp + Ppad;                 // [3][CH+1][ROW]
    float *spre = sact + 3 * SM::ACT;        // [3][CH][8]: i q a cos sin go0 go1 -
    float *sdh = spre + 3 * SM::PRE;         // [2][CH][HP]
    float *sG = sdh + 2 * SM::DH;            // [2][CH][5HP]: af | ag | aa | a1 | a2
    float *sdf = sG + 2 * SM::G;             // [CH][4]: ga gc gs
    float *sl = sdf + SM::DF;                // [2][2HP] chain broadcast lines for (af,ag) — also stored in sG; kept separate for double buffering
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const BwdRange R = bwd_range(a);          // chunking.cuh; adjoint state per chunk: HP floats
    const bool spec = R.spec;
    const int b = R.b, t_ehi = R.t_ehi, t_hi = R.t_hi;
    if (a.mode == 2 && bwd_verify_pass(a, b, HP, HP, H)) return;
    if (threadIdx.x == 0) { mbar_init(bars + 1, 1); mbar_init(bars + 2, 1); mbar_init(bars + 3, 1); }
    stage_params(sp, a.params, L.P, bars);
    const bool act = lane < H;
    const int j = act ? lane : 0, lp = lane < HP ? lane : 0;
    // 32-step blocks [cb, ce) are processed last to first; blocks >= ce_emit are warm-up
    const int cb = R.t_elo / CH, ce = (t_hi + CH - 1) / CH, nchunks = ce - cb, ce_emit = (t_ehi + CH - 1) / CH;
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
    const float *svg = a.saved + (size_t)b * T * ROW;
    (void)sl;

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
                    float fa, fc, fs;
                    pg_features(v.x, v.y, fa, fc, fs);
                    const float2 g = load_gout(go2, oi2, y2, t0 + lane, gs);
                    float4 *d = reinterpret_cast<float4 *>(pr + lane * 8);
                    d[0] = make_float4(v.x, v.y, fa, fc);
                    d[1] = make_float4(fs, g.x, g.y, 0.f);
                }
                __syncwarp();
                if (lane < HP)
                    for (int tl = 0; tl < nt; ++tl) dh[tl * HP + lane] = fmaf(wo0, pr[tl * 8 + 5], wo1 * pr[tl * 8 + 6]);
                mbar_wait(bar, (uint32_t)((s / 3) & 1));
            }
            __syncthreads();
        }
    } else if (warp == 0) {
        // weight columns j: h-part of W_f, W_g, W_a, W_p1, W_p2 and u-part of W_f, W_g
        float cfh[HT], cgh[HT], cfu[HT], cgu[HT], ca[HT], c1[HT], c2[HT];
#pragma unroll
        for (int k = 0; k < HT; ++k) {
            const bool ok = act && k < H;
            cfh[k] = ok ? sp[L.oWf + k * H2 + j] : 0.f; cfu[k] = ok ? sp[L.oWf + k * H2 + H + j] : 0.f;
            cgh[k] = ok ? sp[L.oWg + k * H2 + j] : 0.f; cgu[k] = ok ? sp[L.oWg + k * H2 + H + j] : 0.f;
            ca[k] = ok ? sp[L.oWa + k * H1 + j] : 0.f;  c1[k] = ok ? sp[L.oWp1 + k * H1 + j] : 0.f;  c2[k] = ok ? sp[L.oWp2 + k * H1 + j] : 0.f;
        }
        float gH = 0.f;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 1;
            if (sc >= 0 && sc < nchunks) {
                const int c = ce - 1 - sc, t0 = c * CH, nt = min(CH, t_hi - t0);
                if (spec && c == ce_emit - 1 && t_ehi < T && lane < HP) a.sc_guess[(size_t)blockIdx.x * HP + lane] = gH;
                const float *ac = sact + (sc % 3) * SM::ACT + lp;
                const float *dh = sdh + (sc & 1) * SM::DH + lp;
                float *Gb = sG + (sc & 1) * SM::G;
                const float *row = ac + nt * ROW;
                float an = row[0], p1 = row[HP], p2 = row[2 * HP], f = row[4 * HP], g = row[5 * HP], hp = row[6 * HP - ROW], dht = dh[(nt - 1) * HP];
                for (int tl = nt - 1; tl >= 0; --tl) {
                    const int tp = tl > 0 ? tl - 1 : 0;
                    const float *rn = ac + (tp + 1) * ROW;
                    const float an_n = rn[0], p1_n = rn[HP], p2_n = rn[2 * HP], f_n = rn[4 * HP], g_n = rn[5 * HP], hp_n = rn[6 * HP - ROW], dh_n = dh[tp * HP];
                    gH += dht;
                    const float af = gH * (hp - g) * f * (1.f - f);
                    const float ag = gH * (1.f - f) * (1.f - g * g);
                    float ghp = gH * f;
                    float *G = Gb + tl * 5 * HP;
                    if (lane < HP) { G[lane] = act ? af : 0.f; G[HP + lane] = act ? ag : 0.f; }
                    __syncwarp();
                    float u0 = 0.f, u1 = 0.f, h0 = 0.f, h1 = 0.f;
                    bcast_dot<HT>(G, cfu, u0, u1);
                    bcast_dot<HT>(G + HP, cgu, u0, u1);
                    bcast_dot<HT>(G, cfh, h0, h1);
                    bcast_dot<HT>(G + HP, cgh, h0, h1);
                    const float gu = u0 + u1;
                    const float A = an * (1.f - an), P1 = p1 * (1.f - p1), P2 = p2 * (1.f - p2);
                    const float aa = gu * (1.f - 2.f * an) * P1 * P2 * (1.f - an * an);
                    const float a1 = gu * A * (1.f - 2.f * p1) * P2 * (1.f - p1 * p1);
                    const float a2 = gu * A * P1 * (1.f - 2.f * p2) * (1.f - p2 * p2);
                    if (lane < HP) { G[2 * HP + lane] = act ? aa : 0.f; G[3 * HP + lane] = act ? a1 : 0.f; G[4 * HP + lane] = act ? a2 : 0.f; }
                    __syncwarp();
                    bcast_dot<HT>(G + 2 * HP, ca, h0, h1);
                    bcast_dot<HT>(G + 3 * HP, c1, h0, h1);
                    bcast_dot<HT>(G + 4 * HP, c2, h0, h1);
                    gH = ghp + (h0 + h1);
                    an = an_n; p1 = p1_n; p2 = p2_n; f = f_n; g = g_n; hp = hp_n; dht = dh_n;
                }
            }
            __syncthreads();
        }
        if (spec && lane < HP) a.sc_end[(size_t)blockIdx.x * HP + lane] = gH;
    } else {
        const int fl = lane - H;                   // feature lanes H..H+2: d/da, d/dcos, d/dsin
        const bool isf = fl >= 0 && fl < 3;
        float wx[HT];                              // column H of W_a / W_p1 / W_p2 for the feature lanes
#pragma unroll
        for (int k = 0; k < HT; ++k) {
            float w = 0.f;
            if (isf && k < H) w = sp[(fl == 0 ? L.oWa : (fl == 1 ? L.oWp1 : L.oWp2)) + k * H1 + H];
            wx[k] = w;
        }
        float gfh[DW ? HT : 1], gfu[DW ? HT : 1], ggh[DW ? HT : 1], ggu[DW ? HT : 1], gah[DW ? HT : 1], g1h[DW ? HT : 1], g2h[DW ? HT : 1];
        if constexpr (DW) {
#pragma unroll
            for (int k = 0; k < HT; ++k) { gfh[k] = gfu[k] = ggh[k] = ggu[k] = gah[k] = g1h[k] = g2h[k] = 0.f; }
        }
        float gbf = 0.f, gbg = 0.f, gba = 0.f, gb1 = 0.f, gb2 = 0.f, gxa = 0.f, gx1 = 0.f, gx2v = 0.f, gwo0 = 0.f, gwo1 = 0.f, gbo0 = 0.f, gbo1 = 0.f;
        float2 *gx2 = (a.need_dx && a.gx) ? reinterpret_cast<float2 *>(a.gx) + (size_t)b * T : nullptr;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 2;
            if (sc >= 0 && ce - 1 - sc < ce_emit) {          // warm-up blocks emit nothing
                const int c = ce - 1 - sc, t0 = c * CH, nt = min(CH, t_hi - t0);
                const float *ac = sact + (sc % 3) * SM::ACT, *pr = spre + (sc % 3) * SM::PRE, *Gb = sG + (sc & 1) * SM::G;
                for (int tl = 0; tl < nt; ++tl) {
                    const float *G = Gb + tl * 5 * HP;
                    const float *row = ac + (tl + 1) * ROW;
                    const float hp = row[6 * HP - ROW + lp], ht = row[6 * HP + lp], u = row[3 * HP + lp];
                    const float4 p0 = *reinterpret_cast<const float4 *>(pr + tl * 8), p1v = *reinterpret_cast<const float4 *>(pr + tl * 8 + 4);
                    if constexpr (DW) {
                        const float af = G[lp], ag = G[HP + lp], aa = G[2 * HP + lp], a1 = G[3 * HP + lp], a2 = G[4 * HP + lp];
                        gbf += af; gbg += ag; gba += aa; gb1 += a1; gb2 += a2;
                        gxa = fmaf(aa, p0.z, gxa); gx1 = fmaf(a1, p0.w, gx1); gx2v = fmaf(a2, p1v.x, gx2v);
                        gwo0 = fmaf(p1v.y, ht, gwo0); gwo1 = fmaf(p1v.z, ht, gwo1);
                        const float4 *G4 = reinterpret_cast<const float4 *>(G);
#pragma unroll
                        for (int k4 = 0; k4 < HP / 4; ++k4) {
                            const float4 vf = G4[k4], vg = G4[HP / 4 + k4], va = G4[2 * (HP / 4) + k4], v1 = G4[3 * (HP / 4) + k4], v2 = G4[4 * (HP / 4) + k4];
                            const float ef[4] = {vf.x, vf.y, vf.z, vf.w}, eg[4] = {vg.x, vg.y, vg.z, vg.w}, ea[4] = {va.x, va.y, va.z, va.w};
                            const float e1[4] = {v1.x, v1.y, v1.z, v1.w}, e2[4] = {v2.x, v2.y, v2.z, v2.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int k = k4 * 4 + e;
                                if (k < HT) {
                                    gfh[k] = fmaf(ef[e], hp, gfh[k]); gfu[k] = fmaf(ef[e], u, gfu[k]);
                                    ggh[k] = fmaf(eg[e], hp, ggh[k]); ggu[k] = fmaf(eg[e], u, ggu[k]);
                                    gah[k] = fmaf(ea[e], hp, gah[k]); g1h[k] = fmaf(e1[e], hp, g1h[k]); g2h[k] = fmaf(e2[e], hp, g2h[k]);
                                }
                            }
                        }
                    }
                    if (a.need_dx && isf) {
                        const float *line = G + (2 + fl) * HP;
                        float d0 = 0.f, d1 = 0.f;
#pragma unroll
                        for (int k = 0; k < HT; ++k) { if (k & 1) d1 = fmaf(wx[k], line[k], d1); else d0 = fmaf(wx[k], line[k], d0); }
                        sdf[tl * 4 + fl] = d0 + d1;
                    }
                }
                __syncwarp();
                if (lane < nt) {
                    const float *p = pr + lane * 8;
                    if constexpr (DW) { gbo0 += p[5]; gbo1 += p[6]; }
                    if (gx2) {
                        const float i = p[0], q = p[1], am = p[2], cs = p[3], sn = p[4];
                        const float ga = sdf[lane * 4], gc = sdf[lane * 4 + 1], gsn = sdf[lane * 4 + 2];
                        const float gth = fmaf(-sn, gc, cs * gsn), a2 = am * am;
                        gx2[t0 + lane] = make_float2(ga * i / am - gth * q / a2, ga * q / am + gth * i / a2);
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
                    for (int k = 0; k < HT; ++k)
                        if (k < H) {
                            prt[L.oWf + k * H2 + lane] = gfh[k]; prt[L.oWf + k * H2 + H + lane] = gfu[k];
                            prt[L.oWg + k * H2 + lane] = ggh[k]; prt[L.oWg + k * H2 + H + lane] = ggu[k];
                            prt[L.oWa + k * H1 + lane] = gah[k]; prt[L.oWp1 + k * H1 + lane] = g1h[k]; prt[L.oWp2 + k * H1 + lane] = g2h[k];
                        }
                    prt[L.oWa + lane * H1 + H] = gxa; prt[L.oWp1 + lane * H1 + H] = gx1; prt[L.oWp2 + lane * H1 + H] = gx2v;
                    prt[L.oba + lane] = gba; prt[L.obp1 + lane] = gb1; prt[L.obp2 + lane] = gb2; prt[L.obf + lane] = gbf; prt[L.obg + lane] = gbg;
                    prt[L.oWo + lane] = gwo0; prt[L.oWo + H + lane] = gwo1;
                }
                gbo0 = warp_sum(gbo0); gbo1 = warp_sum(gbo1);
                if (lane == 0) { prt[L.obo] = gbo0; prt[L.obo + 1] = gbo1; }
            }
        }
    }
}

// =====================================================================================================================
//                                                       DVRJANET
// =====================================================================================================================
struct DvLayout {
    int H, K, ocs, oWph, oWpt, oWah, oWax, oWf, obf, oWc, obc, oWs, obs, oWo1, obo1, oWo2, obo2, P;
    __host__ __device__ DvLayout(int h, int k) {
        H = h; K = k; const int hh = h * h;
        ocs = 0; oWph = k; oWpt = oWph + hh; oWah = oWpt + h; oWax = oWah + hh; oWf = oWax + h; obf = oWf + hh; oWc = obf + h; obc = oWc + 2 * hh;
        oWs = obc + h; obs = oWs + 2 * hh; oWo1 = obs + h; obo1 = oWo1 + h; oWo2 = obo1 + 1; obo2 = oWo2 + h; P = obo2 + 1;
    }
};
// row: pa | a~ | cos | sin | f | g_c | g_s | h_I | h_Q
template <int HT> struct DvRow { static constexpr int value = 9 * Pad4<HT>::value; };
template <int HT> struct DvFwdSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = DvRow<HT>::value, XP = CH * 2 * HP, ACT = CH * ROW, PO = CH * 33;
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + ROW + 2 * XP + 2 * ACT + 2 * PO + 2 * 3 * HP; }
};
template <int HT> struct DvBwdSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = DvRow<HT>::value, ACT = (CH + 1) * ROW, PRE = CH * 8, DH = CH * 2 * HP, G = CH * 6 * HP, DF = CH * 2;
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + 3 * ACT + 3 * PRE + 2 * DH + 2 * G + DF; }
};

template <int HT>
__global__ void __launch_bounds__(96, 1) dvrjanet_fwd_kernel(GruArgs a) {
    pdl_enter();   // launched through chunk_launch with the programmatic-serialization attribute (common.cuh)
    constexpr int HP = Pad4<HT>::value, ROW = DvRow<HT>::value;
    using SM = DvFwdSmem<HT>;
    const DvLayout L(a.H, a.K);
    const int H = a.H, T = a.T, K = a.K, H2 = 2 * H;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *zero = sp + Ppad;                 // [ROW]
    float *sxp = zero + ROW;                 // [2][CH][2HP]
    float *sact = sxp + 2 * SM::XP;          // [2][CH][ROW]
    float *spo = sact + 2 * SM::ACT;
    float *sln = spo + 2 * SM::PO;           // [2][3HP]: s | a~cos | a~sin lines
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const FwdRange R = fwd_range(a);          // chunking.cuh; recurrent state per chunk = (h_I, h_Q), 2*HP floats
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
        const float wpt = act ? sp[L.oWpt + j] : 0.f, wax = act ? sp[L.oWax + j] : 0.f;
        for (int s = 0; s < nchunks + 2; ++s) {
            if (s < nchunks) {
                const int t0 = (cb + s) * CH, nt = min(CH, t_hi - t0);
                float *xp = sxp + (s & 1) * SM::XP;
                float fa = 0.f, fth = 0.f;
                if (lane < nt) {
                    const float2 v = __ldg(x2 + t0 + lane);
                    fa = __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));
                    fth = atan2f(v.y, v.x);
                }
                for (int tl = 0; tl < nt; ++tl) {
                    const float av = __shfl_sync(ODPD_FULL, fa, tl), tv = __shfl_sync(ODPD_FULL, fth, tl);
                    if (lane < HP) { xp[tl * 2 * HP + lane] = wpt * tv; xp[tl * 2 * HP + HP + lane] = wax * av; }
                }
            }
            __syncthreads();
        }
    } else if (warp == 0) {
        float wph[HT], wah[HT], wf[HT], wch[HT], wcv[HT], wsh[HT], wsv[HT];
#pragma unroll
        for (int k = 0; k < HT; ++k) {
            const bool ok = act && k < H;
            wph[k] = ok ? sp[L.oWph + j * H + k] : 0.f; wah[k] = ok ? sp[L.oWah + j * H + k] : 0.f; wf[k] = ok ? sp[L.oWf + j * H + k] : 0.f;
            wch[k] = ok ? sp[L.oWc + j * H2 + k] : 0.f; wcv[k] = ok ? sp[L.oWc + j * H2 + H + k] : 0.f;
            wsh[k] = ok ? sp[L.oWs + j * H2 + k] : 0.f; wsv[k] = ok ? sp[L.oWs + j * H2 + H + k] : 0.f;
        }
        const float bf = act ? sp[L.obf + j] : 0.f, bc = act ? sp[L.obc + j] : 0.f, bs = act ? sp[L.obs + j] : 0.f;
        // DVR knots k/K (rounded once from double, as the reference's python float -> tensor dtype conversion) and coefficients
        float knot[8], ck[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { knot[k] = (float)((double)(k + 1) / (double)K); ck[k] = k < K ? sp[L.ocs + k] : 0.f; }
        float hI = 0.f, hQ = 0.f;
        int cur = 0;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int c = s - 1;
            if (c >= 0 && c < nchunks) {
                const int t0 = (cb + c) * CH, nt = min(CH, t_hi - t0);
                if (spec && R.cc > 0 && t0 == t_emit && lane < HP) {
                    a.sc_guess[(size_t)blockIdx.x * 2 * HP + lane] = hI;
                    a.sc_guess[(size_t)blockIdx.x * 2 * HP + HP + lane] = hQ;
                }
                const float *xp = sxp + (c & 1) * SM::XP + lp;
                float *ac = sact + (c & 1) * SM::ACT;
                const float *prow = (c == 0) ? zero : sact + ((c - 1) & 1) * SM::ACT + (CH - 1) * ROW;
                float xt = xp[0], xa = xp[HP];
                for (int tl = 0; tl < nt; ++tl) {
                    const int tn = (tl + 1 < nt) ? tl + 1 : tl;
                    const float nxt = xp[tn * 2 * HP], nxa = xp[tn * 2 * HP + HP];
                    float *ln = sln + cur * 3 * HP;
                    if (lane < HP) ln[lane] = act ? hI + hQ : 0.f;
                    __syncwarp();
                    float t0a = xt, t1a = 0.f, p0 = xa, p1 = 0.f, f0 = bf, f1 = 0.f, gc0 = bc, gc1 = 0.f, gs0 = bs, gs1 = 0.f;
                    bcast_dot<HT>(ln, wph, t0a, t1a);
                    bcast_dot<HT>(ln, wah, p0, p1);
                    bcast_dot<HT>(ln, wf, f0, f1);
                    bcast_dot<HT>(prow + 7 * HP, wch, gc0, gc1);
                    bcast_dot<HT>(prow + 8 * HP, wsh, gs0, gs1);
                    const float tht = t0a + t1a, pa = p0 + p1;
                    float at = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) at = fmaf(fabsf(pa - knot[k]), ck[k], at);   // ck = 0 beyond K
                    float st, ct;
                    sincosf(tht, &st, &ct);
                    const float f = sigmoidf_(f0 + f1);
                    if (lane < HP) { ln[HP + lane] = act ? at * ct : 0.f; ln[2 * HP + lane] = act ? at * st : 0.f; }
                    __syncwarp();
                    bcast_dot<HT>(ln + HP, wcv, gc0, gc1);
                    bcast_dot<HT>(ln + 2 * HP, wsv, gs0, gs1);
                    const float gc = tanhf_(gc0 + gc1), gsv = tanhf_(gs0 + gs1);
                    hI = fmaf(f, hI, (1.f - f) * gc);
                    hQ = fmaf(f, hQ, (1.f - f) * gsv);
                    float *row = ac + tl * ROW;
                    if (lane < HP) {
                        row[7 * HP + lane] = hI; row[8 * HP + lane] = hQ;
                        row[lane] = pa; row[HP + lane] = at; row[2 * HP + lane] = ct; row[3 * HP + lane] = st; row[4 * HP + lane] = f;
                        row[5 * HP + lane] = gc; row[6 * HP + lane] = gsv;
                    }
                    prow = row;
                    cur ^= 1;
                    xt = nxt; xa = nxa;
                }
                __syncwarp();
                fence_async_smem();
            }
            __syncthreads();
        }
        if (spec && lane < HP) {
            a.sc_end[(size_t)blockIdx.x * 2 * HP + lane] = hI;
            a.sc_end[(size_t)blockIdx.x * 2 * HP + HP + lane] = hQ;
        }
    } else {
        const float wo0 = act ? sp[L.oWo1 + j] : 0.f, wo1 = act ? sp[L.oWo2 + j] : 0.f;
        const float bo0 = sp[L.obo1], bo1 = sp[L.obo2];
        const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
        float2 *o2 = reinterpret_cast<float2 *>(a.out) + (size_t)b * T;
        float *svg = a.save ? a.saved + (size_t)b * T * ROW : nullptr;
        float lsum = 0.f;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int c = s - 2;
            if (c >= 0 && (cb + c) * CH >= t_emit) {       // warm-up blocks emit nothing
                const int t0 = (cb + c) * CH, nt = min(CH, t_hi - t0);
                float *ac = sact + (c & 1) * SM::ACT;
                if (svg && lane == 0) tma_store_1d(svg + (size_t)t0 * ROW, ac, (uint32_t)(nt * ROW * 4));
                linear_head_chunk2(ac, ROW, 7 * HP, 8 * HP, HP, H, nt, lane, wo0, wo1, bo0, bo1, spo, nullptr, o2 + t0, y2 ? y2 + t0 : iq_none(), lsum);
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
__global__ void __launch_bounds__(96, 1) dvrjanet_bwd_kernel(GruArgs a) {
    pdl_enter();   // launched through chunk_launch with the programmatic-serialization attribute (common.cuh)
    constexpr int HP = Pad4<HT>::value, ROW = DvRow<HT>::value;
    using SM = DvBwdSmem<HT>;
    const DvLayout L(a.H, a.K);
    const int H = a.H, T = a.T, K = a.K, H2 = 2 * H;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *sact = sp + Ppad;                 // [3][CH+1][ROW]
    float *spre = sact + 3 * SM::ACT;        // [3][CH][8]: i q a th go0 go1
    float *sdh = spre + 3 * SM::PRE;         // [2][CH][2HP]: dL/dh_I | dL/dh_Q from the head
    float *sG = sdh + 2 * SM::DH;            // [2][CH][6HP]: ac | as | af | gtht | gpa | gat
    float *sdf = sG + 2 * SM::G;             // [CH][2]: gth, ga
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const BwdRange R = bwd_range(a);          // chunking.cuh; adjoint state per chunk: 2 * HP floats
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
        const float wo0 = act ? sp[L.oWo1 + j] : 0.f, wo1 = act ? sp[L.oWo2 + j] : 0.f;
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
                    const float am = __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));
                    const float th = atan2f(v.y, v.x);
                    const float2 g = load_gout(go2, oi2, y2, t0 + lane, gs);
                    float4 *d = reinterpret_cast<float4 *>(pr + lane * 8);
                    d[0] = make_float4(v.x, v.y, am, th);
                    d[1] = make_float4(g.x, g.y, 0.f, 0.f);
                }
                __syncwarp();
                if (lane < HP)
                    for (int tl = 0; tl < nt; ++tl) { dh[tl * 2 * HP + lane] = wo0 * pr[tl * 8 + 4]; dh[tl * 2 * HP + HP + lane] = wo1 * pr[tl * 8 + 5]; }
                mbar_wait(bar, (uint32_t)((s / 3) & 1));
            }
            __syncthreads();
        }
    } else if (warp == 0) {
        float cch[HT], ccv[HT], csh[HT], csv[HT], cph[HT], cah[HT], cf[HT];   // weight columns j
#pragma unroll
        for (int k = 0; k < HT; ++k) {
            const bool ok = act && k < H;
            cch[k] = ok ? sp[L.oWc + k * H2 + j] : 0.f; ccv[k] = ok ? sp[L.oWc + k * H2 + H + j] : 0.f;
            csh[k] = ok ? sp[L.oWs + k * H2 + j] : 0.f; csv[k] = ok ? sp[L.oWs + k * H2 + H + j] : 0.f;
            cph[k] = ok ? sp[L.oWph + k * H + j] : 0.f; cah[k] = ok ? sp[L.oWah + k * H + j] : 0.f; cf[k] = ok ? sp[L.oWf + k * H + j] : 0.f;
        }
        float knot[8], ck[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { knot[k] = (float)((double)(k + 1) / (double)K); ck[k] = k < K ? sp[L.ocs + k] : 0.f; }
        float gI = 0.f, gQ = 0.f;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 1;
            if (sc >= 0 && sc < nchunks) {
                const int c = ce - 1 - sc, t0 = c * CH, nt = min(CH, t_hi - t0);
                if (spec && c == ce_emit - 1 && t_ehi < T && lane < HP) {
                    a.sc_guess[(size_t)blockIdx.x * 2 * HP + lane] = gI;
                    a.sc_guess[(size_t)blockIdx.x * 2 * HP + HP + lane] = gQ;
                }
                const float *ac = sact + (sc % 3) * SM::ACT + lp;
                const float *dh = sdh + (sc & 1) * SM::DH + lp;
                float *Gb = sG + (sc & 1) * SM::G;
                for (int tl = nt - 1; tl >= 0; --tl) {
                    const float *row = ac + (tl + 1) * ROW;
                    const float pa = row[0], at = row[HP], ct = row[2 * HP], st = row[3 * HP], f = row[4 * HP], gc = row[5 * HP], gsv = row[6 * HP];
                    const float hI = row[7 * HP - ROW], hQ = row[8 * HP - ROW];
                    gI += dh[tl * 2 * HP]; gQ += dh[tl * 2 * HP + HP];
                    const float af = (gI * (hI - gc) + gQ * (hQ - gsv)) * f * (1.f - f);
                    const float acv = gI * (1.f - f) * (1.f - gc * gc);
                    const float asv = gQ * (1.f - f) * (1.f - gsv * gsv);
                    float gIp = gI * f, gQp = gQ * f;
                    float *G = Gb + tl * 6 * HP;
                    if (lane < HP) { G[lane] = act ? acv : 0.f; G[HP + lane] = act ? asv : 0.f; G[2 * HP + lane] = act ? af : 0.f; }
                    __syncwarp();
                    float i0 = 0.f, i1 = 0.f, q0 = 0.f, q1 = 0.f, vc0 = 0.f, vc1 = 0.f, vs0 = 0.f, vs1 = 0.f;
                    bcast_dot<HT>(G, cch, i0, i1);
                    bcast_dot<HT>(G, ccv, vc0, vc1);
                    bcast_dot<HT>(G + HP, csh, q0, q1);
                    bcast_dot<HT>(G + HP, csv, vs0, vs1);
                    const float gvc = vc0 + vc1, gvs = vs0 + vs1;
                    const float gat = fmaf(gvc, ct, gvs * st);
                    const float gtht = at * fmaf(gvs, ct, -gvc * st);
                    float sg = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float d = pa - knot[k];
                        sg = fmaf(ck[k], d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f), sg);
                    }
                    const float gpa = gat * sg;
                    if (lane < HP) { G[3 * HP + lane] = act ? gtht : 0.f; G[4 * HP + lane] = act ? gpa : 0.f; G[5 * HP + lane] = act ? gat : 0.f; }
                    __syncwarp();
                    float m0 = 0.f, m1 = 0.f;
                    bcast_dot<HT>(G + 3 * HP, cph, m0, m1);
                    bcast_dot<HT>(G + 4 * HP, cah, m0, m1);
                    bcast_dot<HT>(G + 2 * HP, cf, m0, m1);
                    const float gsm = m0 + m1;
                    gI = gIp + (i0 + i1) + gsm;
                    gQ = gQp + (q0 + q1) + gsm;
                }
            }
            __syncthreads();
        }
        if (spec && lane < HP) {
            a.sc_end[(size_t)blockIdx.x * 2 * HP + lane] = gI;
            a.sc_end[(size_t)blockIdx.x * 2 * HP + HP + lane] = gQ;
        }
    } else {
        const int fl = lane - H;                   // feature lanes H (d/dtheta), H+1 (d/da)
        const bool isf = fl >= 0 && fl < 2;
        float wx[HT];
#pragma unroll
        for (int k = 0; k < HT; ++k) wx[k] = (isf && k < H) ? sp[(fl == 0 ? L.oWpt : L.oWax) + k] : 0.f;
        float gch[DW ? HT : 1], gcv[DW ? HT : 1], gsh[DW ? HT : 1], gsvv[DW ? HT : 1], gph[DW ? HT : 1], gah[DW ? HT : 1], gff[DW ? HT : 1];
        if constexpr (DW) {
#pragma unroll
            for (int k = 0; k < HT; ++k) { gch[k] = gcv[k] = gsh[k] = gsvv[k] = gph[k] = gah[k] = gff[k] = 0.f; }
        }
        float gcs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float knot[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) knot[k] = (float)((double)(k + 1) / (double)K);
        float gpt = 0.f, gax = 0.f, gbf = 0.f, gbc = 0.f, gbs = 0.f, gwo0 = 0.f, gwo1 = 0.f, gbo0 = 0.f, gbo1 = 0.f;
        float2 *gx2 = (a.need_dx && a.gx) ? reinterpret_cast<float2 *>(a.gx) + (size_t)b * T : nullptr;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 2;
            if (sc >= 0 && ce - 1 - sc < ce_emit) {          // warm-up blocks emit nothing
                const int c = ce - 1 - sc, t0 = c * CH, nt = min(CH, t_hi - t0);
                const float *ac = sact + (sc % 3) * SM::ACT, *pr = spre + (sc % 3) * SM::PRE, *Gb = sG + (sc & 1) * SM::G;
                for (int tl = 0; tl < nt; ++tl) {
                    const float *G = Gb + tl * 6 * HP;
                    const float *row = ac + (tl + 1) * ROW;
                    const float4 p0 = *reinterpret_cast<const float4 *>(pr + tl * 8);
                    const float2 go = *reinterpret_cast<const float2 *>(pr + tl * 8 + 4);
                    if constexpr (DW) {
                        const float hIp = row[7 * HP - ROW + lp], hQp = row[8 * HP - ROW + lp], sm = hIp + hQp;
                        const float at = row[HP + lp], vc = at * row[2 * HP + lp], vs = at * row[3 * HP + lp], pa = row[lp];
                        const float acv = G[lp], asv = G[HP + lp], af = G[2 * HP + lp], gtht = G[3 * HP + lp], gpa = G[4 * HP + lp], gat = G[5 * HP + lp];
                        gbc += acv; gbs += asv; gbf += af;
                        gpt = fmaf(gtht, p0.w, gpt); gax = fmaf(gpa, p0.z, gax);
                        gwo0 = fmaf(go.x, row[7 * HP + lp], gwo0); gwo1 = fmaf(go.y, row[8 * HP + lp], gwo1);
#pragma unroll
                        for (int k = 0; k < 8; ++k) gcs[k] = fmaf(gat, fabsf(pa - knot[k]), gcs[k]);   // entries >= K are never written out
                        const float4 *G4 = reinterpret_cast<const float4 *>(G);
#pragma unroll
                        for (int k4 = 0; k4 < HP / 4; ++k4) {
                            const float4 vcv = G4[k4], vsv = G4[HP / 4 + k4], vf = G4[2 * (HP / 4) + k4], vt = G4[3 * (HP / 4) + k4], vp = G4[4 * (HP / 4) + k4];
                            const float ec[4] = {vcv.x, vcv.y, vcv.z, vcv.w}, es[4] = {vsv.x, vsv.y, vsv.z, vsv.w}, ef[4] = {vf.x, vf.y, vf.z, vf.w};
                            const float et[4] = {vt.x, vt.y, vt.z, vt.w}, ep[4] = {vp.x, vp.y, vp.z, vp.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int k = k4 * 4 + e;
                                if (k < HT) {
                                    gch[k] = fmaf(ec[e], hIp, gch[k]); gcv[k] = fmaf(ec[e], vc, gcv[k]);
                                    gsh[k] = fmaf(es[e], hQp, gsh[k]); gsvv[k] = fmaf(es[e], vs, gsvv[k]);
                                    gph[k] = fmaf(et[e], sm, gph[k]); gah[k] = fmaf(ep[e], sm, gah[k]); gff[k] = fmaf(ef[e], sm, gff[k]);
                                }
                            }
                        }
                    }
                    if (a.need_dx && isf) {
                        const float *line = G + (3 + fl) * HP;
                        float d0 = 0.f, d1 = 0.f;
#pragma unroll
                        for (int k = 0; k < HT; ++k) { if (k & 1) d1 = fmaf(wx[k], line[k], d1); else d0 = fmaf(wx[k], line[k], d0); }
                        sdf[tl * 2 + fl] = d0 + d1;
                    }
                }
                __syncwarp();
                if (lane < nt) {
                    const float *p = pr + lane * 8;
                    if constexpr (DW) { gbo0 += p[4]; gbo1 += p[5]; }
                    if (gx2) {
                        const float i = p[0], q = p[1], am = p[2], gth = sdf[lane * 2], ga = sdf[lane * 2 + 1], a2 = am * am;
                        gx2[t0 + lane] = make_float2(ga * i / am - gth * q / a2, ga * q / am + gth * i / a2);
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
                    for (int k = 0; k < HT; ++k)
                        if (k < H) {
                            prt[L.oWc + k * H2 + lane] = gch[k]; prt[L.oWc + k * H2 + H + lane] = gcv[k];
                            prt[L.oWs + k * H2 + lane] = gsh[k]; prt[L.oWs + k * H2 + H + lane] = gsvv[k];
                            prt[L.oWph + k * H + lane] = gph[k]; prt[L.oWah + k * H + lane] = gah[k]; prt[L.oWf + k * H + lane] = gff[k];
                        }
                    prt[L.oWpt + lane] = gpt; prt[L.oWax + lane] = gax; prt[L.obf + lane] = gbf; prt[L.obc + lane] = gbc; prt[L.obs + lane] = gbs;
                    prt[L.oWo1 + lane] = gwo0; prt[L.oWo2 + lane] = gwo1;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float v = warp_sum(act ? gcs[k] : 0.f);
                    if (lane == 0 && k < K) prt[L.ocs + k] = v;
                }
                gbo0 = warp_sum(gbo0); gbo1 = warp_sum(gbo1);
                if (lane == 0) { prt[L.obo1] = gbo0; prt[L.obo2] = gbo1; }
            }
        }
    }
}

// ================================================================ host dispatch
#define ODPD_JANET_TIERS(X) X(10) X(15) X(24)
static int janet_tier(int H) {
#define X(HTV) if (H <= HTV) return HTV;
    ODPD_JANET_TIERS(X)
#undef X
    return -1;
}
// The JANET cells forget slowly (gate products; measured on PGJANET at init: the adjoint loses ~0.94x per step, 2e-4 left after
// 128 steps, 4e-8 after 256), so their default warm-up is 256 steps.
static constexpr int JANET_WARM = 256;
// dir: 0 fwd, 1 bwd, +2 = plan only
template <int HT>
static int janet_launch(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *info) {
    constexpr int HP = Pad4<HT>::value;
    const bool pg = a.cell == ODPD_CELL_PGJANET;
    const int P = pg ? PgLayout(a.H).P : DvLayout(a.H, a.K).P;
    const int Ppad = (P + 3) & ~3;
    const bool plan_only = dir >= 2, fwd = (dir & 1) == 0;
    const int ss = pg ? HP : 2 * HP, ROW = pg ? PgRow<HT>::value : DvRow<HT>::value;
    const int64_t off = fwd ? (a.save ? (int64_t)a.B * a.T * ROW : 0) : ((chunk_rows(a.B, a.tchunks_req) * P + 3) & ~(int64_t)3);
    float *base = fwd ? a.saved : a.partials;
    float *scr = base ? base + off : nullptr;
#define LAUNCH(KERN, SMEMT, NAME)                                                                                             \
    {                                                                                                                         \
        static OccCache occ{};                                                                                                   \
        return chunk_launch(KERN, 96, (size_t)SMEMT::total(Ppad) * 4, &occ, a, fwd ? 0 : 1, scr, off, ss, st, plan_only, info, NAME, JANET_WARM); \
    }
    if (pg) {
        if (fwd) LAUNCH(pgjanet_fwd_kernel<HT>, PgFwdSmem<HT>, "pgjanet_fwd_kernel")
        else if (dw) LAUNCH((pgjanet_bwd_kernel<HT, true>), PgBwdSmem<HT>, "pgjanet_bwd_kernel")
        else LAUNCH((pgjanet_bwd_kernel<HT, false>), PgBwdSmem<HT>, "pgjanet_bwd_kernel")
    } else {
        if (fwd) LAUNCH(dvrjanet_fwd_kernel<HT>, DvFwdSmem<HT>, "dvrjanet_fwd_kernel")
        else if (dw) LAUNCH((dvrjanet_bwd_kernel<HT, true>), DvBwdSmem<HT>, "dvrjanet_bwd_kernel")
        else LAUNCH((dvrjanet_bwd_kernel<HT, false>), DvBwdSmem<HT>, "dvrjanet_bwd_kernel")
    }
#undef LAUNCH
}
int64_t janet_saved_floats(int cell, int B, int T, int H, bool save, int tchunks_req) {
    const int ht = janet_tier(H);
    if (ht < 0) return -1;
    const int HP = (ht + 3) & ~3;
    const bool pg = cell == ODPD_CELL_PGJANET;
    return (save ? (int64_t)B * T * (pg ? 7 : 9) * HP : 0) + chunk_fwd_scratch_floats(chunk_rows(B, tchunks_req), pg ? HP : 2 * HP);
}
int64_t janet_workspace_floats(int cell, int B, int H, int64_t P, int tchunks_req) {
    const int ht = janet_tier(H);
    if (ht < 0) return -1;
    const int HP = (ht + 3) & ~3;
    return chunk_workspace_floats(chunk_rows(B, tchunks_req), P, cell == ODPD_CELL_PGJANET ? HP : 2 * HP);
}
int janet_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *info) {
#define X(HTV) if (a.H <= HTV) return janet_launch<HTV>(a, dir, dw, st, info);
    ODPD_JANET_TIERS(X)
#undef X
    set_error("JANET kernels support hidden_size <= 24 (got %d)", a.H);
    return -1;
}

}  // namespace odpd

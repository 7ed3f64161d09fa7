// qgru_qat.cu — fake-quantised (QAT) GRU cell of BASELINE config 5 (W8A8 / W16A16 quant-aware DPD).
//
// Replaces the arithmetic the reference builds by module surgery (quant/quant_envs.py:132-305) around backbones/qgru.py:
//   quant/modules/gru.py GRUCell.forward :32-61, GRU.forward :85-124; INT_Linear quant_layers.py:70-82 (weights and inputs each
//   fake-quantised, float bias, 16-bit out_quantizer only on fc_out and only in eval); Quant_sigmoid/tanh/mult/add quant_ops.py:14-66;
//   INT_Quantizer quantizers.py:56-81:  s = 2^round(log2|scale|),  q(v) = s*rne(clamp(v/s, -2^(b-1), 2^(b-1)-1)),
//   backward = straight-through inside the clamp, 0 outside; the 13 scale parameters receive zero gradient.
//
// One warp per sequence (4 per CTA), lane j owns unit j.  A single rounding-boundary flip costs a whole quantum (2^-6 at 8 bits), so
// this cell uses the accurate libdevice expf/tanhf and an IEEE divide in the sigmoid rather than the 2-MUFU forms of the float cells;
// the quantisers themselves divide by a power of two, done as an exact multiply.  The IQ samples, dLoss/dout and the saved
// h_{t-1} of 32 timesteps are fetched at once (one timestep per lane, coalesced) into shared memory, so the serial step loop
// carries no global-memory latency (round 1 issued 1-4 dependent global loads per step: ~3900 cycles per step).
// Backward recomputes the step's forward from the saved h_{t-1} (only h is saved: T*HP floats per sequence).
#include "cells.h"
#include "pipeline.cuh"

namespace odpd {

struct QLayout {
    int H, oWx, obx, osx, oWh, obh, osh, osop, oWo, obo, oso, P;
    __host__ __device__ explicit QLayout(int h) {
        H = h; oWx = 0; obx = 12 * h; osx = obx + 3 * h; oWh = osx + 3; obh = oWh + 3 * h * h; osh = obh + 3 * h; osop = osh + 3;
        oWo = osop + 4; obo = oWo + 2 * h; oso = obo + 2; P = oso + 3;
    }
};
struct Quant { float s, inv, qn, qp; };
__device__ __forceinline__ Quant mkq(float scale, int bits) {
    Quant q;
    const float e = rintf(log2f(fabsf(scale)));
    q.s = exp2f(e);
    q.inv = exp2f(-e);           // s is an exact power of two: v * (1/s) == v / s bit for bit, at a tenth of the cost of an IEEE divide
    q.qn = -exp2f((float)(bits - 1));
    q.qp = exp2f((float)(bits - 1)) - 1.f;
    return q;
}
__device__ __forceinline__ float qf(const Quant &q, float v, bool &in) {
    float u = v * q.inv;
    in = (u >= q.qn) && (u <= q.qp);
    u = fminf(fmaxf(u, q.qn), q.qp);
    if (!(v == v)) u = v;   // NaN propagates like torch.clamp
    return rintf(u) * q.s;
}
__device__ __forceinline__ float qf(const Quant &q, float v) { bool in; return qf(q, v, in); }
__device__ __forceinline__ float sig_acc(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }

template <int HT> struct QStep {   // forward intermediates of one unit at one timestep (recomputed in the backward)
    float hq, sr, sz, tn, r, z, n, hgn, hnew;
    bool c_hq, c_ar, c_az, c_an, c_r, c_z, c_n, c_m1, c_m2, c_m3, c_h;
};

template <int HT, int FM>
struct QCtx {
    static constexpr int HP = Pad4<HT>::value;
    float wxq[12], whq[3 * HT], bx[3], bh[3];
    Quant qxa, qha, qsig, qtanh, qadd, qmul;
    // one timestep: fq = quantised features (uniform), hp = h_{t-1} of this lane, line = shared broadcast line [HP]
    __device__ __forceinline__ void step(const float *fq, float hp, float *line, int lane, QStep<HT> &o) const {
        o.hq = qf(qha, hp, o.c_hq);
        if (lane < HP) line[lane] = o.hq;
        __syncwarp();
        float xg[3], hg[3];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            xg[g] = bx[g];
#pragma unroll
            for (int k = 0; k < 4; ++k) xg[g] = fmaf(wxq[g * 4 + k], fq[k], xg[g]);
        }
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f, c0 = 0.f, c1 = 0.f;
        const float4 *l4 = reinterpret_cast<const float4 *>(line);
#pragma unroll
        for (int q = 0; q < HP / 4; ++q) {
            const float4 v = l4[q];
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = q * 4 + i;
                if (k < HT) {
                    if (k & 1) { a1 = fmaf(whq[k], e[i], a1); b1 = fmaf(whq[HT + k], e[i], b1); c1 = fmaf(whq[2 * HT + k], e[i], c1); }
                    else       { a0 = fmaf(whq[k], e[i], a0); b0 = fmaf(whq[HT + k], e[i], b0); c0 = fmaf(whq[2 * HT + k], e[i], c0); }
                }
            }
        }
        hg[0] = (a0 + a1) + bh[0]; hg[1] = (b0 + b1) + bh[1]; hg[2] = (c0 + c1) + bh[2];
        const float ar = qf(qadd, xg[0] + hg[0], o.c_ar); o.sr = sig_acc(ar); o.r = qf(qsig, o.sr, o.c_r);
        const float az = qf(qadd, xg[1] + hg[1], o.c_az); o.sz = sig_acc(az); o.z = qf(qsig, o.sz, o.c_z);
        o.hgn = hg[2];
        const float m1 = qf(qmul, o.r * hg[2], o.c_m1);
        const float an = qf(qadd, xg[2] + m1, o.c_an); o.tn = tanhf(an); o.n = qf(qtanh, o.tn, o.c_n);
        const float m2 = qf(qmul, o.z * hp, o.c_m2);
        const float m3 = qf(qmul, (1.f - o.z) * o.n, o.c_m3);
        o.hnew = qf(qadd, m2 + m3, o.c_h);
        __syncwarp();
    }
};

template <int HT, int FM>
__device__ __forceinline__ void qctx_init(QCtx<HT, FM> &c, const float *sp, const QLayout &L, int lane, int bw, int ba, unsigned &cwx, unsigned (&cwh)[3]) {
    const int H = L.H;
    const bool act = lane < H;
    const int j = act ? lane : 0;
    const Quant qxw = mkq(sp[L.osx], bw), qhw = mkq(sp[L.osh], bw);
    c.qxa = mkq(sp[L.osx + 1], ba); c.qha = mkq(sp[L.osh + 1], ba);
    c.qsig = mkq(sp[L.osop], ba); c.qtanh = mkq(sp[L.osop + 1], ba); c.qadd = mkq(sp[L.osop + 2], ba); c.qmul = mkq(sp[L.osop + 3], ba);
    cwx = 0; cwh[0] = cwh[1] = cwh[2] = 0;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bool in; const float w = qf(qxw, sp[L.oWx + (g * H + j) * 4 + k], in);
            c.wxq[g * 4 + k] = act ? w : 0.f;
            if (in) cwx |= 1u << (g * 4 + k);
        }
#pragma unroll
        for (int k = 0; k < HT; ++k) {
            bool in = false; float w = 0.f;
            if (act && k < H) w = qf(qhw, sp[L.oWh + (g * H + j) * H + k], in);
            c.whq[g * HT + k] = w;
            if (in) cwh[g] |= 1u << k;
        }
        c.bx[g] = act ? sp[L.obx + g * H + j] : 0.f;
        c.bh[g] = act ? sp[L.obh + g * H + j] : 0.f;
    }
}

template <int FM>
__device__ __forceinline__ void q_features(IqRow x2, int t, float *f) {
    const float2 v = __ldg(x2 + t);
    float ff[8];
    features_fwd<FM>(v.x, v.y, 0.f, 0.f, ff);
    f[0] = ff[0]; f[1] = ff[1]; f[2] = ff[2]; f[3] = ff[3];
}

// ================================================================ forward
template <int HT, int FM>
__global__ void __launch_bounds__(128) qgru_qat_fwd_kernel(GruArgs a) {
    constexpr int HP = Pad4<HT>::value;
    const QLayout L(a.H);
    const int H = a.H, T = a.T, bw = a.K & 255, ba = (a.K >> 8) & 255, eval = (a.K >> 16) & 1;
    extern __shared__ __align__(16) float smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 4;
    const int Ppad = (L.P + 3) & ~3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    float *line = sp + Ppad + warp * HP;
    stage_params(sp, a.params, L.P, bar);
    const int b = blockIdx.x * wpc + warp;
    if (b >= a.B) return;
    QCtx<HT, FM> c;
    unsigned cwx, cwh[3];
    qctx_init<HT, FM>(c, sp, L, lane, bw, ba, cwx, cwh);
    const bool act = lane < H;
    const Quant qow = mkq(sp[L.oso], bw), qoa = mkq(sp[L.oso + 1], ba), qoo = mkq(sp[L.oso + 2], 16);
    const float wo0 = act ? qf(qow, sp[L.oWo + (act ? lane : 0)]) : 0.f, wo1 = act ? qf(qow, sp[L.oWo + H + (act ? lane : 0)]) : 0.f;
    const float bo0 = sp[L.obo], bo1 = sp[L.obo + 1];
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
    const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
    float2 *o2 = reinterpret_cast<float2 *>(a.out) + (size_t)b * T;
    float *sv = a.save ? a.saved + (size_t)b * T * HP : nullptr;
    float h = 0.f, lsum = 0.f;
    float *sfq = sp + Ppad + wpc * HP + warp * (CH * 4);            // [CH][4] quantised features of the block (this warp)
    for (int t0 = 0; t0 < T; t0 += CH) {
        const int nt = min(CH, T - t0);
        float2 yv = make_float2(0.f, 0.f);
        if (lane < nt) {                                            // one timestep per lane
            float f[4];
            q_features<FM>(x2, t0 + lane, f);
            *reinterpret_cast<float4 *>(sfq + lane * 4) = make_float4(qf(c.qxa, f[0]), qf(c.qxa, f[1]), qf(c.qxa, f[2]), qf(c.qxa, f[3]));
            if (y2) yv = __ldg(y2 + t0 + lane);
        }
        __syncwarp();
        float2 ov = make_float2(0.f, 0.f);                          // output of this lane's timestep
        for (int tl = 0; tl < nt; ++tl) {
            const float4 q4 = *reinterpret_cast<const float4 *>(sfq + tl * 4);
            const float fq[4] = {q4.x, q4.y, q4.z, q4.w};
            QStep<HT> s;
            c.step(fq, h, line, lane, s);
            h = act ? s.hnew : 0.f;
            if (sv && lane < HP) sv[(size_t)(t0 + tl) * HP + lane] = h;
            const float qo = qf(qoa, h);
            float y0 = warp_sum(wo0 * qo) + bo0, y1 = warp_sum(wo1 * qo) + bo1;
            if (eval) { y0 = qf(qoo, y0); y1 = qf(qoo, y1); }
            if (lane == tl) ov = make_float2(y0, y1);
        }
        if (lane < nt) {
            o2[t0 + lane] = ov;                                     // coalesced store of the block's outputs
            if (y2) { const float d0 = ov.x - yv.x, d1 = ov.y - yv.y; lsum = fmaf(d0, d0, fmaf(d1, d1, lsum)); }
        }
        __syncwarp();
    }
    lsum = warp_sum(lsum);
    if (a.loss && y2 && lane == 0) atomicAdd(a.loss, (double)lsum * (double)a.loss_scale);
}

// ================================================================ backward
template <int HT, int FM, bool DW>
__global__ void __launch_bounds__(128) qgru_qat_bwd_kernel(GruArgs a) {
    constexpr int HP = Pad4<HT>::value;
    const QLayout L(a.H);
    const int H = a.H, T = a.T, bw = a.K & 255, ba = (a.K >> 8) & 255;
    extern __shared__ __align__(16) float smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 4;
    const int Ppad = (L.P + 3) & ~3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    float *line = sp + Ppad + warp * 5 * HP;       // [HP] hq line + [3HP] gate-gradient lines (+pad)
    float *gl = line + HP;
    stage_params(sp, a.params, L.P, bar);
    const int b = blockIdx.x * wpc + warp;
    if (b >= a.B) return;
    QCtx<HT, FM> c;
    unsigned cwx, cwh[3];
    qctx_init<HT, FM>(c, sp, L, lane, bw, ba, cwx, cwh);
    const bool act = lane < H;
    const int j = act ? lane : 0;
    const Quant qhw = mkq(sp[L.osh], bw), qow = mkq(sp[L.oso], bw), qoa = mkq(sp[L.oso + 1], ba);
    bool cwo0, cwo1;
    float wo0 = qf(qow, sp[L.oWo + j], cwo0), wo1 = qf(qow, sp[L.oWo + H + j], cwo1);
    if (!act) { wo0 = wo1 = 0.f; }
    float whc[3 * HT];   // column j of the quantised W_h
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int k = 0; k < HT; ++k) whc[g * HT + k] = (act && k < H) ? qf(qhw, sp[L.oWh + (g * H + k) * H + j]) : 0.f;
    float gwh[DW ? 3 * HT : 1], gwx[12], gbx[3] = {0.f, 0.f, 0.f}, gbh[3] = {0.f, 0.f, 0.f}, gwo0 = 0.f, gwo1 = 0.f, gbo0 = 0.f, gbo1 = 0.f;
    if constexpr (DW) {
#pragma unroll
        for (int k = 0; k < 3 * HT; ++k) gwh[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) gwx[k] = 0.f;
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
    const float2 *go2 = a.gout ? reinterpret_cast<const float2 *>(a.gout) + (size_t)b * T : nullptr;
    const float2 *oi2 = a.out_in ? reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * T : nullptr;
    const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
    float2 *gx2 = (a.need_dx && a.gx) ? reinterpret_cast<float2 *>(a.gx) + (size_t)b * T : nullptr;
    const float *sv = a.saved + (size_t)b * T * HP;
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
    float gH = 0.f;
    float *sblk = sp + Ppad + wpc * 5 * HP + warp * (CH * (8 + HP));  // this warp's block staging: [CH][4] feat | [CH][2] go | [CH][2] x | [CH][HP] h_{t-1}
    float *sfe = sblk, *sgo = sblk + CH * 4, *sxx = sblk + CH * 6, *shp = sblk + CH * 8;
    for (int tb = ((T - 1) / CH) * CH; tb >= 0; tb -= CH) {
      const int ntb = min(CH, T - tb);
      __syncwarp();
      if (lane < ntb) {                                               // one timestep per lane: features, dLoss/dout, raw sample
          float f[4];
          q_features<FM>(x2, tb + lane, f);
          *reinterpret_cast<float4 *>(sfe + lane * 4) = make_float4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<float2 *>(sgo + lane * 2) = load_gout(go2, oi2, y2, tb + lane, gs);
          *reinterpret_cast<float2 *>(sxx + lane * 2) = __ldg(x2 + tb + lane);
      }
      for (int i = lane; i < ntb * HP; i += 32) {                     // h_{t-1} of the block's steps (row tb-1 .. tb+ntb-2), coalesced
          const int tt = tb + i / HP - 1;
          shp[i] = tt >= 0 ? __ldg(sv + (size_t)tt * HP + (i % HP)) : 0.f;
      }
      __syncwarp();
      for (int tl = ntb - 1; tl >= 0; --tl) {
        const int t = tb + tl;
        const float hp = lane < HP ? shp[tl * HP + lane] : 0.f;
        float fq[4];
        bool cf[4];
        {
            const float4 f4 = *reinterpret_cast<const float4 *>(sfe + tl * 4);
            const float f[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) fq[k] = qf(c.qxa, f[k], cf[k]);
        }
        QStep<HT> s;
        c.step(fq, hp, line, lane, s);          // leaves hq of all units in `line`
        const float2 go = *reinterpret_cast<const float2 *>(sgo + tl * 2);
        // output head
        bool cq;
        const float hn = act ? s.hnew : 0.f;
        const float qo = qf(qoa, hn, cq);
        if constexpr (DW) { if (cwo0) gwo0 = fmaf(go.x, qo, gwo0); if (cwo1) gwo1 = fmaf(go.y, qo, gwo1); gbo0 += go.x; gbo1 += go.y; }
        if (cq) gH += fmaf(wo0, go.x, wo1 * go.y);
        // cell backward (straight-through masks)
        const float gs3 = s.c_h ? gH : 0.f;
        const float gm2 = s.c_m2 ? gs3 : 0.f, gm3 = s.c_m3 ? gs3 : 0.f;
        const float gz = gm2 * hp - gm3 * s.n;
        const float ghp = gm2 * s.z;
        const float gtn = s.c_n ? gm3 * (1.f - s.z) : 0.f;
        const float gan = gtn * (1.f - s.tn * s.tn);
        const float gsn = s.c_an ? gan : 0.f;
        const float gm1 = s.c_m1 ? gsn : 0.f;
        const float gsr = s.c_r ? gm1 * s.hgn : 0.f;
        const float gar = gsr * s.sr * (1.f - s.sr);
        const float gsum_r = s.c_ar ? gar : 0.f;
        const float gsz = s.c_z ? gz : 0.f;
        const float gaz = gsz * s.sz * (1.f - s.sz);
        const float gsum_z = s.c_az ? gaz : 0.f;
        const float g_xg[3] = {gsum_r, gsum_z, gsn}, g_hg[3] = {gsum_r, gsum_z, gm1 * s.r};
        if (lane < HP) { gl[lane] = act ? g_hg[0] : 0.f; gl[HP + lane] = act ? g_hg[1] : 0.f; gl[2 * HP + lane] = act ? g_hg[2] : 0.f; }
        __syncwarp();
        // d/dh_{t-1} through the quantised recurrent matvec, and dW_h (row form) with the hq line still in shared memory
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const float4 *g4 = reinterpret_cast<const float4 *>(gl + g * HP);
            const float4 *h4 = reinterpret_cast<const float4 *>(line);
#pragma unroll
            for (int q = 0; q < HP / 4; ++q) {
                const float4 gv = g4[q], hv = h4[q];
                const float ge[4] = {gv.x, gv.y, gv.z, gv.w}, he[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = q * 4 + i;
                    if (k < HT) {
                        if (k & 1) q1 = fmaf(whc[g * HT + k], ge[i], q1); else q0 = fmaf(whc[g * HT + k], ge[i], q0);
                        if constexpr (DW) { if ((cwh[g] >> k) & 1u) gwh[g * HT + k] = fmaf(g_hg[g], he[i], gwh[g * HT + k]); }
                    }
                }
            }
        }
        if constexpr (DW) {
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                gbx[g] += g_xg[g]; gbh[g] += g_hg[g];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if ((cwx >> (g * 4 + k)) & 1u) gwx[g * 4 + k] = fmaf(g_xg[g], fq[k], gwx[g * 4 + k]);
            }
        }
        if (gx2) {
            float gf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float part = act ? (c.wxq[k] * g_xg[0] + c.wxq[4 + k] * g_xg[1] + c.wxq[8 + k] * g_xg[2]) : 0.f;
                const float tot = warp_sum(part);
                gf[k] = cf[k] ? tot : 0.f;
            }
            if (lane == 0) {
                const float2 v = *reinterpret_cast<const float2 *>(sxx + tl * 2);
                float gi, gq;
                features_bwd<FM>(v.x, v.y, gf, gi, gq);
                gx2[t] = make_float2(gi, gq);
            }
        }
        gH = act ? ghp + (s.c_hq ? (q0 + q1) : 0.f) : 0.f;
        __syncwarp();
      }
    }
    if constexpr (DW) {
        if (a.partials) {
            float *prt = a.partials + (size_t)b * L.P;
            if (act) {
#pragma unroll
                for (int g = 0; g < 3; ++g) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) prt[L.oWx + (g * H + lane) * 4 + k] = gwx[g * 4 + k];
#pragma unroll
                    for (int k = 0; k < HT; ++k)
                        if (k < H) prt[L.oWh + (g * H + lane) * H + k] = gwh[g * HT + k];
                    prt[L.obx + g * H + lane] = gbx[g]; prt[L.obh + g * H + lane] = gbh[g];
                }
                prt[L.oWo + lane] = gwo0; prt[L.oWo + H + lane] = gwo1;
            }
            if (lane == 0) {
                prt[L.obo] = gbo0; prt[L.obo + 1] = gbo1;
                for (int k = 0; k < 3; ++k) { prt[L.osx + k] = 0.f; prt[L.osh + k] = 0.f; prt[L.oso + k] = 0.f; }   // scale params: zero grad
                for (int k = 0; k < 4; ++k) prt[L.osop + k] = 0.f;
            }
        }
    }
}

#define ODPD_QAT_TIERS(X) X(10) X(16) X(20) X(32)     // 20 / 30: bash_scripts/quant_qgru_dpd_regr.sh:74
static int qat_tier(int H) {
#define X(HTV) if (H <= HTV) return HTV;
    ODPD_QAT_TIERS(X)
#undef X
    return -1;
}
template <int HT, int FM>
static int qat_launch(const GruArgs &a, int dir, bool dw, cudaStream_t st) {
    const QLayout L(a.H);
    const int Ppad = (L.P + 3) & ~3, wpc = 4, HP = Pad4<HT>::value;
    const unsigned grid = (unsigned)((a.B + wpc - 1) / wpc);
    if (dir == 0) {
        const size_t smem = (size_t)(4 + Ppad + wpc * HP + wpc * CH * 4) * 4;
        qgru_qat_fwd_kernel<HT, FM><<<grid, wpc * 32, smem, st>>>(a);
    } else {
        const size_t smem = (size_t)(4 + Ppad + wpc * 5 * HP + wpc * CH * (8 + HP)) * 4;
        if (dw) qgru_qat_bwd_kernel<HT, FM, true><<<grid, wpc * 32, smem, st>>>(a);
        else qgru_qat_bwd_kernel<HT, FM, false><<<grid, wpc * 32, smem, st>>>(a);
    }
    return check_launch("qgru_qat kernel");
}
int64_t qat_nparams(int H) { return QLayout(H).P; }
int64_t qat_saved_floats(int B, int T, int H) {
    const int ht = qat_tier(H);
    return ht < 0 ? -1 : (int64_t)B * T * ((ht + 3) & ~3);
}
int qat_run(const GruArgs &a, int dir, bool dw, cudaStream_t st) {
    const bool amp1 = a.cell == ODPD_CELL_QGRU_AMP1_QAT;
#define X(HTV) if (a.H <= HTV) return amp1 ? qat_launch<HTV, FM_AMP4>(a, dir, dw, st) : qat_launch<HTV, FM_QGRU4>(a, dir, dw, st);
    ODPD_QAT_TIERS(X)
#undef X
    set_error("QAT GRU kernels support hidden_size <= 32 (got %d)", a.H);
    return -1;
}

}  // namespace odpd

// tres_qat.cu — fake-quantised (QAT) TRes-DeltaGRU: the W16A16 stage of bash_scripts/OpenDPDv2.sh:47-49.
//
// Replaces what the reference builds by module surgery (quant/quant_envs.py:286-305) around backbones/deltagru_tcnskip.py:
//   x2h / h2h / fc_out -> INT_Linear (weight and input each fake-quantised, quant_layers.py:70-82; 16-bit output quantiser on fc_out in eval only);
//   the layer's add / mul / sigmoid / tanh modules -> Quant_add / Quant_mult / Quant_sigmoid / Quant_tanh (quant_ops.py:14-66).
//   The delta logic stays in float (deltagru_tcnskip.py:266-291): dx, dh are thresholded exactly as in the float cell (same masks, counters on the
//   unquantised deltas); mac_x = x2h(Q(dx)) + M, mac_h = h2h(Q(dh)); r = Qs(sigmoid(M_r)), z = Qs(sigmoid(M_z));
//   n = Qt(tanh(Qa(M_n + Qm(r M_nh))));  h = Qa(Qm(Qa(1 - z) n) + Qm(z h));  out = fc_out(Q(h)) + tcn(x)   (the TCN skip path stays float).
//   Quantiser (quantizers.py:56-81): s = 2^round(log2|scale|), q(v) = s*rne(clamp(v/s, -2^(b-1), 2^(b-1)-1)); backward = straight-through inside
//   the clamp, 0 outside; the 13 scale parameters receive zero gradient.
//
// One warp per sequence, lane = hidden unit (hidden_size <= 16: the script uses 15).  Like the QAT GRU (qgru_qat.cu) this cell uses the accurate
// expf / tanhf and an IEEE divide: one rounding-boundary flip costs a whole quantum.  Forward saves, per step and unit, the quantised and raw gate
// values, the quantised deltas and the clamp flags (the accumulators M cannot be recomputed from h alone) plus the two keep-mask words; backward
// replays them in reverse (SURVEY §8a-D with the STE flags), accumulates the x2h / h2h / fc_out gradients in registers, and leaves dL/dfeatures per
// step for a time-parallel pass that applies the feature Jacobian, the next-sample roll and the transposed TCN.
//
// Flat parameter layout (named_parameters order of the surgered model): rnn.x2h.weight(3H,6) + weight/act/out scales | rnn.h2h.weight(3H,H) + 3 scales |
// rnn.add / mul / sigmoid / tanh .quantizer.scale | fc_out.weight(2,H) + 3 scales | tcn.0.weight(3,2,3) tcn.2.weight(2,3,1)  = 3H^2 + 20H + 37.
// OdpdDims.K packs n_bits_w | n_bits_a<<8 | eval<<16.
#include "cells.h"
#include "chunking.cuh"

namespace odpd {

static constexpr int TQ_H = 16;                 // lanes that carry units
static constexpr int TQ_ROW = 12 * TQ_H + 12;   // saved floats per step: 12 slots x 16 units | qdx(6) cdxbits - - - | mask_x mask_h

struct TqLayout {
    int H, oWx, osx, oWh, osh, osop, oWo, oso, ow0, ow2, P;
    __host__ __device__ explicit TqLayout(int h) {
        H = h; oWx = 0; osx = 18 * h; oWh = osx + 3; osh = oWh + 3 * h * h; osop = osh + 3; oWo = osop + 4; oso = oWo + 2 * h; ow0 = oso + 3;
        ow2 = ow0 + 18; P = ow2 + 6;
    }
};
struct TQuant { float s, inv, qn, qp; };
__device__ __forceinline__ TQuant tq_mk(float scale, int bits) {
    TQuant q;
    const float e = rintf(log2f(fabsf(scale)));
    q.s = exp2f(e); q.inv = exp2f(-e);
    q.qn = -exp2f((float)(bits - 1)); q.qp = exp2f((float)(bits - 1)) - 1.f;
    return q;
}
__device__ __forceinline__ float tq_f(const TQuant &q, float v, int &in) {
    float u = v * q.inv;
    in = (u >= q.qn) && (u <= q.qp);
    u = fminf(fmaxf(u, q.qn), q.qp);
    if (!(v == v)) u = v;
    return rintf(u) * q.s;
}
__device__ __forceinline__ float tq_sig(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }
__device__ __forceinline__ float tq_hsw(float v) { return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f; }
__device__ __forceinline__ float tq_hsw_grad(float v) { return v < -3.f ? 0.f : (v <= 3.f ? v / 3.f + 0.5f : 1.f); }

// TCN skip path at frame position t (deltagru_tcnskip.py:32-49): cv[0..2] = first conv (taps t-16, t, t+16, zero outside), cv[3..4] = second conv
__device__ __forceinline__ void tq_tcn(const IqRow &x2, const float *w0, const float *w2, int t, int T, float (&cv)[5]) {
    float2 xs[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int tt = t + (k - 1) * 16;
        xs[k] = (tt >= 0 && tt < T) ? x2.ld(tt) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int co = 0; co < 3; ++co) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc = fmaf(w0[(co * 2) * 3 + k], xs[k].x, fmaf(w0[(co * 2 + 1) * 3 + k], xs[k].y, acc));
        cv[co] = acc;
    }
#pragma unroll
    for (int o = 0; o < 2; ++o) cv[3 + o] = fmaf(w2[o * 3], tq_hsw(cv[0]), fmaf(w2[o * 3 + 1], tq_hsw(cv[1]), w2[o * 3 + 2] * tq_hsw(cv[2])));
}

// ================================================================ forward
__global__ void __launch_bounds__(128) tresq_fwd_kernel(GruArgs a) {
    pdl_enter();
    const TqLayout L(a.H);
    const int H = a.H, T = a.T, lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + wi;
    __shared__ float sfeat_all[4][32 * 8];
    if (b >= a.B) return;
    float *sfeat = sfeat_all[wi];
    const int bw = a.K & 255, ba = (a.K >> 8) & 255, evalm = (a.K >> 16) & 1;
    const float *P = a.params;
    const bool act = lane < H;
    const int j = act ? lane : 0;
    const TQuant qxw = tq_mk(__ldg(P + L.osx), bw), qxa = tq_mk(__ldg(P + L.osx + 1), ba), qhw = tq_mk(__ldg(P + L.osh), bw), qha = tq_mk(__ldg(P + L.osh + 1), ba);
    const TQuant qadd = tq_mk(__ldg(P + L.osop), ba), qmul = tq_mk(__ldg(P + L.osop + 1), ba), qsig = tq_mk(__ldg(P + L.osop + 2), ba),
                 qtanh = tq_mk(__ldg(P + L.osop + 3), ba);
    const TQuant qow = tq_mk(__ldg(P + L.oso), bw), qoa = tq_mk(__ldg(P + L.oso + 1), ba), qoo = tq_mk(__ldg(P + L.oso + 2), 16);
    float wx[18], wh[3 * TQ_H], wo[2];
    int dummy;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
#pragma unroll
        for (int k = 0; k < 6; ++k) wx[g * 6 + k] = act ? tq_f(qxw, __ldg(P + L.oWx + (g * H + j) * 6 + k), dummy) : 0.f;
#pragma unroll
        for (int k = 0; k < TQ_H; ++k) wh[g * TQ_H + k] = (act && k < H) ? tq_f(qhw, __ldg(P + L.oWh + (g * H + j) * H + k), dummy) : 0.f;
    }
    wo[0] = act ? tq_f(qow, __ldg(P + L.oWo + j), dummy) : 0.f;
    wo[1] = act ? tq_f(qow, __ldg(P + L.oWo + H + j), dummy) : 0.f;
    float w0[18], w2[6];
#pragma unroll
    for (int i = 0; i < 18; ++i) w0[i] = __ldg(P + L.ow0 + i);
#pragma unroll
    for (int i = 0; i < 6; ++i) w2[i] = __ldg(P + L.ow2 + i);
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
    const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
    float *srow = a.saved + (size_t)b * T * TQ_ROW;
    float h = 0.f, hp = 0.f, Mr = 0.f, Mz = 0.f, Mn = 0.f, Mnh = 0.f, xp[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float lsum = 0.f;
    long long zx = 0, zh = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int nt = min(32, T - t0);
        {   // lane = timestep: features (next sample wraps, deltagru_tcnskip.py:91) and the TCN skip
            float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (lane < nt) {
                const int t = t0 + lane;
                const float2 v = x2.ld(t), vn = x2.ld(t + 1 < T ? t + 1 : 0);
                features_fwd<FM_TRES6>(v.x, v.y, vn.x, vn.y, f);
                float cv[5];
                tq_tcn(x2, w0, w2, t, T, cv);
                f[6] = tq_hsw(cv[3]); f[7] = tq_hsw(cv[4]);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) sfeat[lane * 8 + k] = f[k];
        }
        __syncwarp();
        for (int tl = 0; tl < nt; ++tl) {
            const int t = t0 + tl;
            float *row = srow + (size_t)t * TQ_ROW;
            // delta x (uniform over the lanes)
            float qdx[6];
            unsigned mx = 0, cdxb = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const float fk = sfeat[tl * 8 + k];
                float d = fk - xp[k];
                const float ad = fabsf(d);
                if (ad < a.thx) d = 0.f;
                if (ad >= a.thx) { xp[k] = fk; mx |= 1u << k; }
                zx += (d == 0.f);
                int in;
                qdx[k] = tq_f(qxa, d, in);
                cdxb |= (unsigned)in << k;
            }
            // delta h (lane = unit)
            float dh = h - hp;
            const float adh = fabsf(dh);
            const bool keep = act && adh >= a.thh;
            if (adh < a.thh) dh = 0.f;
            if (keep) hp = h;
            const unsigned mh = __ballot_sync(ODPD_FULL, keep);
            zh += __popc(__ballot_sync(ODPD_FULL, act && dh == 0.f));
            int cdh;
            const float qdh = tq_f(qha, dh, cdh);
            float mxr = Mr, mxz = Mz, mxn = Mn;
            {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int k = 0; k < 6; ++k) { a0 = fmaf(wx[k], qdx[k], a0); a1 = fmaf(wx[6 + k], qdx[k], a1); a2 = fmaf(wx[12 + k], qdx[k], a2); }
                mxr = a0 + Mr; mxz = a1 + Mz; mxn = a2 + Mn;
            }
            float hr = 0.f, hz = 0.f, hn_ = 0.f;
#pragma unroll
            for (int k = 0; k < TQ_H; ++k) {
                const float v = __shfl_sync(ODPD_FULL, qdh, k);
                hr = fmaf(wh[k], v, hr); hz = fmaf(wh[TQ_H + k], v, hz); hn_ = fmaf(wh[2 * TQ_H + k], v, hn_);
            }
            Mr = mxr + hr; Mz = mxz + hz; Mn = mxn; Mnh = hn_ + Mnh;
            int c_r, c_z, c_m1, c_an, c_n, c_omz, c_m3, c_m2, c_h, c_oa;
            const float sr = tq_sig(Mr), sz = tq_sig(Mz);
            const float r = tq_f(qsig, sr, c_r), z = tq_f(qsig, sz, c_z);
            const float m1 = tq_f(qmul, r * Mnh, c_m1);
            const float an = tq_f(qadd, Mn + m1, c_an);
            const float tn = tanhf(an), n = tq_f(qtanh, tn, c_n);
            const float omz = tq_f(qadd, 1.f + (-z), c_omz);
            const float m3 = tq_f(qmul, omz * n, c_m3), m2 = tq_f(qmul, z * h, c_m2);
            const float hnew = tq_f(qadd, m3 + m2, c_h);
            const float hq = tq_f(qoa, hnew, c_oa);
            h = act ? hnew : 0.f;
            float o0 = warp_sum(act ? wo[0] * hq : 0.f), o1 = warp_sum(act ? wo[1] * hq : 0.f);
            if (evalm) { int in; o0 = tq_f(qoo, o0, in); o1 = tq_f(qoo, o1, in); }
            o0 += sfeat[tl * 8 + 6]; o1 += sfeat[tl * 8 + 7];
            if (lane == 0) {
                reinterpret_cast<float2 *>(a.out)[(size_t)b * T + t] = make_float2(o0, o1);
                if (y2) { const float2 y = y2.ld(t); const float d0 = o0 - y.x, d1 = o1 - y.y; lsum = fmaf(d0, d0, fmaf(d1, d1, lsum)); }
            }
            if (a.save) {
                if (lane < TQ_H) {
                    const int fl = c_r | c_z << 1 | c_m1 << 2 | c_an << 3 | c_n << 4 | c_omz << 5 | c_m3 << 6 | c_m2 << 7 | c_h << 8 | c_oa << 9 | cdh << 10;
                    row[lane] = act ? qdh : 0.f; row[TQ_H + lane] = r; row[2 * TQ_H + lane] = z; row[3 * TQ_H + lane] = n; row[4 * TQ_H + lane] = sr;
                    row[5 * TQ_H + lane] = sz; row[6 * TQ_H + lane] = tn; row[7 * TQ_H + lane] = Mnh; row[8 * TQ_H + lane] = omz;
                    row[9 * TQ_H + lane] = h; row[10 * TQ_H + lane] = hq; row[11 * TQ_H + lane] = __int_as_float(act ? fl : 0);
                }
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) row[12 * TQ_H + k] = qdx[k];
                }
                if (lane == 6) row[12 * TQ_H + 6] = __int_as_float((int)cdxb);
                if (lane == 7) { row[TQ_ROW - 2] = __int_as_float((int)mx); row[TQ_ROW - 1] = __int_as_float((int)mh); }
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (a.loss && y2) atomicAdd(a.loss, (double)lsum * (double)a.loss_scale);
        if (a.stats) {
            atomicAdd(reinterpret_cast<unsigned long long *>(a.stats), (unsigned long long)zx);
            atomicAdd(reinterpret_cast<unsigned long long *>(a.stats) + 1, (unsigned long long)T * 6);
            atomicAdd(reinterpret_cast<unsigned long long *>(a.stats) + 2, (unsigned long long)zh);
            atomicAdd(reinterpret_cast<unsigned long long *>(a.stats) + 3, (unsigned long long)T * H);
        }
    }
}

// ================================================================ backward: the cell in reverse, one warp per sequence
// Writes the sequence's gradient-partial row (x2h / h2h / fc_out with the STE flags of the weight quantisers, zeros for the scales) and
// GF[b][t][8] = dL/dfeatures of step t (for tresq_post_kernel).
__global__ void __launch_bounds__(128) tresq_bwd_kernel(GruArgs a, float *gf_out, int dw) {
    pdl_enter();
    const TqLayout L(a.H);
    const int H = a.H, T = a.T, lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + wi;
    __shared__ float sWx[3 * TQ_H * 6], sWh[3 * TQ_H * TQ_H];
    __shared__ __align__(16) float sline_all[4][8 * TQ_H];
    const int bw = a.K & 255, ba = (a.K >> 8) & 255;
    const float *P = a.params;
    const TQuant qxw = tq_mk(__ldg(P + L.osx), bw), qhw = tq_mk(__ldg(P + L.osh), bw), qow = tq_mk(__ldg(P + L.oso), bw);
    (void)ba;
    // quantised weights, [g][j][k] with j, k padded to 16 / 6
    for (int i = threadIdx.x; i < 3 * TQ_H * 6; i += blockDim.x) {
        const int g = i / (TQ_H * 6), jj = (i / 6) % TQ_H, k = i % 6;
        int in;
        sWx[i] = jj < H ? tq_f(qxw, __ldg(P + L.oWx + (g * H + jj) * 6 + k), in) : 0.f;
    }
    for (int i = threadIdx.x; i < 3 * TQ_H * TQ_H; i += blockDim.x) {
        const int g = i / (TQ_H * TQ_H), jj = (i / TQ_H) % TQ_H, k = i % TQ_H;
        int in;
        sWh[i] = (jj < H && k < H) ? tq_f(qhw, __ldg(P + L.oWh + (g * H + jj) * H + k), in) : 0.f;
    }
    __syncthreads();
    if (b >= a.B) return;
    float *line = sline_all[wi];          // [0..47] gM (r,z,n) | [48..95] gk_h (r,z,nh) | [96..111] qdh
    const bool act = lane < H;
    const int j = act ? lane : 0;
    int cwo0, cwo1;
    const float wo0 = tq_f(qow, __ldg(P + L.oWo + j), cwo0), wo1 = tq_f(qow, __ldg(P + L.oWo + H + j), cwo1);
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
    const float *srow = a.saved + (size_t)b * T * TQ_ROW;
    const float2 *go2 = a.gout ? reinterpret_cast<const float2 *>(a.gout) + (size_t)b * T : nullptr;
    const float2 *oi2 = a.out_in ? reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * T : nullptr;
    const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
    float gWx[18], gWh[3 * TQ_H], gWo0 = 0.f, gWo1 = 0.f;
#pragma unroll
    for (int i = 0; i < 18; ++i) gWx[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 3 * TQ_H; ++i) gWh[i] = 0.f;
    float gH = 0.f, gMr = 0.f, gMz = 0.f, gMn = 0.f, gMnh = 0.f, ghp = 0.f, gxp = 0.f;     // gxp: lane k < 6 carries the pending term of feature k
    for (int t = T - 1; t >= 0; --t) {
        const float *row = srow + (size_t)t * TQ_ROW;
        const int lj = lane < TQ_H ? lane : 0;
        const float qdh = row[lj], r = row[TQ_H + lj], z = row[2 * TQ_H + lj], n = row[3 * TQ_H + lj], sr = row[4 * TQ_H + lj], sz = row[5 * TQ_H + lj],
                    tn = row[6 * TQ_H + lj], mn = row[7 * TQ_H + lj], omz = row[8 * TQ_H + lj], hq = row[10 * TQ_H + lj];
        const int fl = act ? __float_as_int(row[11 * TQ_H + lj]) : 0;
        const float hprev = (act && t > 0) ? row[9 * TQ_H + lj - TQ_ROW] : 0.f;
        float qdx[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) qdx[k] = row[12 * TQ_H + k];
        const unsigned cdxb = (unsigned)__float_as_int(row[12 * TQ_H + 6]), mx = (unsigned)__float_as_int(row[TQ_ROW - 2]), mh = (unsigned)__float_as_int(row[TQ_ROW - 1]);
        float2 go;
        if (go2) go = __ldg(go2 + t);
        else { const float2 o = __ldg(oi2 + t); const float2 y = y2.ld(t); go = make_float2(gs * (o.x - y.x), gs * (o.y - y.y)); }
        if (act) { if (cwo0) gWo0 = fmaf(go.x, hq, gWo0); if (cwo1) gWo1 = fmaf(go.y, hq, gWo1); }
        float g = gH + (((fl >> 9) & 1) ? fmaf(wo0, go.x, wo1 * go.y) : 0.f);
        g = ((fl >> 8) & 1) ? g : 0.f;
        const float gm2 = ((fl >> 7) & 1) ? g : 0.f, gm3 = ((fl >> 6) & 1) ? g : 0.f;
        float gz = gm2 * hprev;
        float ghprev = gm2 * z;
        const float gomz = ((fl >> 5) & 1) ? gm3 * n : 0.f;
        gz -= gomz;
        const float gn = ((fl >> 4) & 1) ? gm3 * omz : 0.f;
        const float gan = ((fl >> 3) & 1) ? gn * (1.f - tn * tn) : 0.f;
        const float gm1 = ((fl >> 2) & 1) ? gan : 0.f;
        const float gr = (fl & 1) ? gm1 * mn : 0.f;
        gMr += gr * sr * (1.f - sr);
        gMz += (((fl >> 1) & 1) ? gz : 0.f) * sz * (1.f - sz);
        gMn += gan;
        gMnh += gm1 * r;
        if (lane < TQ_H) {
            line[lane] = act ? gMr : 0.f; line[TQ_H + lane] = act ? gMz : 0.f; line[2 * TQ_H + lane] = act ? gMn : 0.f;
            line[3 * TQ_H + lane] = act ? gMr : 0.f; line[4 * TQ_H + lane] = act ? gMz : 0.f; line[5 * TQ_H + lane] = act ? gMnh : 0.f;
            line[6 * TQ_H + lane] = act ? qdh : 0.f;
        }
        __syncwarp();
        // dL/d(q dh)[q] = sum_{g,j} Whq[g][j][q] gk_h[g][j]   (lane q);   dL/d(q dx)[q] = sum_{g,j} Wxq[g][j][q] gM[g][j]   (lanes q < 6)
        float gdh = 0.f, gdx = 0.f;
        if (lane < TQ_H) {
#pragma unroll 4
            for (int i = 0; i < 3 * TQ_H; ++i) gdh = fmaf(sWh[i * TQ_H + lane], line[3 * TQ_H + i], gdh);
        }
        if (lane < 6) {
#pragma unroll 4
            for (int i = 0; i < 3 * TQ_H; ++i) gdx = fmaf(sWx[i * 6 + lane], line[i], gdx);
        }
        if (!((fl >> 10) & 1)) gdh = 0.f;
        if (!((cdxb >> (lane < 6 ? lane : 0)) & 1u)) gdx = 0.f;
        // weight gradients (this lane's rows)
        if (dw) {
#pragma unroll
            for (int k = 0; k < 6; ++k) { gWx[k] = fmaf(gMr, qdx[k], gWx[k]); gWx[6 + k] = fmaf(gMz, qdx[k], gWx[6 + k]); gWx[12 + k] = fmaf(gMn, qdx[k], gWx[12 + k]); }
#pragma unroll
            for (int k4 = 0; k4 < TQ_H / 4; ++k4) {
                const float4 v = *reinterpret_cast<const float4 *>(line + 6 * TQ_H + 4 * k4);
                const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = 4 * k4 + i;
                    gWh[k] = fmaf(gMr, e[i], gWh[k]); gWh[TQ_H + k] = fmaf(gMz, e[i], gWh[TQ_H + k]); gWh[2 * TQ_H + k] = fmaf(gMnh, e[i], gWh[2 * TQ_H + k]);
                }
            }
        }
        // through the thresholds
        if (lane < 6) {
            float gfk = 0.f;
            if ((mx >> lane) & 1u) { gfk = gxp + gdx; gxp = -gdx; }
            gf_out[((size_t)b * T + t) * 8 + lane] = gfk;
        }
        if (act && ((mh >> lane) & 1u)) { ghprev += ghp + gdh; ghp = -gdh; }
        gH = act ? ghprev : 0.f;
        __syncwarp();
    }
    if (dw && a.partials) {
        float *prt = a.partials + (size_t)b * L.P;
        for (int i = lane; i < L.P; i += 32) prt[i] = 0.f;          // scales (zero gradient) and the TCN slots (tresq_post_kernel adds to them)
        __syncwarp();
        if (act) {
#pragma unroll
            for (int g = 0; g < 3; ++g) {
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    int in; tq_f(qxw, __ldg(P + L.oWx + (g * H + j) * 6 + k), in);
                    prt[L.oWx + (g * H + j) * 6 + k] = in ? gWx[g * 6 + k] : 0.f;
                }
#pragma unroll
                for (int k = 0; k < TQ_H; ++k)
                    if (k < H) {
                        int in; tq_f(qhw, __ldg(P + L.oWh + (g * H + j) * H + k), in);
                        prt[L.oWh + (g * H + j) * H + k] = in ? gWh[g * TQ_H + k] : 0.f;
                    }
            }
            prt[L.oWo + j] = gWo0; prt[L.oWo + H + j] = gWo1;
        }
    }
}

// ================================================================ backward: time-parallel rest, one CTA per sequence
// dL/dx[t] = feature Jacobian of step t (I, Q, |x|, |x|^3)  +  the next-sample features of step t-1 (wrap)  +  the transposed TCN;  TCN weight gradients.
__global__ void __launch_bounds__(128) tresq_post_kernel(GruArgs a, const float *__restrict__ gf, int dw) {
    pdl_enter();
    const TqLayout L(a.H);
    const int T = a.T, tid = threadIdx.x, b = blockIdx.x;
    __shared__ float sred[4][24];
    float w0[18], w2[6];
#pragma unroll
    for (int i = 0; i < 18; ++i) w0[i] = __ldg(a.params + L.ow0 + i);
#pragma unroll
    for (int i = 0; i < 6; ++i) w2[i] = __ldg(a.params + L.ow2 + i);
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
    const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
    const float2 *go2 = a.gout ? reinterpret_cast<const float2 *>(a.gout) + (size_t)b * T : nullptr;
    const float2 *oi2 = a.out_in ? reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * T : nullptr;
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
    auto go_at = [&](int t) {
        if (go2) return __ldg(go2 + t);
        const float2 o = __ldg(oi2 + t); const float2 y = y2.ld(t);
        return make_float2(gs * (o.x - y.x), gs * (o.y - y.y));
    };
    // gradient w.r.t. the first TCN convolution's outputs at position t
    auto gc1_at = [&](int t, float (&gc)[3], float (&cv)[5], float2 &go) {
        tq_tcn(x2, w0, w2, t, T, cv);
        go = go_at(t);
        const float g20 = go.x * tq_hsw_grad(cv[3]), g21 = go.y * tq_hsw_grad(cv[4]);
#pragma unroll
        for (int co = 0; co < 3; ++co) gc[co] = fmaf(w2[co], g20, w2[3 + co] * g21) * tq_hsw_grad(cv[co]);
    };
    float gw[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) gw[i] = 0.f;
    for (int t = tid; t < T; t += blockDim.x) {
        float gc[3], cv[5];
        float2 go;
        gc1_at(t, gc, cv, go);
        if (dw) {
            const float g20 = go.x * tq_hsw_grad(cv[3]), g21 = go.y * tq_hsw_grad(cv[4]);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) { gw[18 + ch] = fmaf(g20, tq_hsw(cv[ch]), gw[18 + ch]); gw[21 + ch] = fmaf(g21, tq_hsw(cv[ch]), gw[21 + ch]); }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int tt = t + (k - 1) * 16;
                if (tt < 0 || tt >= T) continue;
                const float2 v = x2.ld(tt);
#pragma unroll
                for (int co = 0; co < 3; ++co) { gw[(co * 2) * 3 + k] = fmaf(gc[co], v.x, gw[(co * 2) * 3 + k]); gw[(co * 2 + 1) * 3 + k] = fmaf(gc[co], v.y, gw[(co * 2 + 1) * 3 + k]); }
            }
        }
        if (a.need_dx && a.gx) {
            const float *g0 = gf + ((size_t)b * T + t) * 8;
            const float *gm = gf + ((size_t)b * T + (t > 0 ? t - 1 : T - 1)) * 8;      // step t-1 saw this sample as its "next" sample (wrap)
            float gfe[8] = {g0[0], g0[1], g0[2], g0[3], 0.f, 0.f, 0.f, 0.f};
            const float2 v = x2.ld(t);
            float gi, gq;
            features_bwd<FM_TRES6>(v.x, v.y, gfe, gi, gq);
            gi += gm[4]; gq += gm[5];
            // transposed first convolution: position t feeds cv[.][t'] with t' = t - (k-1) 16
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int tp = t - (k - 1) * 16;
                if (tp < 0 || tp >= T) continue;
                float gcp[3], cvp[5];
                float2 gop;
                if (tp == t) { gcp[0] = gc[0]; gcp[1] = gc[1]; gcp[2] = gc[2]; } else gc1_at(tp, gcp, cvp, gop);
#pragma unroll
                for (int co = 0; co < 3; ++co) { gi = fmaf(w0[(co * 2) * 3 + k], gcp[co], gi); gq = fmaf(w0[(co * 2 + 1) * 3 + k], gcp[co], gq); }
            }
            reinterpret_cast<float2 *>(a.gx)[(size_t)b * T + t] = make_float2(gi, gq);
        }
    }
    if (dw && a.partials) {
#pragma unroll
        for (int i = 0; i < 24; ++i) {
            const float s = warp_sum(gw[i]);
            if ((tid & 31) == 0) sred[tid >> 5][i] = s;
        }
        __syncthreads();
        if (tid < 24) a.partials[(size_t)b * L.P + L.ow0 + tid] = sred[0][tid] + sred[1][tid] + sred[2][tid] + sred[3][tid];
    }
}

// ================================================================ host
int64_t tresq_nparams(int H) { return TqLayout(H).P; }
int64_t tresq_saved_floats(int B, int T) { return (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1) * TQ_ROW; }
// workspace = gradient partials [B][P] (4-aligned) | GF [B][T][8]
int64_t tresq_workspace_floats(int B, int T, int H) {
    const int64_t b1 = B > 0 ? B : 1;
    return ((b1 * TqLayout(H).P + 3) & ~(int64_t)3) + b1 * (T > 0 ? T : 1) * 8 + 4;
}

int tresq_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out) {
    if (a.H < 1 || a.H > TQ_H) { set_error("fake-quantised TRes-DeltaGRU: hidden_size %d outside 1..%d", a.H, TQ_H); return -1; }
    const int cgrid = (a.B + 3) / 4;
    if (dir == 0) {
        if (a.save && !a.saved) { set_error("fake-quantised TRes-DeltaGRU: ODPD_F_SAVE without saved"); return -1; }
        launch_pdl(tresq_fwd_kernel, dim3(cgrid), dim3(128), 0, st, a);
        return check_launch("tresq_fwd_kernel");
    }
    if (!a.partials || !a.saved) { set_error("fake-quantised TRes-DeltaGRU backward needs the saved activations and the workspace"); return -1; }
    const TqLayout L(a.H);
    float *gf = a.partials + (((int64_t)a.B * L.P + 3) & ~(int64_t)3);
    launch_pdl(tresq_bwd_kernel, dim3(cgrid), dim3(128), 0, st, a, gf, dw ? 1 : 0);
    launch_pdl(tresq_post_kernel, dim3(a.B), dim3(128), 0, st, a, (const float *)gf, dw ? 1 : 0);
    if (rows_out) *rows_out = a.B;
    return check_launch("tresq backward");
}

}  // namespace odpd

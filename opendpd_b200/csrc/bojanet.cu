// bojanet.cu — BOJANET backbone: FIR front end + vector demodulator + JANET-style recurrence + phase rotation.
//
// Replaces (reference, file:line): backbones/bojanet.py:5-106 —
//   windows of 16 samples, zero before the frame (:73-77: tap m of window t is sample t+m-15);
//   complex FIR with 6 real-weight filter pairs (:80-83):  I_fir = fir_I(I) - fir_Q(Q),  Q_fir = fir_Q(I) + fir_I(Q);
//   vector demodulator (:30-39):  mag = sqrt(I_fir^2+Q_fir^2) + 1e-8, mag^2, sin = Q_fir/mag, cos = I_fir/mag;  L = [mag(6) | mag^2(6)] (:85-86);
//   recurrence (:87-94):  f = sigmoid(W_fi L + b + W_fh h),  g = tanh(W_gi L + b + W_gh h),  h = f h + (1-f) g,   h_0 = 0;
//   phase rotation (:41-52): unit j is rotated by filter j mod 6 (the three concatenation cases of pr_block; the reference itself fails
//   for hidden_size > 18), I_rot = h cos, Q_rot = h sin;  out_I = W_out_I(I_rot) - W_out_Q(Q_rot),  out_Q = W_out_Q(Q_rot) + W_out_I(I_rot) (:100-102).
//
// Only the f/g recurrence is serial.  Everything else runs one thread per timestep:
//   forward   front (FIR, demodulator, input projections XP) -> chain (one warp per sequence: lane j owns unit j, both weight rows in
//             registers, h broadcast by shuffles) -> head (rotation, output, squared error)
//   backward  head_bwd (dL/dh, dL/dsin, dL/dcos, head weight gradients) -> chain_bwd (reverse recurrence, writes the gate gradients) ->
//             front_bwd (through the demodulator to dL/dFIR; every weight gradient that is an outer product over time: each 64-step tile
//             parks its factors in shared memory and the CTA's threads own one parameter each) -> dx (transposed FIR, a gather)
// Gradient partials: one row per CTA of the (fixed) time-parallel grid, reduced in order by reduce_partials_kernel.
//
// Flat parameter layout (named_parameters() order): fir_I.weight(6,16) fir_Q.weight(6,16) W_fi.weight(H,12) W_fi.bias(H) W_fh.weight(H,H)
// W_gi.weight(H,12) W_gi.bias(H) W_gh.weight(H,H) W_out_I.weight(1,H) W_out_I.bias(1) W_out_Q.weight(1,H) W_out_Q.bias(1)  = 2H^2 + 28H + 194.
#include <mutex>
#include "cells.h"
#include "chunking.cuh"

namespace odpd {

static constexpr int BJ_TT = 64;      // timesteps per tile = threads per CTA of the time-parallel kernels
static constexpr int BJ_P = 6;        // FIR filter pairs / vector-demodulator units
static constexpr int BJ_M = 16;       // window
static constexpr int BJ_HMAX = 18;    // 3 * BJ_P: the reference's pr_block covers nothing wider
static constexpr float BJ_EPS = 1e-8f;

struct BjLayout {
    int H, oFI, oFQ, oWfi, obfi, oWfh, oWgi, obgi, oWgh, oWoI, oboI, oWoQ, oboQ, P;
    __host__ __device__ explicit BjLayout(int h) {
        H = h; oFI = 0; oFQ = 96; oWfi = 192; obfi = oWfi + 12 * h; oWfh = obfi + h; oWgi = oWfh + h * h; obgi = oWgi + 12 * h;
        oWgh = obgi + h; oWoI = oWgh + h * h; oboI = oWoI + h; oWoQ = oboI + 1; oboQ = oWoQ + h; P = oboQ + 1;
    }
};

// saved:  FR [B][T][24] = mag | mag^2 | sin | cos      XP [B][T][2H] = f-gate | g-gate input projection      ACT [B][T][3H] = f | g | h
struct BjBufs { float *fr, *xp, *act, *dh, *dsc, *gb, *dfir, *partials; };

__device__ __forceinline__ void bj_tile(int tile, int nts, int T, int tid, int &b, int &t) {
    b = tile / nts;
    t = (tile - b * nts) * BJ_TT + tid;
    (void)T;
}

// ================================================================ forward: FIR + demodulator + input projections, one thread per timestep
__global__ void __launch_bounds__(BJ_TT) boja_front_kernel(GruArgs a, BjBufs u, int nts, int ntiles) {
    pdl_enter();
    const BjLayout L(a.H);
    const int H = a.H, T = a.T, tid = threadIdx.x;
    extern __shared__ __align__(16) float bsm[];
    float *sp = bsm;                       // fir_I | fir_Q | W_fi | b_fi | (W_fh) | W_gi | b_gi  — the flat block up to oWgh
    for (int i = tid; i < L.oWgh; i += BJ_TT) sp[i] = __ldg(a.params + i);
    __syncthreads();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        bj_tile(tile, nts, T, tid, b, t);
        if (t >= T) continue;
        const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
        float fi[BJ_P], fq[BJ_P];
#pragma unroll
        for (int p = 0; p < BJ_P; ++p) { fi[p] = 0.f; fq[p] = 0.f; }
#pragma unroll
        for (int m = 0; m < BJ_M; ++m) {
            const int s = t + m - (BJ_M - 1);
            if (s < 0) continue;
            const float2 v = x2.ld(s);
#pragma unroll
            for (int p = 0; p < BJ_P; ++p) {
                const float wi = sp[L.oFI + p * BJ_M + m], wq = sp[L.oFQ + p * BJ_M + m];
                fi[p] = fmaf(wi, v.x, fmaf(-wq, v.y, fi[p]));
                fq[p] = fmaf(wq, v.x, fmaf(wi, v.y, fq[p]));
            }
        }
        float Lv[2 * BJ_P];
        float *fr = u.fr + ((size_t)b * T + t) * 24;
#pragma unroll
        for (int p = 0; p < BJ_P; ++p) {
            const float mag = sqrtf(fmaf(fi[p], fi[p], fq[p] * fq[p])) + BJ_EPS;
            Lv[p] = mag; Lv[BJ_P + p] = mag * mag;
            fr[p] = mag; fr[BJ_P + p] = mag * mag; fr[2 * BJ_P + p] = fq[p] / mag; fr[3 * BJ_P + p] = fi[p] / mag;
        }
        float *xp = u.xp + ((size_t)b * T + t) * 2 * H;
        for (int j = 0; j < H; ++j) {
            float af = sp[L.obfi + j], ag = sp[L.obgi + j];
#pragma unroll
            for (int k = 0; k < 2 * BJ_P; ++k) { af = fmaf(sp[L.oWfi + j * 12 + k], Lv[k], af); ag = fmaf(sp[L.oWgi + j * 12 + k], Lv[k], ag); }
            xp[j] = af; xp[H + j] = ag;
        }
    }
}

// ================================================================ forward: the recurrence, one warp per sequence
__global__ void __launch_bounds__(128) boja_chain_fwd_kernel(GruArgs a, BjBufs u) {
    pdl_enter();
    const BjLayout L(a.H);
    const int H = a.H, T = a.T, lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= a.B) return;
    const bool act = lane < H;
    const int j = act ? lane : 0;
    float wf[BJ_HMAX], wg[BJ_HMAX];
#pragma unroll
    for (int k = 0; k < BJ_HMAX; ++k) {
        wf[k] = (act && k < H) ? __ldg(a.params + L.oWfh + j * H + k) : 0.f;
        wg[k] = (act && k < H) ? __ldg(a.params + L.oWgh + j * H + k) : 0.f;
    }
    const float *xp = u.xp + (size_t)b * T * 2 * H + j;
    float *arow = u.act + (size_t)b * T * 3 * H + j;
    float h = 0.f;
    float qf[8], qg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        qf[i] = (i < T && act) ? __ldg(xp + (size_t)i * 2 * H) : 0.f;
        qg[i] = (i < T && act) ? __ldg(xp + (size_t)i * 2 * H + H) : 0.f;
    }
    for (int t0 = 0; t0 < T; t0 += 8) {
        float nf[8], ng[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int tt = t0 + 8 + i;
            nf[i] = (tt < T && act) ? __ldg(xp + (size_t)tt * 2 * H) : 0.f;
            ng[i] = (tt < T && act) ? __ldg(xp + (size_t)tt * 2 * H + H) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = t0 + i;
            if (t < T) {
                float f0 = qf[i], f1 = 0.f, g0 = qg[i], g1 = 0.f;
#pragma unroll
                for (int k = 0; k < BJ_HMAX; k += 2) {
                    const float h0 = __shfl_sync(ODPD_FULL, h, k), h1 = __shfl_sync(ODPD_FULL, h, k + 1);
                    f0 = fmaf(wf[k], h0, f0); f1 = fmaf(wf[k + 1], h1, f1);
                    g0 = fmaf(wg[k], h0, g0); g1 = fmaf(wg[k + 1], h1, g1);
                }
                const float f = sigmoidf_(f0 + f1), g = tanhf_(g0 + g1);
                h = act ? fmaf(f, h - g, g) : 0.f;                       // f h + (1-f) g
                if (act) {
                    float *row = arow + (size_t)t * 3 * H;
                    row[0] = f; row[H] = g; row[2 * H] = h;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { qf[i] = nf[i]; qg[i] = ng[i]; }
    }
}

// ================================================================ forward: phase rotation + output + squared error, one thread per timestep
__global__ void __launch_bounds__(BJ_TT) boja_head_fwd_kernel(GruArgs a, BjBufs u, int nts, int ntiles) {
    pdl_enter();
    const BjLayout L(a.H);
    const int H = a.H, T = a.T, tid = threadIdx.x;
    __shared__ float sWI[BJ_HMAX], sWQ[BJ_HMAX], sred[BJ_TT / 32];
    if (tid < H) { sWI[tid] = __ldg(a.params + L.oWoI + tid); sWQ[tid] = __ldg(a.params + L.oWoQ + tid); }
    const float bI = __ldg(a.params + L.oboI), bQ = __ldg(a.params + L.oboQ);
    __syncthreads();
    float lsum = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        bj_tile(tile, nts, T, tid, b, t);
        if (t >= T) continue;
        const float *fr = u.fr + ((size_t)b * T + t) * 24;
        const float *hrow = u.act + ((size_t)b * T + t) * 3 * H + 2 * H;
        float sn[BJ_P], cs[BJ_P];
#pragma unroll
        for (int p = 0; p < BJ_P; ++p) { sn[p] = fr[2 * BJ_P + p]; cs[p] = fr[3 * BJ_P + p]; }
        float pa = bI, pq = bQ;
#pragma unroll
        for (int j = 0; j < BJ_HMAX; ++j) {
            if (j < H) {
                const float hv = hrow[j];
                pa = fmaf(sWI[j], hv * cs[j % BJ_P], pa);
                pq = fmaf(sWQ[j], hv * sn[j % BJ_P], pq);
            }
        }
        const float o0 = pa - pq, o1 = pq + pa;
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

// ================================================================ backward: head.  dL/dh, dL/dsin, dL/dcos per step; head weight gradients
template <bool DW>
__global__ void __launch_bounds__(BJ_TT) boja_head_bwd_kernel(GruArgs a, BjBufs u, int nts, int ntiles) {
    pdl_enter();
    const BjLayout L(a.H);
    const int H = a.H, T = a.T, tid = threadIdx.x;
    __shared__ float sWI[BJ_HMAX], sWQ[BJ_HMAX], sred[2][2 * BJ_HMAX + 2];
    if (tid < H) { sWI[tid] = __ldg(a.params + L.oWoI + tid); sWQ[tid] = __ldg(a.params + L.oWoQ + tid); }
    __syncthreads();
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
    float gwI[BJ_HMAX], gwQ[BJ_HMAX], gbI = 0.f, gbQ = 0.f;
#pragma unroll
    for (int j = 0; j < BJ_HMAX; ++j) { gwI[j] = 0.f; gwQ[j] = 0.f; }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        bj_tile(tile, nts, T, tid, b, t);
        if (t >= T) continue;
        float2 go;
        if (a.gout) go = __ldg(reinterpret_cast<const float2 *>(a.gout) + (size_t)b * T + t);
        else {
            const float2 o = __ldg(reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * T + t);
            const float2 y = iq_row(a.target, a.target_bf16, a.target_starts, b, T).ld(t);
            go = make_float2(gs * (o.x - y.x), gs * (o.y - y.y));
        }
        const float da = go.x + go.y, dq = go.y - go.x;      // out_I = A - Q', out_Q = Q' + A  with A = W_out_I(I_rot), Q' = W_out_Q(Q_rot)
        const float *fr = u.fr + ((size_t)b * T + t) * 24;
        const float *hrow = u.act + ((size_t)b * T + t) * 3 * H + 2 * H;
        float sn[BJ_P], cs[BJ_P], dsn[BJ_P], dcs[BJ_P];
#pragma unroll
        for (int p = 0; p < BJ_P; ++p) { sn[p] = fr[2 * BJ_P + p]; cs[p] = fr[3 * BJ_P + p]; dsn[p] = 0.f; dcs[p] = 0.f; }
        float *dh = u.dh + ((size_t)b * T + t) * H;
#pragma unroll
        for (int j = 0; j < BJ_HMAX; ++j) {
            if (j < H) {
                const float hv = hrow[j], dI = da * sWI[j], dQ = dq * sWQ[j];
                dh[j] = fmaf(dI, cs[j % BJ_P], dQ * sn[j % BJ_P]);
                dcs[j % BJ_P] = fmaf(dI, hv, dcs[j % BJ_P]);
                dsn[j % BJ_P] = fmaf(dQ, hv, dsn[j % BJ_P]);
                if constexpr (DW) { gwI[j] = fmaf(da, hv * cs[j % BJ_P], gwI[j]); gwQ[j] = fmaf(dq, hv * sn[j % BJ_P], gwQ[j]); }
            }
        }
        if constexpr (DW) { gbI += da; gbQ += dq; }
        float *dsc = u.dsc + ((size_t)b * T + t) * 12;
#pragma unroll
        for (int p = 0; p < BJ_P; ++p) { dsc[p] = dsn[p]; dsc[BJ_P + p] = dcs[p]; }
    }
    if constexpr (DW) {
        // CTA reduction of the 2H+2 head gradients (two warps)
        const int wi = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int j = 0; j < BJ_HMAX; ++j) {
            const float s0 = warp_sum(gwI[j]), s1 = warp_sum(gwQ[j]);
            if (lane == 0) { sred[wi][j] = s0; sred[wi][BJ_HMAX + j] = s1; }
        }
        const float s2 = warp_sum(gbI), s3 = warp_sum(gbQ);
        if (lane == 0) { sred[wi][2 * BJ_HMAX] = s2; sred[wi][2 * BJ_HMAX + 1] = s3; }
        __syncthreads();
        float *prt = u.partials + (size_t)blockIdx.x * L.P;
        if (tid < H) { prt[L.oWoI + tid] = sred[0][tid] + sred[1][tid]; prt[L.oWoQ + tid] = sred[0][BJ_HMAX + tid] + sred[1][BJ_HMAX + tid]; }
        if (tid == 0) { prt[L.oboI] = sred[0][2 * BJ_HMAX] + sred[1][2 * BJ_HMAX]; prt[L.oboQ] = sred[0][2 * BJ_HMAX + 1] + sred[1][2 * BJ_HMAX + 1]; }
    }
}

// ================================================================ backward: reverse recurrence, one warp per sequence
//   dh_t = DH_t + rec;  df = dh (h_{t-1} - g), dg = dh (1 - f);  af = df f (1-f), ag = dg (1 - g^2);
//   rec  = dh f + W_fh^T af + W_gh^T ag.        G[b][t] = af | ag
__global__ void __launch_bounds__(128) boja_chain_bwd_kernel(GruArgs a, BjBufs u) {
    pdl_enter();
    const BjLayout L(a.H);
    const int H = a.H, T = a.T, lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= a.B) return;
    const bool act = lane < H;
    const int k = act ? lane : 0;
    float cf[BJ_HMAX], cg[BJ_HMAX];         // column k of W_fh / W_gh
#pragma unroll
    for (int j = 0; j < BJ_HMAX; ++j) {
        cf[j] = (act && j < H) ? __ldg(a.params + L.oWfh + j * H + k) : 0.f;
        cg[j] = (act && j < H) ? __ldg(a.params + L.oWgh + j * H + k) : 0.f;
    }
    const float *arow = u.act + (size_t)b * T * 3 * H + k;
    const float *dhrow = u.dh + (size_t)b * T * H + k;
    float *grow = u.gb + (size_t)b * T * 2 * H + k;
    auto load = [&](int t, float *d) {
        d[0] = d[1] = d[2] = d[3] = 0.f;
        if (act && t >= 0) {
            const float *row = arow + (size_t)t * 3 * H;
            d[0] = __ldg(dhrow + (size_t)t * H);
            d[1] = __ldg(row); d[2] = __ldg(row + H);
            d[3] = t > 0 ? __ldg(row - 3 * H + 2 * H) : 0.f;       // h_{t-1}
        }
    };
    float rec = 0.f;
    float cq[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) load(T - 1 - i, cq[i]);
    for (int t0 = T - 1; t0 >= 0; t0 -= 4) {
        float nq[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) load(t0 - 4 - i, nq[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = t0 - i;
            if (t >= 0) {
                const float dh = cq[i][0] + rec, f = cq[i][1], g = cq[i][2], hp = cq[i][3];
                const float af = dh * (hp - g) * f * (1.f - f), ag = dh * (1.f - f) * (1.f - g * g);
                if (act) { grow[(size_t)t * 2 * H] = af; grow[(size_t)t * 2 * H + H] = ag; }
                float r0 = dh * f, r1 = 0.f;
#pragma unroll
                for (int j = 0; j < BJ_HMAX; j += 2) {
                    const float f0 = __shfl_sync(ODPD_FULL, af, j), f1 = __shfl_sync(ODPD_FULL, af, j + 1);
                    const float g0 = __shfl_sync(ODPD_FULL, ag, j), g1 = __shfl_sync(ODPD_FULL, ag, j + 1);
                    r0 = fmaf(cf[j], f0, fmaf(cg[j], g0, r0));
                    r1 = fmaf(cf[j + 1], f1, fmaf(cg[j + 1], g1, r1));
                }
                rec = act ? r0 + r1 : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) cq[i][q] = nq[i][q];
    }
}

// ================================================================ backward: front.  Gate gradients -> dL/dL -> demodulator -> dL/dFIR; weight gradients
// per-tile shared factors (odd pitches):  G[t][2H] | Lv[t][12] | hp[t][H] | dfir[t][12] | x[t + 15 halo][2]
template <bool DW>
__global__ void __launch_bounds__(BJ_TT) boja_front_bwd_kernel(GruArgs a, BjBufs u, int nts, int ntiles) {
    pdl_enter();
    const BjLayout L(a.H);
    const int H = a.H, T = a.T, tid = threadIdx.x;
    const int GP = (2 * H) | 1, HPi = H | 1;
    extern __shared__ __align__(16) float bsm[];
    float *sWfi = bsm;                          // [H][12]
    float *sWgi = sWfi + 12 * BJ_HMAX;          // [H][12]
    float *sG = sWgi + 12 * BJ_HMAX;            // [64][GP]
    float *sL = sG + BJ_TT * GP;                // [64][13]
    float *sH = sL + BJ_TT * 13;                // [64][HPi]
    float *sD = sH + BJ_TT * HPi;               // [64][13]
    float *sX = sD + BJ_TT * 13;                // [79][2]
    for (int i = tid; i < 12 * H; i += BJ_TT) { sWfi[i] = __ldg(a.params + L.oWfi + i); sWgi[i] = __ldg(a.params + L.oWgi + i); }
    __syncthreads();
    float *prt = (DW && u.partials) ? u.partials + (size_t)blockIdx.x * L.P : nullptr;
    bool first = true;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        bj_tile(tile, nts, T, tid, b, t);
        const bool valid = t < T;
        const int tbase = t - tid;
        float dfi[BJ_P], dfq[BJ_P], Lv[2 * BJ_P];
#pragma unroll
        for (int p = 0; p < BJ_P; ++p) { dfi[p] = 0.f; dfq[p] = 0.f; Lv[p] = 0.f; Lv[BJ_P + p] = 0.f; }
        if (valid) {
            const float *fr = u.fr + ((size_t)b * T + t) * 24;
            const float *g = u.gb + ((size_t)b * T + t) * 2 * H;
            const float *dsc = u.dsc + ((size_t)b * T + t) * 12;
            float dL[2 * BJ_P];
#pragma unroll
            for (int q = 0; q < 2 * BJ_P; ++q) dL[q] = 0.f;
            for (int j = 0; j < H; ++j) {
                const float af = g[j], ag = g[H + j];
                if constexpr (DW) { sG[tid * GP + j] = af; sG[tid * GP + H + j] = ag; }
#pragma unroll
                for (int q = 0; q < 2 * BJ_P; ++q) dL[q] = fmaf(af, sWfi[j * 12 + q], fmaf(ag, sWgi[j * 12 + q], dL[q]));
            }
#pragma unroll
            for (int p = 0; p < BJ_P; ++p) {
                const float mag = fr[p], sn = fr[2 * BJ_P + p], cs = fr[3 * BJ_P + p];
                Lv[p] = mag; Lv[BJ_P + p] = fr[BJ_P + p];
                const float dsn = dsc[p], dcs = dsc[BJ_P + p];
                // mag = sqrt(I^2+Q^2) + eps, sin = Q/mag, cos = I/mag:  I = cos*mag, Q = sin*mag, sqrt(.) = mag - eps
                const float dmag = fmaf(2.f * mag, dL[BJ_P + p], dL[p]) - (dsn * sn + dcs * cs) / mag;
                const float root = mag - BJ_EPS;
                dfi[p] = dcs / mag + dmag * (cs * mag) / root;
                dfq[p] = dsn / mag + dmag * (sn * mag) / root;
            }
            float *df = u.dfir + ((size_t)b * T + t) * 12;
#pragma unroll
            for (int p = 0; p < BJ_P; ++p) { df[p] = dfi[p]; df[BJ_P + p] = dfq[p]; }
        } else if constexpr (DW) {
            for (int j = 0; j < 2 * H; ++j) sG[tid * GP + j] = 0.f;
        }
        if constexpr (DW) {
#pragma unroll
            for (int q = 0; q < 2 * BJ_P; ++q) sL[tid * 13 + q] = Lv[q];
#pragma unroll
            for (int p = 0; p < BJ_P; ++p) { sD[tid * 13 + p] = dfi[p]; sD[tid * 13 + BJ_P + p] = dfq[p]; }
            const float *hprev = u.act + ((size_t)b * T + t - 1) * 3 * H + 2 * H;
            for (int j = 0; j < H; ++j) sH[tid * HPi + j] = (valid && t > 0) ? hprev[j] : 0.f;
            // samples tbase-15 .. tbase+63 of the sequence (zero outside the frame)
            const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
            for (int i = tid; i < BJ_TT + BJ_M - 1; i += BJ_TT) {
                const int s = tbase - (BJ_M - 1) + i;
                float2 v = make_float2(0.f, 0.f);
                if (s >= 0 && s < T) v = x2.ld(s);
                sX[2 * i] = v.x; sX[2 * i + 1] = v.y;
            }
            __syncthreads();
            if (prt) {
                for (int o = tid; o < L.oWoI; o += BJ_TT) {
                    float s = 0.f;
                    if (o < 192) {                                   // fir_I / fir_Q  [p][m]
                        const bool isq = o >= 96;
                        const int p = (o - (isq ? 96 : 0)) / BJ_M, m = (o - (isq ? 96 : 0)) - p * BJ_M;
                        for (int tt = 0; tt < BJ_TT; ++tt) {
                            const float di = sD[tt * 13 + p], dq = sD[tt * 13 + BJ_P + p], xi = sX[2 * (tt + m)], xq = sX[2 * (tt + m) + 1];
                            s += isq ? fmaf(dq, xi, -di * xq) : fmaf(di, xi, dq * xq);
                        }
                    } else if (o < L.obfi) {                         // W_fi[j][k]
                        const int j = (o - L.oWfi) / 12, kk = (o - L.oWfi) - j * 12;
                        for (int tt = 0; tt < BJ_TT; ++tt) s = fmaf(sG[tt * GP + j], sL[tt * 13 + kk], s);
                    } else if (o < L.oWfh) {                         // b_fi[j]
                        const int j = o - L.obfi;
                        for (int tt = 0; tt < BJ_TT; ++tt) s += sG[tt * GP + j];
                    } else if (o < L.oWgi) {                         // W_fh[j][k]
                        const int j = (o - L.oWfh) / H, kk = (o - L.oWfh) - j * H;
                        for (int tt = 0; tt < BJ_TT; ++tt) s = fmaf(sG[tt * GP + j], sH[tt * HPi + kk], s);
                    } else if (o < L.obgi) {                         // W_gi[j][k]
                        const int j = (o - L.oWgi) / 12, kk = (o - L.oWgi) - j * 12;
                        for (int tt = 0; tt < BJ_TT; ++tt) s = fmaf(sG[tt * GP + H + j], sL[tt * 13 + kk], s);
                    } else if (o < L.oWgh) {                         // b_gi[j]
                        const int j = o - L.obgi;
                        for (int tt = 0; tt < BJ_TT; ++tt) s += sG[tt * GP + H + j];
                    } else {                                         // W_gh[j][k]
                        const int j = (o - L.oWgh) / H, kk = (o - L.oWgh) - j * H;
                        for (int tt = 0; tt < BJ_TT; ++tt) s = fmaf(sG[tt * GP + H + j], sH[tt * HPi + kk], s);
                    }
                    prt[o] = first ? s : prt[o] + s;
                }
            }
            first = false;
            __syncthreads();
        }
    }
    if constexpr (DW) {
        if (prt && first)
            for (int o = tid; o < L.oWoI; o += BJ_TT) prt[o] = 0.f;
    }
}

// ================================================================ backward: transposed FIR.  dL/dx[s] gathers the windows t = s .. s+15 that hold sample s
__global__ void __launch_bounds__(BJ_TT) boja_dx_kernel(GruArgs a, BjBufs u, int nts, int ntiles) {
    pdl_enter();
    const int T = a.T, tid = threadIdx.x;
    __shared__ float sF[192];
    __shared__ float sD[(BJ_TT + BJ_M - 1) * 13];        // dL/dFIR of the windows s0 .. s0+78 (pitch 13)
    for (int i = tid; i < 192; i += BJ_TT) sF[i] = __ldg(a.params + i);
    __syncthreads();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, s;
        bj_tile(tile, nts, T, tid, b, s);
        const int s0 = s - tid;
        for (int i = tid; i < (BJ_TT + BJ_M - 1) * 12; i += BJ_TT) {
            const int r = i / 12, c = i - r * 12, t = s0 + r;
            sD[r * 13 + c] = t < T ? __ldg(u.dfir + ((size_t)b * T + t) * 12 + c) : 0.f;
        }
        __syncthreads();
        if (s < T) {
            float gi = 0.f, gq = 0.f;
#pragma unroll
            for (int m = 0; m < BJ_M; ++m) {
                const float *df = sD + (tid + (BJ_M - 1) - m) * 13;      // window t = s+15-m holds sample s at tap m (zero rows beyond the frame)
#pragma unroll
                for (int p = 0; p < BJ_P; ++p) {
                    const float di = df[p], dq = df[BJ_P + p], wi = sF[p * BJ_M + m], wq = sF[96 + p * BJ_M + m];
                    gi = fmaf(di, wi, fmaf(dq, wq, gi));     // I_fir = wi I - wq Q,  Q_fir = wq I + wi Q
                    gq = fmaf(dq, wi, fmaf(-di, wq, gq));
                }
            }
            reinterpret_cast<float2 *>(a.gx)[(size_t)b * T + s] = make_float2(gi, gq);
        }
        __syncthreads();
    }
}

// ================================================================ host
static int bj_grid(int B, int T) {
    const int64_t tiles = (int64_t)B * ((T + BJ_TT - 1) / BJ_TT);
    const int64_t cap = 8 * (int64_t)num_sms();
    return (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
}
int64_t bojanet_nparams(int H) { return BjLayout(H).P; }
int64_t bojanet_saved_floats(int B, int T, int H) {
    const int64_t bt = (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1);
    return bt * (24 + 5 * H) + 4;
}
// workspace = gradient partials [rows][P] (4-aligned) | DH [B][T][H] | DSC [B][T][12] | G [B][T][2H] | DFIR [B][T][12]
int64_t bojanet_workspace_floats(int B, int T, int H) {
    const int64_t bt = (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1);
    return (((int64_t)bj_grid(B, T) * BjLayout(H).P + 3) & ~(int64_t)3) + bt * (3 * H + 24) + 4;
}

static void bj_ensure_smem(const void *k, size_t bytes) {
    if (bytes <= 48 * 1024) return;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int bojanet_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out) {
    if (a.H < 1 || a.H > BJ_HMAX) {
        set_error("BOJANET: hidden_size %d outside 1..%d (the reference's pr_block, bojanet.py:41-52, covers at most 3 x 6 units)", a.H, BJ_HMAX);
        return -1;
    }
    if (a.T < BJ_M - 1) {
        set_error("BOJANET needs frame_length >= 15 (the reference pads with zeros_like(x[:, -15:]), bojanet.py:75-79; got %d)", a.T);
        return -1;
    }
    const BjLayout L(a.H);
    const int H = a.H, nts = (a.T + BJ_TT - 1) / BJ_TT, ntiles = a.B * nts, grid = bj_grid(a.B, a.T);
    const int64_t bt = (int64_t)a.B * a.T;
    BjBufs u{};
    if (!a.saved) { set_error("BOJANET needs the `saved` buffer (odpd_saved_bytes), also without ODPD_F_SAVE"); return -1; }
    u.fr = a.saved; u.xp = u.fr + bt * 24; u.act = u.xp + bt * 2 * H;
    const int wpc = a.B <= 2 * num_sms() ? 1 : 4 /* few sequences: one chain warp per CTA spreads them over the SMs */, cgrid = (a.B + wpc - 1) / wpc;
    if (dir == 0) {
        const size_t fsm = (size_t)L.oWgh * sizeof(float);
        bj_ensure_smem((const void *)boja_front_kernel, fsm);
        launch_pdl(boja_front_kernel, dim3(grid), dim3(BJ_TT), fsm, st, a, u, nts, ntiles);
        launch_pdl(boja_chain_fwd_kernel, dim3(cgrid), dim3(32 * wpc), 0, st, a, u);
        launch_pdl(boja_head_fwd_kernel, dim3(grid), dim3(BJ_TT), 0, st, a, u, nts, ntiles);
        return check_launch("bojanet forward");
    }
    if (!a.partials) { set_error("BOJANET backward needs the workspace (odpd_bwd_workspace_bytes)"); return -1; }
    const int64_t poff = ((int64_t)grid * L.P + 3) & ~(int64_t)3;
    u.partials = a.partials; u.dh = a.partials + poff; u.dsc = u.dh + bt * H; u.gb = u.dsc + bt * 12; u.dfir = u.gb + bt * 2 * H;
    const size_t bsm = (size_t)(24 * BJ_HMAX + BJ_TT * (((2 * H) | 1) + 13 + (H | 1) + 13) + 2 * (BJ_TT + BJ_M - 1)) * sizeof(float);
    if (dw) {
        launch_pdl(boja_head_bwd_kernel<true>, dim3(grid), dim3(BJ_TT), 0, st, a, u, nts, ntiles);
        launch_pdl(boja_chain_bwd_kernel, dim3(cgrid), dim3(32 * wpc), 0, st, a, u);
        bj_ensure_smem((const void *)boja_front_bwd_kernel<true>, bsm);
        launch_pdl(boja_front_bwd_kernel<true>, dim3(grid), dim3(BJ_TT), bsm, st, a, u, nts, ntiles);
    } else {
        launch_pdl(boja_head_bwd_kernel<false>, dim3(grid), dim3(BJ_TT), 0, st, a, u, nts, ntiles);
        launch_pdl(boja_chain_bwd_kernel, dim3(cgrid), dim3(32 * wpc), 0, st, a, u);
        bj_ensure_smem((const void *)boja_front_bwd_kernel<false>, bsm);
        launch_pdl(boja_front_bwd_kernel<false>, dim3(grid), dim3(BJ_TT), bsm, st, a, u, nts, ntiles);
    }
    if (a.need_dx) {
        if (!a.gx) { set_error("BOJANET backward: ODPD_F_NEED_DX without gx"); return -1; }
        launch_pdl(boja_dx_kernel, dim3(grid), dim3(BJ_TT), 0, st, a, u, nts, ntiles);
    }
    if (rows_out) *rows_out = grid;
    return check_launch("bojanet backward");
}

}  // namespace odpd

// apnrru.cu — APNRRU backbone: phase-normalised recurrent unit, forward / backward (+ fused I/Q MSE).
//
// Replaces (reference, file:line): backbones/apnrru.py:5-135 —
//   16-tap complex FIR with 3 real-weight filter pairs over windows that are zero before the frame (:67-71, :84-87: I_f = fir_I(I) - fir_Q(Q),
//   Q_f = fir_Q(I) + fir_I(Q)) plus the raw sample: 4 complex inputs;  r = conj(x_t)/|x_t| (:74-77) rotates the inputs (:93-95) and, at every
//   step, the complex state h_I + j h_Q (:101-102);  the RRU cell (:22-31) on u = [inputs(8), h_I', h_Q', h_A(3)], hnew = [h_I', h_Q', h_A]:
//   v = sigmoid(C hnew) + Z * tanh(W_h tanh(W_u u + b_u) + b_h);  (v[:H] + j v[H:2H]) is rotated back by conj(r) (:115-119), h_A = v[2H:];
//   out_I = o_I(h_I) - o_Q(h_Q),  out_Q = o_Q(h_Q) + o_I(h_I) (:123-125).
// S = 2H+3 state values; one warp carries one sequence with lane = state value, so hidden_size <= 14 (S <= 31).
//
// Only the cell is serial.  forward: front (one thread per timestep: FIR, r, rotated inputs, their share of W_u) -> chain (one warp per
// sequence: the two rotations are lane-pair shuffles, the two dense layers broadcast their inputs through a shared line; W_u rows / W_h rows
// in registers) -> head (one thread per timestep).  backward: chain_bwd (reverse; columns of W_h / W_u in registers; stores the per-step
// adjoints gv, ga1 and the partial dL/dr) -> front_bwd (one thread per timestep: through the input rotation and r to dL/dx and dL/dFIR;
// every weight gradient as outer products over 64-step tiles held in shared memory, one parameter per thread at a time) -> dx (transposed FIR).
//
// Flat parameter layout (named_parameters() order): fir_I.weight(3,16) fir_Q.weight(3,16) rru.C(1) rru.Z(1,S) rru.W_u.weight(16,S+8) rru.W_u.bias(16)
// rru.W_h.weight(S,16) rru.W_h.bias(S) output_layer_I.weight(1,H) output_layer_Q.weight(1,H)   = 241 + 34 S + 2 H.
#include <mutex>
#include "cells.h"
#include "chunking.cuh"

namespace odpd {

static constexpr int AP_TT = 64;       // timesteps per tile = threads per CTA of the time-parallel kernels
static constexpr int AP_M = 16;        // FIR window
static constexpr int AP_HMAX = 14;
static constexpr int AP_SMAX = 2 * AP_HMAX + 3;   // 31

struct ApLayout {
    int H, S, U, oC, oZ, oWu, obu, oWh, obh, oI, oQ, P;
    __host__ __device__ explicit ApLayout(int h) {
        H = h; S = 2 * h + 3; U = S + 8;
        oC = 96; oZ = 97; oWu = oZ + S; obu = oWu + 16 * U; oWh = obu + 16; obh = oWh + 16 * S; oI = obh + S; oQ = oI + h; P = oQ + h;
    }
    // saved row per step:  hnew(S) | v1(16) | v2(S) | sg(S) | hd(2H)
    __host__ __device__ int row() const { return 3 * S + 16 + 2 * H; }
};
// saved:  FR [B][T][12] = fir_I(3) fir_Q(3) rr ri mag - - -   |  XP [B][T][16]  |  ST [B][T][row]
// workspace: partials (4-aligned) | GXD [B][T][2] (float2 accesses: kept at the aligned front) | GV [B][T][S] | GA1 [B][T][16] | GRR [B][T][2] | DFIR [B][T][6]
struct ApBufs { float *fr, *xp, *st, *gv, *ga1, *grr, *dfir, *gxd, *partials; };

__device__ __forceinline__ void ap_tile(int tile, int nts, int tid, int &b, int &t) {
    b = tile / nts;
    t = (tile - b * nts) * AP_TT + tid;
}

// FIR outputs, r and the 8 rotated inputs of timestep t
__device__ __forceinline__ void ap_inputs(const IqRow &x2, const float *sF, int t, float *fir, float &rr, float &ri, float &mag, float *xin) {
#pragma unroll
    for (int p = 0; p < 6; ++p) fir[p] = 0.f;
#pragma unroll
    for (int m = 0; m < AP_M; ++m) {
        const int s = t + m - (AP_M - 1);
        if (s < 0) continue;
        const float2 v = x2.ld(s);
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const float wi = sF[p * AP_M + m], wq = sF[48 + p * AP_M + m];
            fir[p] = fmaf(wi, v.x, fmaf(-wq, v.y, fir[p]));
            fir[3 + p] = fmaf(wq, v.x, fmaf(wi, v.y, fir[3 + p]));
        }
    }
    const float2 c = x2.ld(t);
    mag = sqrtf(fmaf(c.x, c.x, c.y * c.y));
    rr = c.x / mag; ri = -c.y / mag;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float a = k < 3 ? fir[k] : c.x, bq = k < 3 ? fir[3 + k] : c.y;
        xin[2 * k] = fmaf(rr, a, -ri * bq);
        xin[2 * k + 1] = fmaf(ri, a, rr * bq);
    }
}

// ================================================================ forward: front
__global__ void __launch_bounds__(AP_TT) apn_front_kernel(GruArgs a, ApBufs u, int nts, int ntiles) {
    pdl_enter();
    const ApLayout L(a.H);
    const int T = a.T, tid = threadIdx.x, U = L.U;
    __shared__ float sF[96], sWx[16 * 8], sbu[16];
    for (int i = tid; i < 96; i += AP_TT) sF[i] = __ldg(a.params + i);
    for (int i = tid; i < 128; i += AP_TT) sWx[i] = __ldg(a.params + L.oWu + (i >> 3) * U + (i & 7));
    if (tid < 16) sbu[tid] = __ldg(a.params + L.obu + tid);
    __syncthreads();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        ap_tile(tile, nts, tid, b, t);
        if (t >= T) continue;
        const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
        float fir[6], xin[8], rr, ri, mag;
        ap_inputs(x2, sF, t, fir, rr, ri, mag, xin);
        float *fr = u.fr + ((size_t)b * T + t) * 12;
#pragma unroll
        for (int p = 0; p < 6; ++p) fr[p] = fir[p];
        fr[6] = rr; fr[7] = ri; fr[8] = mag;
        float *xp = u.xp + ((size_t)b * T + t) * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float acc = sbu[i];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc = fmaf(sWx[i * 8 + k], xin[k], acc);
            xp[i] = acc;
        }
    }
}

// ================================================================ forward: the cell, one warp per sequence (lane = state value)
__global__ void __launch_bounds__(128) apn_chain_fwd_kernel(GruArgs a, ApBufs u) {
    pdl_enter();
    const ApLayout L(a.H);
    const int H = a.H, S = L.S, U = L.U, T = a.T, lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + wi;
    __shared__ __align__(16) float sline[4][32 + 16];
    if (b >= a.B) return;
    float *line = sline[wi], *line1 = line + 32;
    const bool act = lane < S;
    const int partner = lane < H ? lane + H : (lane < 2 * H ? lane - H : lane);
    const bool lowI = lane < H, upQ = lane >= H && lane < 2 * H;
    float wu[AP_SMAX], wh[16];
#pragma unroll
    for (int s = 0; s < AP_SMAX; ++s) wu[s] = (lane < 16 && s < S) ? __ldg(a.params + L.oWu + lane * U + 8 + s) : 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) wh[i] = act ? __ldg(a.params + L.oWh + lane * 16 + i) : 0.f;
    const float bh = act ? __ldg(a.params + L.obh + lane) : 0.f, Zs = act ? __ldg(a.params + L.oZ + lane) : 0.f, Cc = __ldg(a.params + L.oC);
    const float *fr = u.fr + (size_t)b * T * 12;
    const float *xp = u.xp + (size_t)b * T * 16 + (lane & 15);
    const int R = L.row();
    float *st = u.st + (size_t)b * T * R;
    float h = 0.f;                                    // lanes < H: h_I, H..2H-1: h_Q, 2H..2H+2: h_A
    // inputs of the next 4 steps are fetched while the current 4 run (a one-step-ahead register prefetch stalls on its own hand-over:
    // ncu showed a third of all stall samples on the MOV that ends the iteration)
    float cx[4], cr[4], ci[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        cx[i] = i < T ? __ldg(xp + (size_t)i * 16) : 0.f;
        cr[i] = i < T ? __ldg(fr + (size_t)i * 12 + 6) : 0.f;
        ci[i] = i < T ? __ldg(fr + (size_t)i * 12 + 7) : 0.f;
    }
    for (int t0 = 0; t0 < T; t0 += 4) {
        float nx[4], nr[4], ni[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int tt = t0 + 4 + i;
            nx[i] = tt < T ? __ldg(xp + (size_t)tt * 16) : 0.f;
            nr[i] = tt < T ? __ldg(fr + (size_t)tt * 12 + 6) : 0.f;
            ni[i] = tt < T ? __ldg(fr + (size_t)tt * 12 + 7) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
        const int t = t0 + i;
        if (t < T) {
        const float xpv = cx[i], rr = cr[i], ri = ci[i];
        // rotate the complex state by r
        const float hp = __shfl_sync(ODPD_FULL, h, partner);
        const float hn = lowI ? fmaf(h, rr, -hp * ri) : (upQ ? fmaf(hp, ri, h * rr) : h);
        line[lane] = act ? hn : 0.f;
        __syncwarp();
        float a0 = xpv, a1 = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 v = reinterpret_cast<const float4 *>(line)[q];
            if (4 * q < AP_SMAX) a0 = fmaf(wu[4 * q], v.x, a0);
            if (4 * q + 1 < AP_SMAX) a1 = fmaf(wu[4 * q + 1], v.y, a1);
            if (4 * q + 2 < AP_SMAX) a0 = fmaf(wu[4 * q + 2], v.z, a0);
            if (4 * q + 3 < AP_SMAX) a1 = fmaf(wu[4 * q + 3], v.w, a1);
        }
        const float v1 = tanhf_(a0 + a1);
        if (lane < 16) line1[lane] = v1;
        __syncwarp();
        float c0 = bh, c1 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = reinterpret_cast<const float4 *>(line1)[q];
            c0 = fmaf(wh[4 * q], v.x, c0); c1 = fmaf(wh[4 * q + 1], v.y, c1); c0 = fmaf(wh[4 * q + 2], v.z, c0); c1 = fmaf(wh[4 * q + 3], v.w, c1);
        }
        const float v2 = tanhf_(c0 + c1), sg = sigmoidf_(Cc * hn);
        const float vv = act ? fmaf(Zs, v2, sg) : 0.f;
        // rotate back by conj(r)
        const float vp = __shfl_sync(ODPD_FULL, vv, partner);
        h = lowI ? fmaf(rr, vv, ri * vp) : (upQ ? fmaf(rr, vv, -ri * vp) : vv);
        float *row = st + (size_t)t * R;
        if (act) { row[lane] = hn; row[S + 16 + lane] = v2; row[2 * S + 16 + lane] = sg; }
        if (lane < 16) row[S + lane] = v1;
        if (lane < 2 * H) row[3 * S + 16 + lane] = h;
        __syncwarp();
        }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { cx[i] = nx[i]; cr[i] = nr[i]; ci[i] = ni[i]; }
    }
}

// ================================================================ forward: output + squared error, one thread per timestep
__global__ void __launch_bounds__(AP_TT) apn_head_fwd_kernel(GruArgs a, ApBufs u, int nts, int ntiles) {
    pdl_enter();
    const ApLayout L(a.H);
    const int H = a.H, T = a.T, tid = threadIdx.x, R = L.row();
    __shared__ float sWI[AP_HMAX], sWQ[AP_HMAX], sred[AP_TT / 32];
    if (tid < H) { sWI[tid] = __ldg(a.params + L.oI + tid); sWQ[tid] = __ldg(a.params + L.oQ + tid); }
    __syncthreads();
    float lsum = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        ap_tile(tile, nts, tid, b, t);
        if (t >= T) continue;
        const float *hd = u.st + ((size_t)b * T + t) * R + 3 * L.S + 16;
        float oi = 0.f, oq = 0.f;
#pragma unroll
        for (int j = 0; j < AP_HMAX; ++j)
            if (j < H) { oi = fmaf(sWI[j], hd[j], oi); oq = fmaf(sWQ[j], hd[H + j], oq); }
        const float o0 = oi - oq, o1 = oq + oi;
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

__device__ __forceinline__ float2 apn_go(const GruArgs &a, int b, int t, float gs) {
    if (a.gout) return __ldg(reinterpret_cast<const float2 *>(a.gout) + (size_t)b * a.T + t);
    const float2 o = __ldg(reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * a.T + t);
    const float2 y = iq_row(a.target, a.target_bf16, a.target_starts, b, a.T).ld(t);
    return make_float2(gs * (o.x - y.x), gs * (o.y - y.y));
}

// ================================================================ backward: the cell in reverse, one warp per sequence
// Per step it stores gv (dL/dv, S), ga1 (dL/d pre-activation of the first dense layer, 16) and the state-rotation share of dL/d(rr, ri).
__global__ void __launch_bounds__(128) apn_chain_bwd_kernel(GruArgs a, ApBufs u) {
    pdl_enter();
    const ApLayout L(a.H);
    const int H = a.H, S = L.S, U = L.U, T = a.T, lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + wi;
    __shared__ __align__(16) float sline[4][32 + 16];
    if (b >= a.B) return;
    float *line = sline[wi], *line1 = line + 32;
    const bool act = lane < S;
    const int partner = lane < H ? lane + H : (lane < 2 * H ? lane - H : lane);
    const bool lowI = lane < H, upQ = lane >= H && lane < 2 * H;
    float whc[AP_SMAX], wuc[16];        // column `lane` of W_h (lanes < 16), column 8+lane of W_u (lanes < S)
#pragma unroll
    for (int s = 0; s < AP_SMAX; ++s) whc[s] = (lane < 16 && s < S) ? __ldg(a.params + L.oWh + s * 16 + lane) : 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) wuc[i] = act ? __ldg(a.params + L.oWu + i * U + 8 + lane) : 0.f;
    const float Zs = act ? __ldg(a.params + L.oZ + lane) : 0.f, Cc = __ldg(a.params + L.oC);
    const float wo = lowI ? __ldg(a.params + L.oI + lane) : (upQ ? __ldg(a.params + L.oQ + lane - H) : 0.f);
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
    const float *fr = u.fr + (size_t)b * T * 12;
    const int R = L.row();
    const float *st = u.st + (size_t)b * T * R;
    float gh = 0.f;                                   // adjoint of the state after step t (same lane layout as h)
    // everything a step reads from memory is fetched one step ahead: the serial loop never waits on a global load
    struct StepIn { float rr, ri, da, dq, hn, v2, sg, v1, hprev; };
    auto fetch = [&](int t) {
        StepIn q{};
        if (t < 0) return q;
        const float *row = st + (size_t)t * R;
        q.rr = __ldg(fr + (size_t)t * 12 + 6); q.ri = __ldg(fr + (size_t)t * 12 + 7);
        const float2 go = apn_go(a, b, t, gs);
        q.da = go.x + go.y; q.dq = go.y - go.x;
        q.hn = act ? __ldg(row + lane) : 0.f; q.v2 = act ? __ldg(row + S + 16 + lane) : 0.f; q.sg = act ? __ldg(row + 2 * S + 16 + lane) : 0.f;
        q.v1 = lane < 16 ? __ldg(row + S + lane) : 0.f;
        q.hprev = (lane < 2 * H && t > 0) ? __ldg(row - R + 3 * S + 16 + lane) : 0.f;
        return q;
    };
    auto step = [&](int t, const StepIn &c) {
        const float rr = c.rr, ri = c.ri, da = c.da, dq = c.dq, v2 = c.v2, sg = c.sg, v1 = c.v1, hprev = c.hprev;
        const float vv = fmaf(Zs, v2, sg);
        // head + rotation back by conj(r):  h_I = rr aI + ri aQ,  h_Q = rr aQ - ri aI
        const float g = gh + (lowI ? da : dq) * wo;                         // lanes >= 2H: wo = 0
        const float gp_ = __shfl_sync(ODPD_FULL, g, partner), vp = __shfl_sync(ODPD_FULL, vv, partner);
        const float gv = lowI ? fmaf(g, rr, -gp_ * ri) : (upQ ? fmaf(gp_, ri, g * rr) : gh);
        float prr = (lane < 2 * H) ? g * vv : 0.f;                          // dL/drr: gI aI + gQ aQ
        float pri = lowI ? g * vp : (upQ ? -g * vp : 0.f);                  // dL/dri: gI aQ - gQ aI
        // v = sigmoid(C hnew) + Z v2
        const float ds = sg * (1.f - sg);
        float ghn = gv * ds * Cc;
        const float ga2 = gv * Zs * (1.f - v2 * v2);
        line[lane] = act ? ga2 : 0.f;
        __syncwarp();
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 v = reinterpret_cast<const float4 *>(line)[q];
            if (4 * q < AP_SMAX) a0 = fmaf(whc[4 * q], v.x, a0);
            if (4 * q + 1 < AP_SMAX) a1 = fmaf(whc[4 * q + 1], v.y, a1);
            if (4 * q + 2 < AP_SMAX) a0 = fmaf(whc[4 * q + 2], v.z, a0);
            if (4 * q + 3 < AP_SMAX) a1 = fmaf(whc[4 * q + 3], v.w, a1);
        }
        const float ga1 = (a0 + a1) * (1.f - v1 * v1);                      // lanes < 16
        if (lane < 16) line1[lane] = ga1;
        __syncwarp();
        float c0 = 0.f, c1 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = reinterpret_cast<const float4 *>(line1)[q];
            c0 = fmaf(wuc[4 * q], v.x, c0); c1 = fmaf(wuc[4 * q + 1], v.y, c1); c0 = fmaf(wuc[4 * q + 2], v.z, c0); c1 = fmaf(wuc[4 * q + 3], v.w, c1);
        }
        ghn += c0 + c1;                                                    // + dL/du[8+s]
        // rotation of the previous state by r:  h_I' = pI rr - pQ ri,  h_Q' = pI ri + pQ rr
        const float gnp = __shfl_sync(ODPD_FULL, ghn, partner), hpp = __shfl_sync(ODPD_FULL, hprev, partner);
        gh = lowI ? fmaf(ghn, rr, gnp * ri) : (upQ ? fmaf(-gnp, ri, ghn * rr) : (act ? ghn : 0.f));
        prr += (lane < 2 * H) ? ghn * hprev : 0.f;                          // gI' pI + gQ' pQ
        pri += lowI ? -ghn * hpp : (upQ ? ghn * hpp : 0.f);                 // -gI' pQ + gQ' pI
        prr = warp_sum(prr); pri = warp_sum(pri);
        if (act) u.gv[((size_t)b * T + t) * S + lane] = gv;
        if (lane < 16) u.ga1[((size_t)b * T + t) * 16 + lane] = ga1;
        if (lane == 0) { u.grr[((size_t)b * T + t) * 2] = prr; u.grr[((size_t)b * T + t) * 2 + 1] = pri; }
        __syncwarp();
    };
    StepIn A = fetch(T - 1), Bq = fetch(T - 2);
    for (int t = T - 1; t >= 0; t -= 2) {
        const StepIn An = fetch(t - 2);
        step(t, A);
        const StepIn Bn = fetch(t - 3);
        if (t - 1 >= 0) step(t - 1, Bq);
        A = An; Bq = Bn;
    }
}

// ================================================================ backward: front.  Input rotation and r -> dL/dx (direct part), dL/dFIR; weight gradients
// per-tile shared factors (odd pitches):  ga1[t][16] | uu[t][U] (rotated inputs + hnew) | ga2[t][S] | v1[t][16] | gz[t][S] (gv v2) | gc[t] | hd[t][2H] |
//                                         dd[t][2] (da, dq) | dfir[t][6] | x[t + 15 halo][2]
template <bool DW>
__global__ void __launch_bounds__(AP_TT) apn_front_bwd_kernel(GruArgs a, ApBufs u, int nts, int ntiles) {
    pdl_enter();
    const ApLayout L(a.H);
    const int H = a.H, S = L.S, U = L.U, T = a.T, tid = threadIdx.x, R = L.row();
    const int UP = U | 1, SP = S | 1, HP2 = (2 * H) | 1;
    extern __shared__ __align__(16) float asm_[];
    float *sF = asm_;                       // [96] FIR weights
    float *sWx = sF + 96;                   // [16][8] W_u columns of the rotated inputs
    float *sZ = sWx + 128;                  // [32]
    float *sA1 = sZ + 32;                   // [64][17]
    float *sU = sA1 + AP_TT * 17;           // [64][UP]
    float *sA2 = sU + AP_TT * UP;           // [64][SP]
    float *sV1 = sA2 + AP_TT * SP;          // [64][17]
    float *sGz = sV1 + AP_TT * 17;          // [64][SP]
    float *sGc = sGz + AP_TT * SP;          // [64]
    float *sHd = sGc + AP_TT;               // [64][HP2]
    float *sDd = sHd + AP_TT * HP2;         // [64][3]
    float *sDf = sDd + AP_TT * 3;           // [64][7]
    float *sX = sDf + AP_TT * 7;            // [79][2]
    for (int i = tid; i < 96; i += AP_TT) sF[i] = __ldg(a.params + i);
    for (int i = tid; i < 128; i += AP_TT) sWx[i] = __ldg(a.params + L.oWu + (i >> 3) * U + (i & 7));
    if (tid < 32) sZ[tid] = tid < S ? __ldg(a.params + L.oZ + tid) : 0.f;
    __syncthreads();
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
    float *prt = (DW && u.partials) ? u.partials + (size_t)blockIdx.x * L.P : nullptr;
    bool first = true;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        ap_tile(tile, nts, tid, b, t);
        const bool valid = t < T;
        const int tbase = t - tid;
        const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
        float dfir[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, xin[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ga1[16];
        float2 go = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) ga1[i] = 0.f;
        if (valid) {
            const float *fr = u.fr + ((size_t)b * T + t) * 12;
            const float rr = fr[6], ri = fr[7], mag = fr[8];
            const float2 c = x2.ld(t);
            go = apn_go(a, b, t, gs);
#pragma unroll
            for (int i = 0; i < 16; ++i) ga1[i] = u.ga1[((size_t)b * T + t) * 16 + i];
            float grr = u.grr[((size_t)b * T + t) * 2], gri = u.grr[((size_t)b * T + t) * 2 + 1];
            float gxi = 0.f, gxq = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float gre = 0.f, gim = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) { gre = fmaf(ga1[i], sWx[i * 8 + 2 * k], gre); gim = fmaf(ga1[i], sWx[i * 8 + 2 * k + 1], gim); }
                const float ak = k < 3 ? fr[k] : c.x, bk = k < 3 ? fr[3 + k] : c.y;
                xin[2 * k] = fmaf(rr, ak, -ri * bk); xin[2 * k + 1] = fmaf(ri, ak, rr * bk);
                const float gak = fmaf(gre, rr, gim * ri), gbk = fmaf(-gre, ri, gim * rr);
                grr += fmaf(gre, ak, gim * bk); gri += fmaf(-gre, bk, gim * ak);
                if (k < 3) { dfir[k] = gak; dfir[3 + k] = gbk; } else { gxi = gak; gxq = gbk; }
            }
            // rr = I/|x|, ri = -Q/|x|
            const float im = 1.f / mag, im3 = im * im * im;
            gxi += grr * (im - c.x * c.x * im3) + gri * (c.y * c.x * im3);
            gxq += grr * (-c.x * c.y * im3) + gri * (-im + c.y * c.y * im3);
            float *df = u.dfir + ((size_t)b * T + t) * 6;
#pragma unroll
            for (int p = 0; p < 6; ++p) df[p] = dfir[p];
            reinterpret_cast<float2 *>(u.gxd)[(size_t)b * T + t] = make_float2(gxi, gxq);
        }
        if constexpr (DW) {
            const float *row = u.st + ((size_t)b * T + t) * R;
            const float *gvp = u.gv + ((size_t)b * T + t) * S;
#pragma unroll
            for (int i = 0; i < 16; ++i) { sA1[tid * 17 + i] = ga1[i]; sV1[tid * 17 + i] = valid ? row[S + i] : 0.f; }
#pragma unroll
            for (int k = 0; k < 8; ++k) sU[tid * UP + k] = xin[k];
            float gc = 0.f;
            for (int s = 0; s < S; ++s) {
                float hn = 0.f, v2 = 0.f, sg = 0.f, gv = 0.f;
                if (valid) { hn = row[s]; v2 = row[S + 16 + s]; sg = row[2 * S + 16 + s]; gv = gvp[s]; }
                sU[tid * UP + 8 + s] = hn;
                sA2[tid * SP + s] = gv * sZ[s] * (1.f - v2 * v2);
                sGz[tid * SP + s] = gv * v2;
                gc = fmaf(gv * sg * (1.f - sg), hn, gc);
            }
            sGc[tid] = gc;
            for (int j = 0; j < 2 * H; ++j) sHd[tid * HP2 + j] = valid ? row[3 * S + 16 + j] : 0.f;
            sDd[tid * 3] = go.x + go.y; sDd[tid * 3 + 1] = go.y - go.x;
#pragma unroll
            for (int p = 0; p < 6; ++p) sDf[tid * 7 + p] = dfir[p];
            for (int i = tid; i < AP_TT + AP_M - 1; i += AP_TT) {
                const int s = tbase - (AP_M - 1) + i;
                float2 v = make_float2(0.f, 0.f);
                if (s >= 0 && s < T) v = x2.ld(s);
                sX[2 * i] = v.x; sX[2 * i + 1] = v.y;
            }
            __syncthreads();
            if (prt) {
                for (int o = tid; o < L.P; o += AP_TT) {
                    float s = 0.f;
                    if (o < 96) {                                    // fir_I / fir_Q [p][m]
                        const bool isq = o >= 48;
                        const int p = (o - (isq ? 48 : 0)) / AP_M, m = (o - (isq ? 48 : 0)) - p * AP_M;
                        for (int tt = 0; tt < AP_TT; ++tt) {
                            const float di = sDf[tt * 7 + p], dq = sDf[tt * 7 + 3 + p], xi = sX[2 * (tt + m)], xq = sX[2 * (tt + m) + 1];
                            s += isq ? fmaf(dq, xi, -di * xq) : fmaf(di, xi, dq * xq);
                        }
                    } else if (o == L.oC) {
                        for (int tt = 0; tt < AP_TT; ++tt) s += sGc[tt];
                    } else if (o < L.oWu) {                          // Z[s]
                        const int q = o - L.oZ;
                        for (int tt = 0; tt < AP_TT; ++tt) s += sGz[tt * SP + q];
                    } else if (o < L.obu) {                          // W_u[i][k]
                        const int i = (o - L.oWu) / U, k = (o - L.oWu) - i * U;
                        for (int tt = 0; tt < AP_TT; ++tt) s = fmaf(sA1[tt * 17 + i], sU[tt * UP + k], s);
                    } else if (o < L.oWh) {                          // b_u[i]
                        const int i = o - L.obu;
                        for (int tt = 0; tt < AP_TT; ++tt) s += sA1[tt * 17 + i];
                    } else if (o < L.obh) {                          // W_h[s][i]
                        const int q = (o - L.oWh) >> 4, i = (o - L.oWh) & 15;
                        for (int tt = 0; tt < AP_TT; ++tt) s = fmaf(sA2[tt * SP + q], sV1[tt * 17 + i], s);
                    } else if (o < L.oI) {                           // b_h[s]
                        const int q = o - L.obh;
                        for (int tt = 0; tt < AP_TT; ++tt) s += sA2[tt * SP + q];
                    } else if (o < L.oQ) {                           // output_layer_I[j]
                        const int j = o - L.oI;
                        for (int tt = 0; tt < AP_TT; ++tt) s = fmaf(sDd[tt * 3], sHd[tt * HP2 + j], s);
                    } else {                                         // output_layer_Q[j]
                        const int j = o - L.oQ;
                        for (int tt = 0; tt < AP_TT; ++tt) s = fmaf(sDd[tt * 3 + 1], sHd[tt * HP2 + H + j], s);
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
            for (int o = tid; o < L.P; o += AP_TT) prt[o] = 0.f;
    }
}

// ================================================================ backward: transposed FIR (+ the direct part)
__global__ void __launch_bounds__(AP_TT) apn_dx_kernel(GruArgs a, ApBufs u, int nts, int ntiles) {
    pdl_enter();
    const int T = a.T, tid = threadIdx.x;
    __shared__ float sF[96];
    __shared__ float sD[(AP_TT + AP_M - 1) * 7];
    for (int i = tid; i < 96; i += AP_TT) sF[i] = __ldg(a.params + i);
    __syncthreads();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, s;
        ap_tile(tile, nts, tid, b, s);
        const int s0 = s - tid;
        for (int i = tid; i < (AP_TT + AP_M - 1) * 6; i += AP_TT) {
            const int r = i / 6, c = i - r * 6, t = s0 + r;
            sD[r * 7 + c] = t < T ? __ldg(u.dfir + ((size_t)b * T + t) * 6 + c) : 0.f;
        }
        __syncthreads();
        if (s < T) {
            const float2 d0 = reinterpret_cast<const float2 *>(u.gxd)[(size_t)b * T + s];
            float gi = d0.x, gq = d0.y;
#pragma unroll
            for (int m = 0; m < AP_M; ++m) {
                const float *df = sD + (tid + (AP_M - 1) - m) * 7;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const float di = df[p], dq = df[3 + p], wi = sF[p * AP_M + m], wq = sF[48 + p * AP_M + m];
                    gi = fmaf(di, wi, fmaf(dq, wq, gi));
                    gq = fmaf(dq, wi, fmaf(-di, wq, gq));
                }
            }
            reinterpret_cast<float2 *>(a.gx)[(size_t)b * T + s] = make_float2(gi, gq);
        }
        __syncthreads();
    }
}

// ================================================================ host
static int ap_grid(int B, int T) {
    const int64_t tiles = (int64_t)B * ((T + AP_TT - 1) / AP_TT);
    const int64_t cap = 8 * (int64_t)num_sms();
    return (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
}
int64_t apnrru_nparams(int H) { return ApLayout(H).P; }
int64_t apnrru_saved_floats(int B, int T, int H) {
    const int64_t bt = (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1);
    return bt * (12 + 16 + ApLayout(H).row()) + 4;
}
int64_t apnrru_workspace_floats(int B, int T, int H) {
    const ApLayout L(H);
    const int64_t bt = (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1);
    return (((int64_t)ap_grid(B, T) * L.P + 3) & ~(int64_t)3) + bt * (L.S + 16 + 2 + 6 + 2) + 4;
}

int apnrru_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out) {
    if (a.H < 1 || a.H > AP_HMAX) { set_error("APNRRU: hidden_size %d outside 1..%d (one warp lane per state value: 2H+3 <= 31)", a.H, AP_HMAX); return -1; }
    if (a.T < AP_M - 1) { set_error("APNRRU needs frame_length >= 15 (the reference pads with zeros_like(x[:, -15:]), apnrru.py:69-71; got %d)", a.T); return -1; }
    const ApLayout L(a.H);
    const int nts = (a.T + AP_TT - 1) / AP_TT, ntiles = a.B * nts, grid = ap_grid(a.B, a.T), wpc = a.B <= 2 * num_sms() ? 1 : 4 /* few sequences: one chain warp per CTA spreads them over the SMs */, cgrid = (a.B + wpc - 1) / wpc;
    const int64_t bt = (int64_t)a.B * a.T;
    if (!a.saved) { set_error("APNRRU needs the `saved` buffer (odpd_saved_bytes), also without ODPD_F_SAVE"); return -1; }
    ApBufs u{};
    u.fr = a.saved; u.xp = u.fr + bt * 12; u.st = u.xp + bt * 16;
    if (dir == 0) {
        launch_pdl(apn_front_kernel, dim3(grid), dim3(AP_TT), 0, st, a, u, nts, ntiles);
        launch_pdl(apn_chain_fwd_kernel, dim3(cgrid), dim3(32 * wpc), 0, st, a, u);
        launch_pdl(apn_head_fwd_kernel, dim3(grid), dim3(AP_TT), 0, st, a, u, nts, ntiles);
        return check_launch("apnrru forward");
    }
    if (!a.partials) { set_error("APNRRU backward needs the workspace (odpd_bwd_workspace_bytes)"); return -1; }
    if (a.need_dx && !a.gx) { set_error("APNRRU backward: ODPD_F_NEED_DX without gx"); return -1; }
    const int64_t poff = ((int64_t)grid * L.P + 3) & ~(int64_t)3;
    // gxd is read and written as float2: it sits first, at a 4-float-aligned offset (behind the odd-sized sections it would be misaligned whenever B*T is odd)
    u.partials = a.partials; u.gxd = a.partials + poff; u.gv = u.gxd + bt * 2; u.ga1 = u.gv + bt * L.S; u.grr = u.ga1 + bt * 16; u.dfir = u.grr + bt * 2;
    launch_pdl(apn_chain_bwd_kernel, dim3(cgrid), dim3(32 * wpc), 0, st, a, u);
    const size_t bsm = (size_t)(96 + 128 + 32 + AP_TT * (17 + (L.U | 1) + (L.S | 1) + 17 + (L.S | 1) + 1 + ((2 * a.H) | 1) + 3 + 7) + 2 * (AP_TT + AP_M - 1)) * sizeof(float);
    if (dw) {
        static std::mutex mu;
        { std::lock_guard<std::mutex> lock(mu); cudaFuncSetAttribute(apn_front_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); }
        launch_pdl(apn_front_bwd_kernel<true>, dim3(grid), dim3(AP_TT), bsm, st, a, u, nts, ntiles);
    } else {
        launch_pdl(apn_front_bwd_kernel<false>, dim3(grid), dim3(AP_TT), bsm, st, a, u, nts, ntiles);
    }
    if (a.need_dx) launch_pdl(apn_dx_kernel, dim3(grid), dim3(AP_TT), 0, st, a, u, nts, ntiles);
    if (rows_out) *rows_out = grid;
    return check_launch("apnrru backward");
}

}  // namespace odpd

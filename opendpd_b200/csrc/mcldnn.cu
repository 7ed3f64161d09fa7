// mcldnn.cu — MCLDNN backbone: convolutional front end + LSTM(8) + two linear layers, forward / backward.
//
// Replaces (reference, file:line): backbones/mcldnn.py:9-113 —
//   features (I,Q,|x|,|x|^2,|x|^3) (:88-93); the window of timestep t holds samples (t-4+m) mod T, m = 0..4 (`pad = x[:, -(memory_length-1):, :]`,
//   :96-99) as a 5x5 image (feature, memory); conv2d_1 (1 -> C, 3x3, pad 1) (:101); conv1d over memory with the features as 5 groups (5 -> 5C,
//   k=3, pad 1), re-viewed as (C,5,5): channel oc lands at [oc / 5][oc % 5] (:102-103); both concatenated along the height (:104), transposed,
//   conv2d_2 (10 -> 1, 3x3, pad 1) over (C,5) (:106) -> 5C values per timestep; LSTM(5C -> 8) (:108); fc_out(8 -> 16), fc_out_2(16 -> 2) with no
//   activation in between (:109-110).  C = the CLI's hidden size (models.py:136-138), 1..12 here.
//
// The front end has NO nonlinearity: the 5C LSTM inputs are a fixed linear map K (5C x 25) + k0 of the 25 window features, so the LSTM's input
// projection is one 32 x 25 matrix  Mx = W_ih K,  bx = b_ih + b_hh + W_ih k0  applied to the window.  A one-CTA kernel composes K, k0, Mx, bx from
// the convolution weights per call (parameter space: ~10^5 MACs); everything that scales with time only sees Mx.  Backward likewise: the
// time-parallel kernels produce dL/dMx, dL/dbx (+ dW_hh and the two linear layers) as tile outer products, and a one-CTA kernel pulls them back
// through the composition to the convolution / W_ih gradients and writes ONE finished gradient row for reduce_partials_kernel.
//   forward   compose -> xp (one thread per timestep: window features, XP = bx + Mx w) -> chain (one warp per sequence: lane = (gate, unit),
//             W_hh row in registers, h and the gate values exchanged by shuffles) -> head (one thread per timestep)
//   backward  head_bwd (dL/dh from the two linear layers) -> chain_bwd (reverse, stores the 32 gate gradients per step) -> post (dL/dw = Mx^T ga ->
//             feature Jacobian -> one (dI,dQ) contribution per (timestep, tap); tile outer products into intermediate partial rows) -> gather
//             (dL/dx, taps summed in order) -> params (ordered reduction of the intermediate rows, pull-back through the composition)
//
// Flat parameter layout (named_parameters() order): conv2d_1.weight(C,1,3,3) .bias(C) conv1d.weight(5C,1,3) .bias(5C) conv2d_2.weight(1,10,3,3) .bias(1)
// lstm.weight_ih_l0(32,5C) weight_hh_l0(32,8) bias_ih_l0(32) bias_hh_l0(32) fc_out.weight(16,8) .bias(16) fc_out_2.weight(2,16) .bias(2) = 190 C + 589.
#include <mutex>
#include "cells.h"
#include "chunking.cuh"

namespace odpd {

static constexpr int MC_TT = 64;
static constexpr int MC_CMAX = 12;
static constexpr int MC_PS = 800 + 32 + 256 + 128 + 16 + 32 + 2;   // intermediate gradient row: dMx | dbx | dW_hh | dF1 | dF1b | dF2 | dF2b  (1266)

struct McLayout {
    int C, IN, oW1, ob1, oWc, obc, oW2, ob2, oWih, oWhh, obih, obhh, oF1, oF1b, oF2, oF2b, P;
    __host__ __device__ explicit McLayout(int c) {
        C = c; IN = 5 * c;
        oW1 = 0; ob1 = 9 * c; oWc = ob1 + c; obc = oWc + 15 * c; oW2 = obc + 5 * c; ob2 = oW2 + 90; oWih = ob2 + 1; oWhh = oWih + 32 * IN;
        obih = oWhh + 256; obhh = obih + 32; oF1 = obhh + 32; oF1b = oF1 + 128; oF2 = oF1b + 16; oF2b = oF2 + 32; P = oF2b + 2;
    }
};
// composed maps (head of `saved`):  K [IN][25] | k0 [IN] | Mx [32][25] | bx [32]   (padded to 4 floats)
__host__ __device__ inline int mc_comp_floats(int C) { return ((5 * C * 26 + 832) + 3) & ~3; }
// saved:  comp | XP [B][T][32] | ACT [B][T][48] = gates i f g o (32) | c (8) | h (8)
// workspace: final row [P] (4-aligned) | intermediate rows [R][MC_PS] | DH [B][T][8] | GA [B][T][32] | contributions [B][T][5] float2
struct McBufs { float *comp, *xp, *act, *row0, *inter, *dh, *ga; float2 *contrib; int rows; };

__device__ __forceinline__ void mc_tile(int tile, int nts, int tid, int &b, int &t) {
    b = tile / nts;
    t = (tile - b * nts) * MC_TT + tid;
}
// impulse response of the two first-stage convolutions: value at Z[c2][r][m2] for a unit input at window position (f, m)
__device__ __forceinline__ float mc_zimp(const float *sp, const McLayout &L, int f, int m, int c2, int r, int m2) {
    const int dm = m - m2 + 1;
    if (dm < 0 || dm > 2) return 0.f;
    if (r < 5) {
        const int df = f - r + 1;
        return (df < 0 || df > 2) ? 0.f : sp[L.oW1 + c2 * 9 + df * 3 + dm];
    }
    const int oc = c2 * 5 + r - 5;
    return (oc / L.C == f) ? sp[L.oWc + oc * 3 + dm] : 0.f;
}
__device__ __forceinline__ float mc_z0(const float *sp, const McLayout &L, int c2, int r) { return r < 5 ? sp[L.ob1 + c2] : sp[L.obc + c2 * 5 + r - 5]; }

// ================================================================ forward: compose K, k0, Mx, bx   (one CTA)
__global__ void __launch_bounds__(256) mcl_compose_kernel(GruArgs a, McBufs u) {
    pdl_enter();
    const McLayout L(a.H);
    const int C = L.C, IN = L.IN, tid = threadIdx.x;
    extern __shared__ __align__(16) float msm[];
    float *sp = msm;                       // parameters up to the LSTM's W_hh (convolutions + W_ih)
    float *sK = sp + ((L.oWhh + 3) & ~3);  // [IN][26]  K | k0
    for (int i = tid; i < L.oWhh; i += 256) sp[i] = __ldg(a.params + i);
    __syncthreads();
    for (int e = tid; e < IN * 26; e += 256) {
        const int o = e / 26, fm = e - o * 26, c = o / 5, mo = o - c * 5;
        float acc = fm == 25 ? sp[L.ob2] : 0.f;
        for (int r = 0; r < 10; ++r)
            for (int dc = 0; dc < 3; ++dc) {
                const int c2 = c + dc - 1;
                if (c2 < 0 || c2 >= C) continue;
                for (int dm = 0; dm < 3; ++dm) {
                    const int m2 = mo + dm - 1;
                    if (m2 < 0 || m2 > 4) continue;
                    const float z = fm == 25 ? mc_z0(sp, L, c2, r) : mc_zimp(sp, L, fm / 5, fm % 5, c2, r, m2);
                    acc = fmaf(sp[L.oW2 + r * 9 + dc * 3 + dm], z, acc);
                }
            }
        sK[e] = acc;
        if (fm < 25) u.comp[o * 25 + fm] = acc; else u.comp[IN * 25 + o] = acc;
    }
    __syncthreads();
    float *Mx = u.comp + IN * 26, *bx = Mx + 800;
    for (int e = tid; e < 32 * 26; e += 256) {
        const int g = e / 26, fm = e - g * 26;
        float acc = fm == 25 ? __ldg(a.params + L.obih + g) + __ldg(a.params + L.obhh + g) : 0.f;
        for (int o = 0; o < IN; ++o) acc = fmaf(sp[L.oWih + g * IN + o], sK[o * 26 + fm], acc);
        if (fm < 25) Mx[g * 25 + fm] = acc; else bx[g] = acc;
    }
}

// window features of timestep t: w[f*5 + m], f = (I,Q,|x|,|x|^2,|x|^3), m = tap (sample (t-4+m) mod T)
__device__ __forceinline__ void mc_window(const IqRow &x2, int t, int T, float (&w)[25], float (&iq)[5][2]) {
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        int s = (t - 4 + m) % T;
        if (s < 0) s += T;
        const float2 v = x2.ld(s);
        const float a2 = __fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y));
        const float am = __fsqrt_rn(a2);
        w[m] = v.x; w[5 + m] = v.y; w[10 + m] = am; w[15 + m] = a2; w[20 + m] = __fmul_rn(__fmul_rn(am, am), am);
        iq[m][0] = v.x; iq[m][1] = v.y;
    }
}

// ================================================================ forward: XP = bx + Mx w, one thread per timestep
__global__ void __launch_bounds__(MC_TT) mcl_xp_kernel(GruArgs a, McBufs u, int nts, int ntiles) {
    pdl_enter();
    const int T = a.T, tid = threadIdx.x, IN = 5 * a.H;
    __shared__ __align__(16) float sM[832];
    for (int i = tid; i < 832; i += MC_TT) sM[i] = u.comp[IN * 26 + i];
    __syncthreads();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        mc_tile(tile, nts, tid, b, t);
        if (t >= T) continue;
        float w[25], iq[5][2];
        mc_window(iq_row(a.x, a.x_bf16, a.x_starts, b, T), t, T, w, iq);
        float *xp = u.xp + ((size_t)b * T + t) * 32;
#pragma unroll 4
        for (int g = 0; g < 32; ++g) {
            float acc = sM[800 + g];
#pragma unroll
            for (int k = 0; k < 25; ++k) acc = fmaf(sM[g * 25 + k], w[k], acc);
            xp[g] = acc;
        }
    }
}

// ================================================================ forward: LSTM(8), one warp per sequence, lane = (gate g = lane >> 3, unit j = lane & 7)
__global__ void __launch_bounds__(128) mcl_chain_fwd_kernel(GruArgs a, McBufs u) {
    pdl_enter();
    const McLayout L(a.H);
    const int T = a.T, lane = threadIdx.x & 31, g = lane >> 3, j = lane & 7;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= a.B) return;
    float wh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) wh[k] = __ldg(a.params + L.oWhh + lane * 8 + k);
    const float *xp = u.xp + (size_t)b * T * 32 + lane;
    float *act = u.act + (size_t)b * T * 48;
    float h = 0.f, c = 0.f;                 // every lane carries the state of unit j
    float cx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) cx[i] = i < T ? __ldg(xp + (size_t)i * 32) : 0.f;
    for (int t0 = 0; t0 < T; t0 += 4) {
        float nx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) nx[i] = t0 + 4 + i < T ? __ldg(xp + (size_t)(t0 + 4 + i) * 32) : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = t0 + i;
            if (t < T) {
                float p0 = cx[i], p1 = 0.f;
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    p0 = fmaf(wh[k], __shfl_sync(ODPD_FULL, h, k), p0);
                    p1 = fmaf(wh[k + 1], __shfl_sync(ODPD_FULL, h, k + 1), p1);
                }
                const float pre = p0 + p1;
                const float gv = g == 2 ? tanhf_(pre) : sigmoidf_(pre);
                const float ig = __shfl_sync(ODPD_FULL, gv, j), fg = __shfl_sync(ODPD_FULL, gv, 8 + j), gg = __shfl_sync(ODPD_FULL, gv, 16 + j),
                            og = __shfl_sync(ODPD_FULL, gv, 24 + j);
                c = fmaf(fg, c, ig * gg);
                h = og * tanhf_(c);
                float *row = act + (size_t)t * 48;
                row[lane] = gv;
                if (lane < 8) { row[32 + lane] = c; row[40 + lane] = h; }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) cx[i] = nx[i];
    }
}

// ================================================================ forward: out = fc_out_2(fc_out(h)), squared error, one thread per timestep
__global__ void __launch_bounds__(MC_TT) mcl_head_fwd_kernel(GruArgs a, McBufs u, int nts, int ntiles) {
    pdl_enter();
    const McLayout L(a.H);
    const int T = a.T, tid = threadIdx.x;
    __shared__ float sH[178], sred[MC_TT / 32];
    for (int i = tid; i < 178; i += MC_TT) sH[i] = __ldg(a.params + L.oF1 + i);
    __syncthreads();
    float lsum = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        mc_tile(tile, nts, tid, b, t);
        if (t >= T) continue;
        const float *hp = u.act + ((size_t)b * T + t) * 48 + 40;
        float hv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) hv[k] = hp[k];
        float o0 = sH[176], o1 = sH[177];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float y = sH[128 + k];
#pragma unroll
            for (int q = 0; q < 8; ++q) y = fmaf(sH[k * 8 + q], hv[q], y);
            o0 = fmaf(sH[144 + k], y, o0);
            o1 = fmaf(sH[160 + k], y, o1);
        }
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

__device__ __forceinline__ float2 mcl_go(const GruArgs &a, int b, int t, float gs) {
    if (a.gout) return __ldg(reinterpret_cast<const float2 *>(a.gout) + (size_t)b * a.T + t);
    const float2 o = __ldg(reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * a.T + t);
    const float2 y = iq_row(a.target, a.target_bf16, a.target_starts, b, a.T).ld(t);
    return make_float2(gs * (o.x - y.x), gs * (o.y - y.y));
}

// ================================================================ backward: dL/dh of the two linear layers, one thread per timestep
__global__ void __launch_bounds__(MC_TT) mcl_head_bwd_kernel(GruArgs a, McBufs u, int nts, int ntiles) {
    pdl_enter();
    const McLayout L(a.H);
    const int T = a.T, tid = threadIdx.x;
    __shared__ float sH[178];
    for (int i = tid; i < 178; i += MC_TT) sH[i] = __ldg(a.params + L.oF1 + i);
    __syncthreads();
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        mc_tile(tile, nts, tid, b, t);
        if (t >= T) continue;
        const float2 go = mcl_go(a, b, t, gs);
        float dh[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float dy = fmaf(go.x, sH[144 + k], go.y * sH[160 + k]);
#pragma unroll
            for (int q = 0; q < 8; ++q) dh[q] = fmaf(dy, sH[k * 8 + q], dh[q]);
        }
        float *d = u.dh + ((size_t)b * T + t) * 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] = dh[q];
    }
}

// ================================================================ backward: LSTM in reverse, one warp per sequence.  GA[b][t][32] = gate gradients
__global__ void __launch_bounds__(128) mcl_chain_bwd_kernel(GruArgs a, McBufs u) {
    pdl_enter();
    const McLayout L(a.H);
    const int T = a.T, lane = threadIdx.x & 31, g = lane >> 3, j = lane & 7;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= a.B) return;
    float whc[32];                          // column j of W_hh
#pragma unroll
    for (int r = 0; r < 32; ++r) whc[r] = __ldg(a.params + L.oWhh + r * 8 + j);
    const float *act = u.act + (size_t)b * T * 48;
    const float *dhp = u.dh + (size_t)b * T * 8 + j;
    float *gap = u.ga + (size_t)b * T * 32 + lane;
    float rec = 0.f, dcr = 0.f;             // every lane carries the adjoints of unit j
    struct In { float dh, ig, fg, gg, og, cc, cp; };
    auto fetch = [&](int t) {
        In q{};
        if (t < 0) return q;
        const float *row = act + (size_t)t * 48;
        q.dh = __ldg(dhp + (size_t)t * 8);
        q.ig = __ldg(row + j); q.fg = __ldg(row + 8 + j); q.gg = __ldg(row + 16 + j); q.og = __ldg(row + 24 + j); q.cc = __ldg(row + 32 + j);
        q.cp = t > 0 ? __ldg(row - 48 + 32 + j) : 0.f;
        return q;
    };
    auto step = [&](int t, const In &q) {
        const float dh = q.dh + rec, tc = tanhf_(q.cc);
        const float dc = fmaf(dh * q.og, 1.f - tc * tc, dcr);
        const float gi = dc * q.gg * q.ig * (1.f - q.ig), gf = dc * q.cp * q.fg * (1.f - q.fg), gg2 = dc * q.ig * (1.f - q.gg * q.gg),
                    go = dh * tc * q.og * (1.f - q.og);
        dcr = dc * q.fg;
        const float mine = g == 0 ? gi : (g == 1 ? gf : (g == 2 ? gg2 : go));      // gradient of gate (g, j)
        gap[(size_t)t * 32] = mine;
        float r0 = 0.f, r1 = 0.f;
#pragma unroll
        for (int r = 0; r < 32; r += 2) {
            r0 = fmaf(whc[r], __shfl_sync(ODPD_FULL, mine, r), r0);
            r1 = fmaf(whc[r + 1], __shfl_sync(ODPD_FULL, mine, r + 1), r1);
        }
        rec = r0 + r1;
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

// ================================================================ backward: post.  dL/dw = Mx^T ga -> (dI,dQ) per tap; tile outer products -> intermediate rows
// per-tile shared factors (odd pitches): ga[t][33] | w[t][25] | hprev[t][9] | h[t][9] | go[t][3] | y1[t][17] | dy1[t][17]
template <bool DW>
__global__ void __launch_bounds__(MC_TT) mcl_post_kernel(GruArgs a, McBufs u, int nts, int ntiles) {
    pdl_enter();
    const McLayout L(a.H);
    const int T = a.T, tid = threadIdx.x, IN = L.IN;
    __shared__ __align__(16) float sM[800];
    __shared__ float sH[178];
    extern __shared__ __align__(16) float msm[];
    float *sGa = msm;                      // [64][33]
    float *sW = sGa + MC_TT * 33;          // [64][25]
    float *sHp = sW + MC_TT * 25;          // [64][9]
    float *sHh = sHp + MC_TT * 9;          // [64][9]
    float *sGo = sHh + MC_TT * 9;          // [64][3]
    float *sY1 = sGo + MC_TT * 3;          // [64][17]
    float *sDy = sY1 + MC_TT * 17;         // [64][17]
    for (int i = tid; i < 800; i += MC_TT) sM[i] = u.comp[IN * 26 + i];
    for (int i = tid; i < 178; i += MC_TT) sH[i] = __ldg(a.params + L.oF1 + i);
    __syncthreads();
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
    float *prt = DW ? u.inter + (size_t)blockIdx.x * MC_PS : nullptr;
    bool first = true;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, t;
        mc_tile(tile, nts, tid, b, t);
        const bool valid = t < T;
        float w[25], iq[5][2], ga[32];
#pragma unroll
        for (int k = 0; k < 25; ++k) w[k] = 0.f;
#pragma unroll
        for (int g = 0; g < 32; ++g) ga[g] = 0.f;
        if (valid) {
            mc_window(iq_row(a.x, a.x_bf16, a.x_starts, b, T), t, T, w, iq);
            const float *gp = u.ga + ((size_t)b * T + t) * 32;
#pragma unroll
            for (int g = 0; g < 32; ++g) ga[g] = gp[g];
            if (a.need_dx && u.contrib) {
                float dw[25];
#pragma unroll
                for (int k = 0; k < 25; ++k) dw[k] = 0.f;
#pragma unroll 4
                for (int g = 0; g < 32; ++g)
#pragma unroll
                    for (int k = 0; k < 25; ++k) dw[k] = fmaf(ga[g], sM[g * 25 + k], dw[k]);
#pragma unroll
                for (int m = 0; m < 5; ++m) {
                    const float am = w[10 + m], a2 = w[15 + m];
                    const float gam = fmaf(3.f * a2, dw[20 + m], dw[10 + m]);
                    const float sc = fmaf(2.f, dw[15 + m], gam / am);
                    u.contrib[((size_t)b * T + t) * 5 + m] = make_float2(fmaf(iq[m][0], sc, dw[m]), fmaf(iq[m][1], sc, dw[5 + m]));
                }
            }
        }
        if constexpr (DW) {
#pragma unroll
            for (int g = 0; g < 32; ++g) sGa[tid * 33 + g] = ga[g];
#pragma unroll
            for (int k = 0; k < 25; ++k) sW[tid * 25 + k] = w[k];
            float2 go = make_float2(0.f, 0.f);
            float hv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (valid) {
                go = mcl_go(a, b, t, gs);
                const float *row = u.act + ((size_t)b * T + t) * 48;
#pragma unroll
                for (int q = 0; q < 8; ++q) { hv[q] = row[40 + q]; sHp[tid * 9 + q] = t > 0 ? row[-48 + 40 + q] : 0.f; }
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) sHp[tid * 9 + q] = 0.f;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) sHh[tid * 9 + q] = hv[q];
            sGo[tid * 3] = go.x; sGo[tid * 3 + 1] = go.y;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float y = sH[128 + k];
#pragma unroll
                for (int q = 0; q < 8; ++q) y = fmaf(sH[k * 8 + q], hv[q], y);
                sY1[tid * 17 + k] = valid ? y : 0.f;
                sDy[tid * 17 + k] = fmaf(go.x, sH[144 + k], go.y * sH[160 + k]);
            }
            __syncthreads();
            for (int o = tid; o < MC_PS; o += MC_TT) {
                float s = 0.f;
                if (o < 800) {                       // dMx[g][fm]
                    const int g = o / 25, k = o - g * 25;
                    for (int tt = 0; tt < MC_TT; ++tt) s = fmaf(sGa[tt * 33 + g], sW[tt * 25 + k], s);
                } else if (o < 832) {                // dbx[g]
                    for (int tt = 0; tt < MC_TT; ++tt) s += sGa[tt * 33 + o - 800];
                } else if (o < 1088) {               // dW_hh[r][k]
                    const int r = (o - 832) >> 3, k = (o - 832) & 7;
                    for (int tt = 0; tt < MC_TT; ++tt) s = fmaf(sGa[tt * 33 + r], sHp[tt * 9 + k], s);
                } else if (o < 1216) {               // fc_out.weight[k][q]
                    const int k = (o - 1088) >> 3, q = (o - 1088) & 7;
                    for (int tt = 0; tt < MC_TT; ++tt) s = fmaf(sDy[tt * 17 + k], sHh[tt * 9 + q], s);
                } else if (o < 1232) {               // fc_out.bias[k]
                    for (int tt = 0; tt < MC_TT; ++tt) s += sDy[tt * 17 + o - 1216];
                } else if (o < 1264) {               // fc_out_2.weight[c][k]
                    const int c = (o - 1232) >> 4, k = (o - 1232) & 15;
                    for (int tt = 0; tt < MC_TT; ++tt) s = fmaf(sGo[tt * 3 + c], sY1[tt * 17 + k], s);
                } else {                             // fc_out_2.bias[c]
                    for (int tt = 0; tt < MC_TT; ++tt) s += sGo[tt * 3 + o - 1264];
                }
                prt[o] = first ? s : prt[o] + s;
            }
            first = false;
            __syncthreads();
        }
    }
    if constexpr (DW) {
        if (first)
            for (int o = tid; o < MC_PS; o += MC_TT) prt[o] = 0.f;
    }
}

// dL/dx[s] = sum over the five window taps that touch sample s: timestep (s+4-m) mod T, tap m   (fixed order)
__global__ void mcl_gather_kernel(const float2 *__restrict__ ctr, float2 *__restrict__ gx, int B, int T) {
    pdl_enter();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * T) return;
    const int b = (int)(i / T), s = (int)(i - (int64_t)b * T);
    float gi = 0.f, gq = 0.f;
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        const int t = (s + 4 - m) % T;
        const float2 v = ctr[((size_t)b * T + t) * 5 + m];
        gi += v.x; gq += v.y;
    }
    gx[i] = make_float2(gi, gq);
}

// ================================================================ backward: parameters (one CTA).  Ordered reduction of the intermediate rows, then the
// pull-back through  Mx = W_ih K, bx = b_ih + b_hh + W_ih k0  and  K, k0 = conv2d_2 o (conv2d_1 | conv1d)  to the finished gradient row.
__global__ void __launch_bounds__(256) mcl_params_kernel(GruArgs a, McBufs u) {
    pdl_enter();
    const McLayout L(a.H);
    const int C = L.C, IN = L.IN, tid = threadIdx.x;
    extern __shared__ __align__(16) float msm[];
    float *sp = msm;                                   // parameters up to W_hh
    float *sI = sp + ((L.oWhh + 3) & ~3);              // [MC_PS] reduced intermediate gradients
    float *sdK = sI + ((MC_PS + 3) & ~3);              // [IN][26]  dK | dk0
    float *row = u.row0;
    for (int i = tid; i < L.oWhh; i += 256) sp[i] = __ldg(a.params + i);
    for (int o = tid; o < MC_PS; o += 256) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;          // four interleaved partial sums: a fixed tree, loads in flight together
        int r = 0;
        for (; r + 3 < u.rows; r += 4) {
            s0 += u.inter[(size_t)r * MC_PS + o]; s1 += u.inter[(size_t)(r + 1) * MC_PS + o];
            s2 += u.inter[(size_t)(r + 2) * MC_PS + o]; s3 += u.inter[(size_t)(r + 3) * MC_PS + o];
        }
        for (; r < u.rows; ++r) s0 += u.inter[(size_t)r * MC_PS + o];
        sI[o] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    const float *K = u.comp, *k0 = u.comp + IN * 25;
    // LSTM + linear layers: direct
    for (int o = tid; o < 256; o += 256) row[L.oWhh + o] = sI[832 + o];
    if (tid < 32) { row[L.obih + tid] = sI[800 + tid]; row[L.obhh + tid] = sI[800 + tid]; }
    for (int o = tid; o < 178; o += 256) row[L.oF1 + o] = sI[1088 + o];
    // dW_ih[g][o] = sum_fm dMx[g][fm] K[o][fm] + dbx[g] k0[o]
    for (int e = tid; e < 32 * IN; e += 256) {
        const int g = e / IN, o = e - g * IN;
        float s = sI[800 + g] * k0[o];
        for (int k = 0; k < 25; ++k) s = fmaf(sI[g * 25 + k], K[o * 25 + k], s);
        row[L.oWih + e] = s;
    }
    // dK[o][fm] = sum_g W_ih[g][o] dMx[g][fm],  dk0[o] = sum_g W_ih[g][o] dbx[g]
    for (int e = tid; e < IN * 26; e += 256) {
        const int o = e / 26, fm = e - o * 26;
        float s = 0.f;
        for (int g = 0; g < 32; ++g) s = fmaf(sp[L.oWih + g * IN + o], fm < 25 ? sI[g * 25 + fm] : sI[800 + g], s);
        sdK[e] = s;
    }
    __syncthreads();
    // conv2d_2.weight[r][dc][dm] and bias
    for (int e = tid; e < 91; e += 256) {
        float s = 0.f;
        if (e == 90) {
            for (int o = 0; o < IN; ++o) s += sdK[o * 26 + 25];
            row[L.ob2] = s;
        } else {
            const int r = e / 9, dc = (e - r * 9) / 3, dm = e - r * 9 - dc * 3;
            for (int o = 0; o < IN; ++o) {
                const int c = o / 5, mo = o - c * 5, c2 = c + dc - 1, m2 = mo + dm - 1;
                if (c2 < 0 || c2 >= C || m2 < 0 || m2 > 4) continue;
                for (int fm = 0; fm < 25; ++fm) s = fmaf(sdK[o * 26 + fm], mc_zimp(sp, L, fm / 5, fm % 5, c2, r, m2), s);
                s = fmaf(sdK[o * 26 + 25], mc_z0(sp, L, c2, r), s);
            }
            row[L.oW2 + e] = s;
        }
    }
    // dZ[(fm)][c2][r][m2] = sum over (c, mo, dc, dm) with c+dc-1 = c2, mo+dm-1 = m2 of dK[(c,mo)][fm] W2[r][dc][dm]   (fm = 25: the bias path)
    auto dz = [&](int fm, int c2, int r, int m2) {
        float s = 0.f;
        for (int dc = 0; dc < 3; ++dc) {
            const int c = c2 - dc + 1;
            if (c < 0 || c >= C) continue;
            for (int dm = 0; dm < 3; ++dm) {
                const int mo = m2 - dm + 1;
                if (mo < 0 || mo > 4) continue;
                s = fmaf(sdK[(c * 5 + mo) * 26 + fm], sp[L.oW2 + r * 9 + dc * 3 + dm], s);
            }
        }
        return s;
    };
    // conv2d_1.weight[c2][df][dm'] : impulse at (f, m) reaches Z[c2][r][m2] with df = f - r + 1, dm' = m - m2 + 1
    for (int e = tid; e < 9 * C; e += 256) {
        const int c2 = e / 9, df = (e - c2 * 9) / 3, dmm = e - c2 * 9 - df * 3;
        float s = 0.f;
        for (int r = 0; r < 5; ++r) {
            const int f = r + df - 1;
            if (f < 0 || f > 4) continue;
            for (int m2 = 0; m2 < 5; ++m2) {
                const int m = m2 + dmm - 1;
                if (m < 0 || m > 4) continue;
                s += dz(f * 5 + m, c2, r, m2);
            }
        }
        row[L.oW1 + e] = s;
    }
    // conv2d_1.bias[c2] = sum_{r<5, m2} dZ0[c2][r][m2]
    for (int c2 = tid; c2 < C; c2 += 256) {
        float s = 0.f;
        for (int r = 0; r < 5; ++r)
            for (int m2 = 0; m2 < 5; ++m2) s += dz(25, c2, r, m2);
        row[L.ob1 + c2] = s;
    }
    // conv1d.weight[oc][dm'] (input channel f = oc / C, lands at Z[oc / 5][5 + oc % 5]) and bias
    for (int e = tid; e < 20 * C; e += 256) {
        const int oc = e / 4, q = e - oc * 4, c2 = oc / 5, r = 5 + oc % 5, f = oc / C;
        float s = 0.f;
        if (q < 3) {
            for (int m2 = 0; m2 < 5; ++m2) {
                const int m = m2 + q - 1;
                if (m < 0 || m > 4) continue;
                s += dz(f * 5 + m, c2, r, m2);
            }
            row[L.oWc + oc * 3 + q] = s;
        } else {
            for (int m2 = 0; m2 < 5; ++m2) s += dz(25, c2, r, m2);
            row[L.obc + oc] = s;
        }
    }
}

// ================================================================ host
static int mc_grid(int B, int T) {
    const int64_t tiles = (int64_t)B * ((T + MC_TT - 1) / MC_TT);
    const int64_t cap = (int64_t)num_sms();          // few rows: the one-CTA parameter kernel reduces them
    return (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
}
int64_t mcldnn_nparams(int C) { return McLayout(C).P; }
int64_t mcldnn_saved_floats(int B, int T, int C) {
    const int64_t bt = (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1);
    return mc_comp_floats(C) + bt * 80 + 4;
}
int64_t mcldnn_workspace_floats(int B, int T, int C) {
    const int64_t bt = (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1);
    return ((McLayout(C).P + 3) & ~3) + (int64_t)mc_grid(B, T) * MC_PS + bt * (8 + 32 + 10) + 8;
}

int mcldnn_run(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out) {
    if (a.H < 1 || a.H > MC_CMAX) { set_error("MCLDNN: hidden size (conv channels) %d outside 1..%d", a.H, MC_CMAX); return -1; }
    if (a.T < 4) { set_error("MCLDNN needs frame_length >= 4 (the reference's wrap-around window, mcldnn.py:96-99; got %d)", a.T); return -1; }
    if (!a.saved) { set_error("MCLDNN needs the `saved` buffer (odpd_saved_bytes), also without ODPD_F_SAVE"); return -1; }
    const McLayout L(a.H);
    const int nts = (a.T + MC_TT - 1) / MC_TT, ntiles = a.B * nts, grid = mc_grid(a.B, a.T), wpc = a.B <= 2 * num_sms() ? 1 : 4 /* few sequences: one chain warp per CTA spreads them over the SMs */, cgrid = (a.B + wpc - 1) / wpc;
    const int64_t bt = (int64_t)a.B * a.T;
    McBufs u{};
    u.comp = a.saved; u.xp = a.saved + mc_comp_floats(a.H); u.act = u.xp + bt * 32;
    const size_t csm = (size_t)(((L.oWhh + 3) & ~3) + L.IN * 26) * sizeof(float);
    if (dir == 0) {
        launch_pdl(mcl_compose_kernel, dim3(1), dim3(256), csm, st, a, u);
        launch_pdl(mcl_xp_kernel, dim3(grid), dim3(MC_TT), 0, st, a, u, nts, ntiles);
        launch_pdl(mcl_chain_fwd_kernel, dim3(cgrid), dim3(32 * wpc), 0, st, a, u);
        launch_pdl(mcl_head_fwd_kernel, dim3(grid), dim3(MC_TT), 0, st, a, u, nts, ntiles);
        return check_launch("mcldnn forward");
    }
    if (!a.partials) { set_error("MCLDNN backward needs the workspace (odpd_bwd_workspace_bytes)"); return -1; }
    if (a.need_dx && !a.gx) { set_error("MCLDNN backward: ODPD_F_NEED_DX without gx"); return -1; }
    u.row0 = a.partials; u.inter = a.partials + ((L.P + 3) & ~3); u.dh = u.inter + (int64_t)grid * MC_PS; u.ga = u.dh + bt * 8;
    u.contrib = reinterpret_cast<float2 *>(u.ga + bt * 32);       // even float offset from the (8-byte aligned) workspace base
    u.rows = grid;
    launch_pdl(mcl_head_bwd_kernel, dim3(grid), dim3(MC_TT), 0, st, a, u, nts, ntiles);
    launch_pdl(mcl_chain_bwd_kernel, dim3(cgrid), dim3(32 * wpc), 0, st, a, u);
    const size_t psm = (size_t)MC_TT * (33 + 25 + 9 + 9 + 3 + 17 + 17) * sizeof(float);
    if (dw) launch_pdl(mcl_post_kernel<true>, dim3(grid), dim3(MC_TT), psm, st, a, u, nts, ntiles);
    else if (a.need_dx) launch_pdl(mcl_post_kernel<false>, dim3(grid), dim3(MC_TT), psm, st, a, u, nts, ntiles);
    if (a.need_dx) {
        const int64_t n = bt;
        launch_pdl(mcl_gather_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, (const float2 *)u.contrib, reinterpret_cast<float2 *>(a.gx), a.B, a.T);
    }
    if (dw) {
        const size_t qsm = (size_t)(((L.oWhh + 3) & ~3) + ((MC_PS + 3) & ~3) + L.IN * 26) * sizeof(float);
        launch_pdl(mcl_params_kernel, dim3(1), dim3(256), qsm, st, a, u);
    }
    if (rows_out) *rows_out = 1;
    return check_launch("mcldnn backward");
}

}  // namespace odpd

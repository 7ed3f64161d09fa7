// delta.cu — fused forward / backward of the delta-network GRU backbones.
//
// Replaces (reference, file:line):
//   DeltaGRU      backbones/deltagru.py:60-77 (features, fc_out) + DeltaGRULayer :149-266 (biases folded into the delta
//                 memories :164-170, compute_deltas :174-185, update_states :187-192, compute_gates :194-206)
//   TRes-DeltaGRU backbones/deltagru_tcnskip.py:89-103 (TCN skip :32-49, next-sample features :91-100, out=fc_out(h)+skip)
//                 + DeltaGRULayer :232-304 (bias-free x2h/h2h)
// Backward recurrences: SURVEY.md §8a-D (validated against autograd through oracle/odpd_oracle.c).
//
// Same 3-warp chunk pipeline as gru_family.cu.  What is specific here: the input-side delta logic (|x - x_hat| >= thx,
// x_hat update, W_ih * delta_x running sums) does not depend on h, so it runs in the "pre" warp one chunk ahead — with
// the reference's separately-rounded feature arithmetic, so the delta-x keep mask is bit exact; the chain warp only
// carries the h-side delta (threshold thh, delta memories of W_hh * delta_h) and the gates.
#include "cells.h"
#include "pipeline.cuh"

namespace odpd {

template <bool TRES>
struct DeltaLayout {
    int H, oWih, oWhh, obih, obhh, oWo, obo, oc0, oc2, P;
    __host__ __device__ explicit DeltaLayout(int h) {
        H = h; oWih = 0; oWhh = 18 * h;
        if (TRES) { obih = obhh = obo = -1; oWo = oWhh + 3 * h * h; oc0 = oWo + 2 * h; oc2 = oc0 + 18; P = oc2 + 6; }
        else { obih = oWhh + 3 * h * h; obhh = obih + 3 * h; oWo = obhh + 3 * h; obo = oWo + 2 * h; oc0 = oc2 = -1; P = obo + 2; }
    }
};
// saved row per step: r | z | n | Mnh | h_t | delta_h (masked)  (6 x HP)  then  delta_x[6] (masked), mask_x bits, mask_h bits
template <int HT> struct DRow { static constexpr int HP = Pad4<HT>::value; static constexpr int value = 6 * HP + 8; };

__device__ __forceinline__ float hswish(float v) { return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f; }
__device__ __forceinline__ float hswish_grad(float v) { return v < -3.f ? 0.f : (v <= 3.f ? v / 3.f + 0.5f : 1.f); }

// TCN skip at one timestep (deltagru_tcnskip.py:32-49): Conv1d(2->3,k3,dil16,pad16) -> Hardswish -> Conv1d(3->2,k1) -> Hardswish
__device__ __forceinline__ void tcn_point(IqRow x2, int T, int t, const float *w0, const float *w2, float *c1, float *c2) {
    float2 xs[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int tt = t + (k - 1) * 16;
        xs[k] = (tt >= 0 && tt < T) ? __ldg(x2 + tt) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int co = 0; co < 3; ++co) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc = fmaf(w0[(co * 2 + 1) * 3 + k], xs[k].y, fmaf(w0[(co * 2) * 3 + k], xs[k].x, acc));
        c1[co] = acc;
    }
#pragma unroll
    for (int o = 0; o < 2; ++o) c2[o] = fmaf(w2[o * 3 + 2], hswish(c1[2]), fmaf(w2[o * 3 + 1], hswish(c1[1]), w2[o * 3] * hswish(c1[0])));
}

template <int HT, bool TRES>
struct DFwdSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = DRow<HT>::value;
    static constexpr int XP = (CH + 1) * 3 * HP /* +1 spare row for the chain's last-step prefetch */, FT = CH * 8, ACT = CH * ROW, PO = CH * 33, SK = CH * 2;
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + HP + 2 * XP + 3 * FT + 3 * SK + 2 * ACT + 2 * PO + 2 * HP + CH * 8 + ROW; }
};
template <int HT, bool TRES>
struct DBwdSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = DRow<HT>::value;
    static constexpr int ACT = (CH + 1) * ROW, PRE = CH * 12, DH = CH * HP, G = CH * 4 * HP, DF = CH * 8;
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + 3 * ACT + 3 * PRE + 2 * DH + 2 * G + DF + 8; }
};

// ================================================================ forward
template <int HT, bool TRES>
__global__ void __launch_bounds__(96, 1) delta_fwd_kernel(GruArgs a) {
    constexpr int HP = Pad4<HT>::value, ROW = DRow<HT>::value, F = 6;
    constexpr int FM = TRES ? FM_TRES6 : FM_DGRU6;
    using SM = DFwdSmem<HT, TRES>;
    const DeltaLayout<TRES> L(a.H);
    const int H = a.H, T = a.T;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *zero = sp + Ppad;                 // [HP]
    float *sxp = zero + HP;                  // [2][CH][3HP]   running W_ih*delta_x sums (+ folded biases)
    float *sft = sxp + 2 * SM::XP;           // [3][CH][8]     masked delta_x[6], mask_x bits, -
    float *ssk = sft + 3 * SM::FT;           // [3][CH][2]     TCN skip
    float *sact = ssk + 3 * SM::SK;          // [2][CH][ROW]
    float *spo = sact + 2 * SM::ACT;         // [2][CH][33]
    float *sdl = spo + 2 * SM::PO;           // [2][HP]        delta_h broadcast line (double buffered)
    float *sraw = sdl + 2 * HP;              // [CH][8]        raw features of the block (pre warp only)
    float *sdump = sraw + CH * 8;            // [ROW]          dump row for the chain's first deferred store
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.x;
    stage_params(sp, a.params, L.P, bars);
    if (threadIdx.x < HP) zero[threadIdx.x] = 0.f;
    __syncthreads();
    const bool act = lane < H;
    const int j = act ? lane : 0, lp = lane < HP ? lane : 0;
    const int nchunks = (T + CH - 1) / CH;
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
    const float thx = a.thx, thh = a.thh;

    if (warp == 1) {
        // =============================== pre: features, TCN skip, input-side delta logic
        float wir[F], wiz[F], win[F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            wir[f] = act ? sp[L.oWih + (0 * H + j) * F + f] : 0.f;
            wiz[f] = act ? sp[L.oWih + (1 * H + j) * F + f] : 0.f;
            win[f] = act ? sp[L.oWih + (2 * H + j) * F + f] : 0.f;
        }
        // delta memories start from the folded biases (deltagru.py:164-170); zero for TRes (:206-211)
        float Mr = (!TRES && act) ? sp[L.obih + j] + sp[L.obhh + j] : 0.f;
        float Mz = (!TRES && act) ? sp[L.obih + H + j] + sp[L.obhh + H + j] : 0.f;
        float Mn = (!TRES && act) ? sp[L.obih + 2 * H + j] : 0.f;
        // Three phases per 32-step block, each with the lane mapping that suits it (the first version ran everything in one loop over
        // the timesteps with lane = unit and broadcast every feature with a shuffle: ~350 cycles per step, as slow as the chain warp):
        //   (i)   lane = timestep: load the samples, features (IEEE, no contraction), TCN skip
        //   (ii)  lane = feature : the x_hat recurrence |f - x_hat| >= thx (serial over time, 6 lanes, ~4 dependent ops per step);
        //                          masked delta_x, keep mask and zero count via ballots
        //   (iii) lane = unit    : running sums M += W_ih delta_x_t (the dot product first, then one add: the reference's order,
        //                          deltagru.py:196-199 `mm(delta_x, W^T) + M`), no cross-lane traffic
        float xh = 0.f;                              // x_hat of feature `lane` (lanes 0..5)
        const bool fl = lane < F;
        long long zx = 0;
        for (int s = 0; s < nchunks + 2; ++s) {
            if (s < nchunks) {
                const int t0 = s * CH, nt = min(CH, T - t0);
                float *ft = sft + (s % 3) * SM::FT, *xp = sxp + (s & 1) * SM::XP, *sk = ssk + (s % 3) * SM::SK;
                if (lane < nt) {
                    const int t = t0 + lane;
                    const float2 v = __ldg(x2 + t);
                    float2 vn = make_float2(0.f, 0.f);
                    if (TRES) vn = __ldg(x2 + ((t + 1 < T) ? t + 1 : 0));    // torch.roll(x,-1): wraps to x[0]
                    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    features_fwd<FM>(v.x, v.y, vn.x, vn.y, f);
                    float4 *r4 = reinterpret_cast<float4 *>(sraw + lane * 8);
                    r4[0] = make_float4(f[0], f[1], f[2], f[3]);
                    r4[1] = make_float4(f[4], f[5], 0.f, 0.f);
                    if (TRES) {
                        float c1[3], c2[2];
                        tcn_point(x2, T, t, sp + L.oc0, sp + L.oc2, c1, c2);
                        *reinterpret_cast<float2 *>(sk + lane * 2) = make_float2(hswish(c2[0]), hswish(c2[1]));
                    }
                }
                __syncwarp();
                unsigned zc = 0;
                const float *rawq = sraw + (fl ? lane : 0);
                float *ftq = ft + (fl ? lane : 7);
#pragma unroll 4
                for (int tl = 0; tl < nt; ++tl) {
                    const float fv = rawq[tl * 8];
                    const float d = __fsub_rn(fv, xh);
                    const float ad = fabsf(d);
                    const float dxv = (ad < thx) ? 0.f : d;          // masked_fill(|d| < th, 0)
                    const bool keep = ad >= thx;                      // where(|d| >= th, x, x_hat)
                    xh = keep ? fv : xh;
                    const unsigned mx = __ballot_sync(ODPD_FULL, keep && fl);
                    zc += __popc(__ballot_sync(ODPD_FULL, dxv == 0.f && fl));
                    ftq[tl * 8] = fl ? dxv : 0.f;                     // lanes >= 6 all write the unused word 7
                    if (lane == 0) ft[tl * 8 + 6] = __uint_as_float(mx);
                }
                zx += zc;
                __syncwarp();
                if (lane < HP) {
#pragma unroll 4
                    for (int tl = 0; tl < nt; ++tl) {
                        const float4 *d4 = reinterpret_cast<const float4 *>(ft + tl * 8);
                        const float4 d0 = d4[0], d1 = d4[1];
                        const float dx[F] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y};
                        const float pr = fmaf(wir[2], dx[2], fmaf(wir[1], dx[1], wir[0] * dx[0])) + fmaf(wir[5], dx[5], fmaf(wir[4], dx[4], wir[3] * dx[3]));
                        const float pz = fmaf(wiz[2], dx[2], fmaf(wiz[1], dx[1], wiz[0] * dx[0])) + fmaf(wiz[5], dx[5], fmaf(wiz[4], dx[4], wiz[3] * dx[3]));
                        const float pn = fmaf(win[2], dx[2], fmaf(win[1], dx[1], win[0] * dx[0])) + fmaf(win[5], dx[5], fmaf(win[4], dx[4], win[3] * dx[3]));
                        Mr += pr; Mz += pz; Mn += pn;
                        float *o = xp + tl * 3 * HP + lane;
                        o[0] = Mr; o[HP] = Mz; o[2 * HP] = Mn;
                    }
                }
            }
            __syncthreads();
        }
        if (a.stats && lane == 0) {
            atomicAdd(reinterpret_cast<unsigned long long *>(a.stats), (unsigned long long)zx);
            atomicAdd(reinterpret_cast<unsigned long long *>(a.stats) + 1, (unsigned long long)T * F);
        }
    } else if (warp == 0) {
        // =============================== chain
        if constexpr (HP <= 16) {
            // half-warp split (see gru_family.cu): lanes 0..15 own the r / n-side rows of unit u, lanes 16..31 the z row
            const int half = lane >> 4, u = lane & 15;
            const bool au = u < H, lo = (half == 0);
            const int ju = au ? u : 0, up = u < HP ? u : 0;
            float wA[HT], wB[HT];
#pragma unroll
            for (int k = 0; k < HT; ++k) {
                const bool ok = au && k < H;
                wA[k] = ok ? sp[L.oWhh + ((half ? 1 : 0) * H + ju) * H + k] : 0.f;
                wB[k] = (ok && lo) ? sp[L.oWhh + (2 * H + ju) * H + k] : 0.f;
            }
            float h = 0.f, hh = 0.f, MhA = 0.f;
            float Mnh = (!TRES && au && lo) ? sp[L.obhh + 2 * H + ju] : 0.f;
            int cur = 0;
            for (int s = 0; s < nchunks + 2; ++s) {
                const int c = s - 1;
                if (c >= 0 && c < nchunks) {
                    const int t0 = c * CH, nt = min(CH, T - t0);
                    const float *xa_p = sxp + (c & 1) * SM::XP + (half ? HP : 0) + up;
                    const float *xn_p = sxp + (c & 1) * SM::XP + 2 * HP + up;
                    float *row = sact + (c & 1) * SM::ACT + lane;
                    const bool wr = lane < HP;
                    float xa = *xa_p, xn = *xn_p;
                    // The gate values of a step are stored one iteration late, while the next step's broadcast loads are in flight; the
                    // first iteration "stores" into a dump row so that the loop body carries no branch.  The keep mask and the zero
                    // count are NOT produced here: the post warp derives both from the stored delta_h (vote + popc on the chain warp cost
                    // ~35 cycles per step: issue is in order, so even off-path consumers of a slow result stall the chain).
                    float p_s = 0.f, p_z = 0.f, p_n = 0.f, p_m = 0.f, p_h = 0.f, p_d = 0.f;
                    float *prow = sdump + lane;
#pragma unroll 2
                    for (int tl = 0; tl < nt; ++tl) {
                        xa_p += 3 * HP; xn_p += 3 * HP;
                        const float nxa = *xa_p, nxn = *xn_p;       // the last step reads the spare row of its own buffer: unused
                        const float d = h - hh;
                        const bool keep = fabsf(d) >= thh;
                        const float dh = keep ? d : 0.f;
                        hh = keep ? h : hh;
                        float *line = sdl + cur * HP;
                        if (wr) line[lane] = dh;
                        __syncwarp();
                        const float4 *l4 = reinterpret_cast<const float4 *>(line);
                        float4 lv[HP / 4];
#pragma unroll
                        for (int q = 0; q < HP / 4; ++q) lv[q] = l4[q];
                        if (wr) { prow[0] = p_s; prow[HP] = p_z; prow[2 * HP] = p_n; prow[3 * HP] = p_m; prow[4 * HP] = p_h; prow[5 * HP] = p_d; }
                        const float base = xa + MhA;
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;   // 4 chains per dot product
#pragma unroll
                        for (int q = 0; q < HP / 4; ++q) {
                            const float e[4] = {lv[q].x, lv[q].y, lv[q].z, lv[q].w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int k = q * 4 + i;
                                if (k < HT) {
                                    if (i == 0) { a0 = fmaf(wA[k], e[i], a0); b0 = fmaf(wB[k], e[i], b0); }
                                    else if (i == 1) { a1 = fmaf(wA[k], e[i], a1); b1 = fmaf(wB[k], e[i], b1); }
                                    else if (i == 2) { a2 = fmaf(wA[k], e[i], a2); b2 = fmaf(wB[k], e[i], b2); }
                                    else { a3 = fmaf(wA[k], e[i], a3); b3 = fmaf(wB[k], e[i], b3); }
                                }
                            }
                        }
                        const float dA = (a0 + a1) + (a2 + a3), dB = (b0 + b1) + (b2 + b3);
                        const float sA = sigmoidf_(base + dA);                    // r (lanes 0..15) / z (lanes 16..31)
                        MhA += dA; Mnh += dB;
                        const float z = __shfl_down_sync(ODPD_FULL, sA, 16);
                        const float n = tanhf_(fmaf(sA, Mnh, xn));
                        h = fmaf(z, h - n, n);                                    // z*h + (1-z)*n
                        p_s = sA; p_z = z; p_n = n; p_m = Mnh; p_h = h; p_d = dh;
                        prow = row;
                        row += ROW;
                        cur ^= 1;
                        xa = nxa; xn = nxn;
                    }
                    if (wr) { prow[0] = p_s; prow[HP] = p_z; prow[2 * HP] = p_n; prow[3 * HP] = p_m; prow[4 * HP] = p_h; prow[5 * HP] = p_d; }
                    __syncwarp();
                    fence_async_smem();
                }
                __syncthreads();
            }
        } else {
        float whr[HT], whz[HT], whn[HT];
#pragma unroll
            for (int k = 0; k < HT; ++k) {
                const bool ok = act && k < H;
                whr[k] = ok ? sp[L.oWhh + (0 * H + j) * H + k] : 0.f;
                whz[k] = ok ? sp[L.oWhh + (1 * H + j) * H + k] : 0.f;
                whn[k] = ok ? sp[L.oWhh + (2 * H + j) * H + k] : 0.f;
            }
            float h = 0.f, hh = 0.f, Mhr = 0.f, Mhz = 0.f;
            float Mnh = (!TRES && act) ? sp[L.obhh + 2 * H + j] : 0.f;
            long long zh = 0;
            int cur = 0;
            for (int s = 0; s < nchunks + 2; ++s) {
                const int c = s - 1;
                if (c >= 0 && c < nchunks) {
                    const int t0 = c * CH, nt = min(CH, T - t0);
                    const float *xp = sxp + (c & 1) * SM::XP + lp;
                    float *ac = sact + (c & 1) * SM::ACT;
                    float xr = xp[0], xz = xp[HP], xn = xp[2 * HP];
                    for (int tl = 0; tl < nt; ++tl) {
                        const int tn = (tl + 1 < nt) ? tl + 1 : tl;
                        const float nxr = xp[tn * 3 * HP], nxz = xp[tn * 3 * HP + HP], nxn = xp[tn * 3 * HP + 2 * HP];
                        // h-side delta (compute_deltas / update_states)
                        const float d = h - hh, ad = fabsf(d);
                        const float dh = (ad < thh) ? 0.f : d;
                        const bool keep = ad >= thh;
                        if (keep) hh = h;
                        if (act) zh += (dh == 0.f);
                        const unsigned mh = __ballot_sync(ODPD_FULL, keep && act);
                        float *line = sdl + cur * HP;
                        if (lane < HP) line[lane] = act ? dh : 0.f;
                        __syncwarp();
                        float r0 = 0.f, r1 = 0.f, z0 = 0.f, z1 = 0.f, n0 = 0.f, n1 = 0.f;
                        bcast_dot<HT>(line, whr, r0, r1);
                        bcast_dot<HT>(line, whz, z0, z1);
                        bcast_dot<HT>(line, whn, n0, n1);
                        Mhr += r0 + r1; Mhz += z0 + z1; Mnh += n0 + n1;
                        const float r = sigmoidf_(xr + Mhr);
                        const float z = sigmoidf_(xz + Mhz);
                        const float n = tanhf_(fmaf(r, Mnh, xn));
                        h = fmaf(z, h, (1.f - z) * n);
                        float *row = ac + tl * ROW;
                        if (lane < HP) {
                            row[lane] = r; row[HP + lane] = z; row[2 * HP + lane] = n; row[3 * HP + lane] = Mnh; row[4 * HP + lane] = h;
                            row[5 * HP + lane] = act ? dh : 0.f;
                        }
                        if (lane == 0) row[6 * HP + 7] = __uint_as_float(mh);
                        cur ^= 1;
                        xr = nxr; xz = nxz; xn = nxn;
                    }
                    __syncwarp();
                    fence_async_smem();
                }
                __syncthreads();
            }
            if (a.stats) {
                unsigned long long tot = (unsigned long long)zh;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(ODPD_FULL, tot, o);
                if (lane == 0) {
                    atomicAdd(reinterpret_cast<unsigned long long *>(a.stats) + 2, tot);
                    atomicAdd(reinterpret_cast<unsigned long long *>(a.stats) + 3, (unsigned long long)T * H);
                }
            }
        }
    } else {
        // =============================== post: head (+skip), loss, activation store
        const float wo0 = act ? sp[L.oWo + j] : 0.f, wo1 = act ? sp[L.oWo + H + j] : 0.f;
        const float bo0 = TRES ? 0.f : sp[L.obo], bo1 = TRES ? 0.f : sp[L.obo + 1];
        const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
        float2 *o2 = reinterpret_cast<float2 *>(a.out) + (size_t)b * T;
        float *svg = a.save ? a.saved + (size_t)b * T * ROW : nullptr;
        float lsum = 0.f;
        long long zh_post = 0;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int c = s - 2;
            if (c >= 0) {
                const int t0 = c * CH, nt = min(CH, T - t0);
                float *ac = sact + (c & 1) * SM::ACT;
                const float *ft = sft + (c % 3) * SM::FT;
                if (lane < nt) {   // complete the rows: masked delta_x and mask_x bits from the pre warp
                    const float4 *s4 = reinterpret_cast<const float4 *>(ft + lane * 8);
                    const float4 d0 = s4[0], d1 = s4[1];
                    float *row = ac + lane * ROW + 6 * HP;
                    *reinterpret_cast<float4 *>(row) = d0;
                    row[4] = d1.x; row[5] = d1.y; row[6] = d1.z;
                }
                if constexpr (HP <= 16) {
                    // keep mask and zero count of delta_h, derived from the stored masked delta_h (lane = unit): with thh > 0 a unit is
                    // kept iff its masked delta is non-zero (|d| >= thh > 0); with thh <= 0 every unit is kept (deltagru.py:176-183)
                    const bool valid = lane < H;
                    const float *dcol = ac + 5 * HP + (valid ? lane : 0);
                    unsigned zc = 0;
#pragma unroll 4
                    for (int tl = 0; tl < nt; ++tl) {
                        const float dv = dcol[tl * ROW];
                        const unsigned mh = __ballot_sync(ODPD_FULL, valid && (thh > 0.f ? dv != 0.f : true));
                        zc += __popc(__ballot_sync(ODPD_FULL, valid && dv == 0.f));
                        if (lane == 0) ac[tl * ROW + 6 * HP + 7] = __uint_as_float(mh);
                    }
                    zh_post += zc;
                }
                fence_async_smem();
                __syncwarp();
                if (svg && lane == 0) tma_store_1d(svg + (size_t)t0 * ROW, ac, (uint32_t)(nt * ROW * 4));
                linear_head_chunk(ac, ROW, 4 * HP, HP, H, nt, lane, wo0, wo1, bo0, bo1, spo,
                                  TRES ? reinterpret_cast<const float2 *>(ssk + (c % 3) * SM::SK) : nullptr, o2 + t0, y2 ? y2 + t0 : iq_none(), lsum);
                if (svg && lane == 0) tma_store_wait_read();
                __syncwarp();
            }
            __syncthreads();
        }
        if (a.loss && y2) {
            lsum = warp_sum(lsum);
            if (lane == 0) atomicAdd(a.loss, (double)lsum * (double)a.loss_scale);
        }
        if constexpr (HP <= 16) {
            if (a.stats && lane == 0) {
                atomicAdd(reinterpret_cast<unsigned long long *>(a.stats) + 2, (unsigned long long)zh_post);
                atomicAdd(reinterpret_cast<unsigned long long *>(a.stats) + 3, (unsigned long long)T * H);
            }
        }
    }
}

// ================================================================ backward
template <int HT, bool TRES, bool DW>
__global__ void __launch_bounds__(128, 1) delta_bwd_kernel(GruArgs a) {
    constexpr int HP = Pad4<HT>::value, ROW = DRow<HT>::value, F = 6;
    constexpr int FM = TRES ? FM_TRES6 : FM_DGRU6;
    using SM = DBwdSmem<HT, TRES>;
    const DeltaLayout<TRES> L(a.H);
    const int H = a.H, T = a.T;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *sact = sp + Ppad;                 // [3][CH+1][ROW]
    float *spre = sact + 3 * SM::ACT;        // [3][CH][12]  go(2) | skip-path dL/dx (2) | -
    float *sdh = spre + 3 * SM::PRE;         // [2][CH][HP]
    float *sG = sdh + 2 * SM::DH;            // [2][CH][4HP]: gM_r | gM_z | gMnh | gM_n
    float *sdf = sG + 2 * SM::G;             // [CH][8]
    float *swrap = sdf + SM::DF;             // [2] grad of (I_next,Q_next) at step T-1 -> sample 0
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.x;
    if (threadIdx.x == 0) { mbar_init(bars + 1, 1); mbar_init(bars + 2, 1); mbar_init(bars + 3, 1); }
    stage_params(sp, a.params, L.P, bars);
    const bool act = lane < H;
    const int j = act ? lane : 0, lp = lane < HP ? lane : 0;
    const int nchunks = (T + CH - 1) / CH;
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
    const float *svg = a.saved + (size_t)b * T * ROW;
    const float2 *go2 = a.gout ? reinterpret_cast<const float2 *>(a.gout) + (size_t)b * T : nullptr;
    const float2 *oi2 = a.out_in ? reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * T : nullptr;
    const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);

    if (warp == 1) {
        // =============================== pre: saved rows (TMA), dLoss/dout, head back-projection, TCN-skip backward
        const float wo0 = act ? sp[L.oWo + j] : 0.f, wo1 = act ? sp[L.oWo + H + j] : 0.f;
        float gw0[TRES ? 18 : 1], gw2[TRES ? 6 : 1], gbo0 = 0.f, gbo1 = 0.f;
        if constexpr (TRES) {
#pragma unroll
            for (int q = 0; q < 18; ++q) gw0[q] = 0.f;
#pragma unroll
            for (int q = 0; q < 6; ++q) gw2[q] = 0.f;
        }
        for (int s = 0; s < nchunks + 2; ++s) {
            if (s < nchunks) {
                const int c = nchunks - 1 - s, t0 = c * CH, nt = min(CH, T - t0), slot = s % 3;
                float *ac = sact + slot * SM::ACT, *pr = spre + slot * SM::PRE, *dh = sdh + (s & 1) * SM::DH;
                uint64_t *bar = bars + 1 + slot;
                load_rows_with_prev(ac, svg, ROW, t0, nt, lane, bar);
                if (lane < nt) {
                    const int t = t0 + lane;
                    const float2 g = load_gout(go2, oi2, y2, t, gs);
                    float sx = 0.f, sy = 0.f;
                    if constexpr (TRES) {
                        const float *w0 = sp + L.oc0, *w2 = sp + L.oc2;
                        // own step: weight grads of the two convs
                        float c1[3], c2[2];
                        tcn_point(x2, T, t, w0, w2, c1, c2);
                        if constexpr (DW) {
                            const float g2[2] = {g.x * hswish_grad(c2[0]), g.y * hswish_grad(c2[1])};
#pragma unroll
                            for (int o = 0; o < 2; ++o)
#pragma unroll
                                for (int ch = 0; ch < 3; ++ch) gw2[o * 3 + ch] = fmaf(g2[o], hswish(c1[ch]), gw2[o * 3 + ch]);
#pragma unroll
                            for (int co = 0; co < 3; ++co) {
                                const float gc1 = (w2[co] * g2[0] + w2[3 + co] * g2[1]) * hswish_grad(c1[co]);
#pragma unroll
                                for (int k = 0; k < 3; ++k) {
                                    const int tt = t + (k - 1) * 16;
                                    if (tt >= 0 && tt < T) {
                                        const float2 xv = __ldg(x2 + tt);
                                        gw0[(co * 2) * 3 + k] = fmaf(gc1, xv.x, gw0[(co * 2) * 3 + k]);
                                        gw0[(co * 2 + 1) * 3 + k] = fmaf(gc1, xv.y, gw0[(co * 2 + 1) * 3 + k]);
                                    }
                                }
                            }
                        }
                        // dL/dx through the skip path, as a gather: taps at t-(k-1)*16
                        if (a.need_dx) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                const int ts = t - (k - 1) * 16;
                                if (ts >= 0 && ts < T) {
                                    float d1[3], d2[2];
                                    tcn_point(x2, T, ts, w0, w2, d1, d2);
                                    const float2 gg = load_gout(go2, oi2, y2, ts, gs);
                                    const float g2x = gg.x * hswish_grad(d2[0]), g2y = gg.y * hswish_grad(d2[1]);
#pragma unroll
                                    for (int co = 0; co < 3; ++co) {
                                        const float gc1 = (w2[co] * g2x + w2[3 + co] * g2y) * hswish_grad(d1[co]);
                                        sx = fmaf(w0[(co * 2) * 3 + k], gc1, sx);
                                        sy = fmaf(w0[(co * 2 + 1) * 3 + k], gc1, sy);
                                    }
                                }
                            }
                        }
                    } else if constexpr (DW) {
                        gbo0 += g.x; gbo1 += g.y;
                    }
                    *reinterpret_cast<float4 *>(pr + lane * 12) = make_float4(g.x, g.y, sx, sy);
                }
                __syncwarp();
                if (lane < HP)
                    for (int tl = 0; tl < nt; ++tl) dh[tl * HP + lane] = fmaf(wo0, pr[tl * 12], wo1 * pr[tl * 12 + 1]);
                mbar_wait(bar, (uint32_t)((s / 3) & 1));
            }
            __syncthreads();
        }
        if constexpr (DW) {
            if (a.partials) {
                float *prt = a.partials + (size_t)b * L.P;
                if constexpr (TRES) {
#pragma unroll
                    for (int q = 0; q < 18; ++q) { const float v = warp_sum(gw0[q]); if (lane == 0) prt[L.oc0 + q] = v; }
#pragma unroll
                    for (int q = 0; q < 6; ++q) { const float v = warp_sum(gw2[q]); if (lane == 0) prt[L.oc2 + q] = v; }
                } else {
                    gbo0 = warp_sum(gbo0); gbo1 = warp_sum(gbo1);
                    if (lane == 0) { prt[L.obo] = gbo0; prt[L.obo + 1] = gbo1; }
                }
            }
        }
    } else if (warp == 0) {
        // =============================== chain
        float wcr[HT], wcz[HT], wcn[HT];   // column j of W_hh per gate block
#pragma unroll
        for (int k = 0; k < HT; ++k) {
            const bool ok = act && k < H;
            wcr[k] = ok ? sp[L.oWhh + (0 * H + k) * H + j] : 0.f;
            wcz[k] = ok ? sp[L.oWhh + (1 * H + k) * H + j] : 0.f;
            wcn[k] = ok ? sp[L.oWhh + (2 * H + k) * H + j] : 0.f;
        }
        float gH = 0.f, gMr = 0.f, gMz = 0.f, gMn = 0.f, gMnh = 0.f, ghh = 0.f;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 1;
            if (sc >= 0 && sc < nchunks) {
                const int c = nchunks - 1 - sc, t0 = c * CH, nt = min(CH, T - t0);
                const float *acb = sact + (sc % 3) * SM::ACT;
                const float *ac = acb + lp;
                const float *dh = sdh + (sc & 1) * SM::DH + lp;
                float *Gb = sG + (sc & 1) * SM::G;
                const float *row = ac + nt * ROW;
                float r = row[0], z = row[HP], n = row[2 * HP], mnh = row[3 * HP], hp = row[4 * HP - ROW], dht = dh[(nt - 1) * HP];
                unsigned mh = __float_as_uint(acb[nt * ROW + 6 * HP + 7]);
                for (int tl = nt - 1; tl >= 0; --tl) {
                    const int tp = tl > 0 ? tl - 1 : 0;
                    const float *rn = ac + (tp + 1) * ROW;
                    const float r_n = rn[0], z_n = rn[HP], n_n = rn[2 * HP], mnh_n = rn[3 * HP], hp_n = rn[4 * HP - ROW], dh_n = dh[tp * HP];
                    const unsigned mh_n = __float_as_uint(acb[(tp + 1) * ROW + 6 * HP + 7]);
                    gH += dht;
                    const float gz = gH * (hp - n), gn = gH * (1.f - z);
                    float ghp = gH * z;
                    const float ga = gn * (1.f - n * n);
                    gMr = fmaf(ga * mnh, r * (1.f - r), gMr);
                    gMz = fmaf(gz, z * (1.f - z), gMz);
                    gMn += ga;
                    gMnh = fmaf(ga, r, gMnh);
                    float *G = Gb + tl * 4 * HP;
                    // no `act ?` select here: lanes >= H carry exact zeros by construction (zero head and W_hh columns => gH == 0), and the
                    // select made ptxas re-load H from the constant bank on the dependent chain of every step
                    if (lane < HP) { G[lane] = gMr; G[HP + lane] = gMz; G[2 * HP + lane] = gMnh; G[3 * HP + lane] = gMn; }
                    __syncwarp();
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f;     // three H-term dots as six chains of depth H/2
                    bcast_dot<HT>(G, wcr, a0, a1);
                    bcast_dot<HT>(G + HP, wcz, a2, a3);
                    bcast_dot<HT>(G + 2 * HP, wcn, a4, a5);
                    const float gdh = ((a0 + a1) + (a2 + a3)) + (a4 + a5);
                    if ((mh >> lane) & 1u) { ghp += ghh + gdh; ghh = -gdh; }
                    gH = ghp;
                    r = r_n; z = z_n; n = n_n; mnh = mnh_n; hp = hp_n; dht = dh_n; mh = mh_n;
                }
            }
            __syncthreads();
        }
        // bias gradients of the folded initial delta memories (deltagru.py:164-170)
        if constexpr (DW && !TRES) {
            if (a.partials && act) {
                float *prt = a.partials + (size_t)b * L.P;
                prt[L.obih + lane] = gMr; prt[L.obhh + lane] = gMr;
                prt[L.obih + H + lane] = gMz; prt[L.obhh + H + lane] = gMz;
                prt[L.obih + 2 * H + lane] = gMn; prt[L.obhh + 2 * H + lane] = gMnh;
            }
        }
    } else {
        // =============================== post, two warps (FFMA issue rate of one warp is the limit):
        //   warp 2 "post-A": dL/dW_hh;   warp 3 "post-B": dL/dW_ih, head, dL/dfeatures with the x_hat routing, dL/dx
        const bool roleA = (warp == 2);
        const int fl = lane - H;                 // feature lanes H..H+5 (H+6<=32 enforced by the host)
        const bool isf = !roleA && fl >= 0 && fl < F;
        float wic[3 * HT];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int k = 0; k < HT; ++k) wic[g * HT + k] = (isf && k < H) ? sp[L.oWih + (g * H + k) * F + fl] : 0.f;
        float gwhh[DW ? 3 * HT : 1], gwih[DW ? 3 * F : 1], gwo0 = 0.f, gwo1 = 0.f;
        if constexpr (DW) {
#pragma unroll
            for (int k = 0; k < 3 * HT; ++k) gwhh[k] = 0.f;
#pragma unroll
            for (int k = 0; k < 3 * F; ++k) gwih[k] = 0.f;
        }
        float gxh = 0.f;                          // dL/dx_hat of this lane's feature
        float2 *gx2 = (a.need_dx && a.gx) ? reinterpret_cast<float2 *>(a.gx) + (size_t)b * T : nullptr;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 2;
            if (sc >= 0) {
                const int c = nchunks - 1 - sc, t0 = c * CH, nt = min(CH, T - t0);
                const float *ac = sact + (sc % 3) * SM::ACT, *pr = spre + (sc % 3) * SM::PRE, *Gb = sG + (sc & 1) * SM::G;
                if (roleA) {
                    if constexpr (DW) {
#pragma unroll 2
                        for (int tl = 0; tl < nt; ++tl) {
                            const float4 *G4 = reinterpret_cast<const float4 *>(Gb + tl * 4 * HP);
                            const float dhm = ac[(tl + 1) * ROW + 5 * HP + lp];
#pragma unroll
                            for (int k4 = 0; k4 < HP / 4; ++k4) {
                                const float4 v_r = G4[k4], v_z = G4[HP / 4 + k4], v_nh = G4[2 * (HP / 4) + k4];
                                const float kr[4] = {v_r.x, v_r.y, v_r.z, v_r.w}, kz[4] = {v_z.x, v_z.y, v_z.z, v_z.w};
                                const float knh[4] = {v_nh.x, v_nh.y, v_nh.z, v_nh.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int k = k4 * 4 + e;
                                    if (k < HT) {
                                        gwhh[k] = fmaf(kr[e], dhm, gwhh[k]);
                                        gwhh[HT + k] = fmaf(kz[e], dhm, gwhh[HT + k]);
                                        gwhh[2 * HT + k] = fmaf(knh[e], dhm, gwhh[2 * HT + k]);
                                    }
                                }
                            }
                        }
                    }
                } else {
                for (int tl = nt - 1; tl >= 0; --tl) {
                    const float *G = Gb + tl * 4 * HP;
                    const float *row = ac + (tl + 1) * ROW;
                    const float ht = row[4 * HP + lp];
                    const float4 dx0 = *reinterpret_cast<const float4 *>(row + 6 * HP), dx1 = *reinterpret_cast<const float4 *>(row + 6 * HP + 4);
                    const float dx[F] = {dx0.x, dx0.y, dx0.z, dx0.w, dx1.x, dx1.y};
                    const unsigned mx = __float_as_uint(dx1.z);
                    if constexpr (DW) {
                        const float gr = G[lp], gzv = G[HP + lp], gnx = G[3 * HP + lp];
#pragma unroll
                        for (int q = 0; q < F; ++q) {
                            gwih[q] = fmaf(gr, dx[q], gwih[q]);
                            gwih[F + q] = fmaf(gzv, dx[q], gwih[F + q]);
                            gwih[2 * F + q] = fmaf(gnx, dx[q], gwih[2 * F + q]);
                        }
                        gwo0 = fmaf(pr[tl * 12], ht, gwo0); gwo1 = fmaf(pr[tl * 12 + 1], ht, gwo1);
                    }
                    float fa0 = 0.f, fa1 = 0.f;
                    if (gx2) {
                        const float4 *G4 = reinterpret_cast<const float4 *>(G);
#pragma unroll
                        for (int k4 = 0; k4 < HP / 4; ++k4) {
                            const float4 v_r = G4[k4], v_z = G4[HP / 4 + k4], v_nx = G4[3 * (HP / 4) + k4];
                            const float kr[4] = {v_r.x, v_r.y, v_r.z, v_r.w}, kz[4] = {v_z.x, v_z.y, v_z.z, v_z.w};
                            const float knx[4] = {v_nx.x, v_nx.y, v_nx.z, v_nx.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int k = k4 * 4 + e;
                                if (k < HT) {
                                    fa0 = fmaf(wic[k], kr[e], fa0);
                                    fa1 = fmaf(wic[HT + k], kz[e], fa1);
                                    fa0 = fmaf(wic[2 * HT + k], knx[e], fa0);
                                }
                            }
                        }
                    }
                    if (isf) {
                        const float gdx = fa0 + fa1;
                        float gf = 0.f;
                        if ((mx >> fl) & 1u) { gf = gxh + gdx; gxh = -gdx; }
                        sdf[tl * 8 + fl] = gf;
                    }
                }
                __syncwarp();
                if (gx2) {
                    const int t = t0 + lane;
                    float gi = 0.f, gq = 0.f;
                    if (lane < nt) {
                        const float2 v = __ldg(x2 + t);
                        float gf[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) gf[q] = (q < F) ? sdf[lane * 8 + q] : 0.f;
                        features_bwd<FM>(v.x, v.y, gf, gi, gq);
                        if constexpr (TRES) {
                            gi += pr[lane * 12 + 2]; gq += pr[lane * 12 + 3];                 // skip path
                            if (lane >= 1) { gi += sdf[(lane - 1) * 8 + 4]; gq += sdf[(lane - 1) * 8 + 5]; }   // (I_next,Q_next) of step t-1
                            if (t == T - 1) { swrap[0] = sdf[lane * 8 + 4]; swrap[1] = sdf[lane * 8 + 5]; }   // wraps to sample 0
                        }
                    }
                    if constexpr (TRES) {
                        __syncwarp();
                        if (lane == 0 && t0 == 0) { gi += swrap[0]; gq += swrap[1]; }
                        // the first sample of the chunk handled in the PREVIOUS stage (chunk c+1) still misses this chunk's last step
                        if (lane == nt - 1 && t0 + nt < T) {
                            float2 *p = gx2 + t0 + nt;
                            float2 v = *p;
                            v.x += sdf[lane * 8 + 4]; v.y += sdf[lane * 8 + 5];
                            *p = v;
                        }
                    }
                    if (lane < nt) gx2[t] = make_float2(gi, gq);
                }
                __syncwarp();
                }   // post-B
            }
            __syncthreads();
        }
        if constexpr (DW) {
            if (a.partials) {
                float *prt = a.partials + (size_t)b * L.P;
                if (act) {
                    if (roleA) {
#pragma unroll
                        for (int g = 0; g < 3; ++g)
#pragma unroll
                            for (int k = 0; k < HT; ++k)
                                if (k < H) prt[L.oWhh + (g * H + k) * H + lane] = gwhh[g * HT + k];
                    } else {
#pragma unroll
                        for (int g = 0; g < 3; ++g)
#pragma unroll
                            for (int q = 0; q < F; ++q) prt[L.oWih + (g * H + lane) * F + q] = gwih[g * F + q];
                        prt[L.oWo + lane] = gwo0; prt[L.oWo + H + lane] = gwo1;
                    }
                }
            }
        }
    }
}

#define ODPD_DELTA_TIERS(X) X(10) X(15) X(26)
static int delta_tier(int H) {
#define X(HTV) if (H <= HTV) return HTV;
    ODPD_DELTA_TIERS(X)
#undef X
    return -1;
}
template <int HT, bool TRES>
static int delta_launch(const GruArgs &a, int dir, bool dw, cudaStream_t st) {
    const DeltaLayout<TRES> L(a.H);
    const int Ppad = (L.P + 3) & ~3;
    if (dir == 0) {
        const size_t smem = (size_t)DFwdSmem<HT, TRES>::total(Ppad) * 4;
        auto k = delta_fwd_kernel<HT, TRES>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<a.B, 96, smem, st>>>(a);
    } else {
        const size_t smem = (size_t)DBwdSmem<HT, TRES>::total(Ppad) * 4;
        if (dw) { auto k = delta_bwd_kernel<HT, TRES, true>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<<<a.B, 128, smem, st>>>(a); }
        else { auto k = delta_bwd_kernel<HT, TRES, false>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<<<a.B, 128, smem, st>>>(a); }
    }
    return check_launch("delta kernel");
}
int64_t delta_saved_floats(int cell, int B, int T, int H) {
    const int ht = delta_tier(H);
    return ht < 0 ? -1 : (int64_t)B * T * (6 * ((ht + 3) & ~3) + 8);
}
int delta_run(const GruArgs &a, int dir, bool dw, cudaStream_t st) {
    const bool tres = a.cell == ODPD_CELL_TRES;
#define X(HTV) if (a.H <= HTV) return tres ? delta_launch<HTV, true>(a, dir, dw, st) : delta_launch<HTV, false>(a, dir, dw, st);
    ODPD_DELTA_TIERS(X)
#undef X
    set_error("delta-GRU kernels support hidden_size <= 26 (got %d)", a.H);
    return -1;
}

}  // namespace odpd

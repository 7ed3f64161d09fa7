// gru_family.cu — fused forward / backward for the nn.GRU-based backbones: GRU, DGRU, QGRU, QGRU_AMP1.
//
// Replaces (reference, file:line): backbones/gru.py:45-48, backbones/dgru.py:59-74, backbones/qgru.py:59-71,
// backbones/qgru_amp1.py:59-76 and the ATen GRU cell behind torch.nn.GRU (gate order r,z,n;
// n = tanh(W_in x + b_in + r*(W_hn h + b_hn)); h' = (h-n)*z + n), plus nn.MSELoss (project.py:262-272).
//
// Mapping (B200).  The path is a T-step serial recurrence per sequence, so everything is organised around the
// length of ONE timestep's dependent chain.  One CTA = one sequence = three specialised warps that form a
// software pipeline over 32-step chunks (double/triple-buffered in shared memory, one __syncthreads per chunk):
//     warp 1 "pre"   : coalesced float2 loads of the IQ stream, feature extraction (IEEE sqrt/div), input
//                      projection W_ih f_t + b for the whole chunk (no recurrence -> fully pipelined);
//                      backward: TMA bulk load of the saved activations, dLoss/dout, fc_hid back-projection
//     warp 0 "chain" : the recurrence only.  Lane j owns hidden unit j: weight rows/columns in registers, h in a
//                      register, broadcast through shared memory (1 STS + H/4 LDS.128), 2 MUFU per gate
//     warp 2 "post"  : output head + squared error + TMA bulk store of the activations (forward);
//                      weight-gradient accumulation in registers, dL/dfeatures -> dL/dx (backward)
// The flat parameter block is staged once per CTA into shared memory by a TMA bulk copy.
#include <cstdlib>
#include "cells.h"
#include "pipeline.cuh"
#include "chunking.cuh"

namespace odpd {

template <int FM, int HEAD>
struct GruLayout {
    static constexpr int F = FeatN<FM>::value;
    int H, O, oWih, oWhh, obih, obhh, oWo, obo, oWh, obh, P;
    __host__ __device__ explicit GruLayout(int h) {
        H = h; O = HEAD ? h + F : h;
        oWih = 0; oWhh = 3 * h * F; obih = oWhh + 3 * h * h; obhh = obih + 3 * h; oWo = obhh + 3 * h;
        obo = oWo + 2 * O; oWh = obo + 2; obh = oWh + h * h; P = HEAD ? obh + h : obo + 2;
    }
};

// activation row saved per timestep: r | z | n | hgn(=W_hn h + b_hn) | h_t | g(=relu(fc_hid), DGRU only), each HP floats
template <int HT, int HEAD> struct Row { static constexpr int NS = HEAD ? 6 : 5; static constexpr int value = NS * Pad4<HT>::value; };

// shared-memory carve-up (floats).  [0,16): 4 mbarriers (params + 3 activation slots)
template <int HT, int FM, int HEAD>
struct FwdSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = Row<HT, HEAD>::value, F = FeatN<FM>::value;
    static constexpr int XPP = 3 * HP + 4;         // row pitch of the input projection: lane-per-timestep 16-byte stores hit distinct banks
    static constexpr int XP = (CH + 1) * XPP;      // input projection of one chunk (+1 row: the chain's last-step prefetch reads one row past
                                                   // the chunk; the spare row keeps that read inside its own buffer, away from the pre warp's writes)
    static constexpr int FT = CH * 8;          // features of one chunk
    static constexpr int ACT = CH * ROW;       // activation rows of one chunk
    // zero-padded, 16-byte aligned copies of the weights the lane-per-timestep helper warps broadcast-read:
    //   WI [F][3*HP] (k-major) | BI [3*HP] | WH [HT][HP] + BH [HP] (DGRU head) | WO [2][HP]
    // LEAN (hidden tiers above 16): the helper warps work lane-per-timestep on these copies.  Up to 16 units the lane-per-unit helpers
    // are kept — measured on the headline (DGRU H13, 8 chunks x 4 CTAs/SM): lean helpers 59.2 us, lane-per-unit 55.6 us per forward
    // (their bursts of conflicting 16-byte row accesses cost the co-resident chain warps more than the shorter streams save); above
    // 16 units the lane-per-unit sweeps get longer and the smaller footprint (no [CH][33] transposition buffers) buys a CTA per SM.
    static constexpr bool LEAN = HP > 16;
    static constexpr int WI = LEAN ? F * 3 * HP : 0, BI = LEAN ? 3 * HP : 0, WH = (LEAN && HEAD) ? HT * HP + HP : 0, WO = LEAN ? 2 * HP : 0;
    static constexpr int PO = LEAN ? 0 : CH * 33;
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + HP + 2 * XP + 3 * FT + 2 * ACT + WI + BI + WH + WO + 2 * PO; }
};
template <int HT, int HEAD>
struct BwdSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = Row<HT, HEAD>::value;
    static constexpr int ACT = (CH + 1) * ROW;  // chunk rows + the row before the chunk
    static constexpr int PRE = CH * 12;         // features(8) + dLoss/dout(2) + pad per step
    static constexpr int DH = CH * HP;
    static constexpr int G = CH * 4 * HP;       // ar | az | an*r | an
    static constexpr int DF = CH * 8;
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + 3 * ACT + 3 * PRE + 2 * DH + 3 * DH + 2 * G + DF; }
};

// ================================================================ forward
template <int HT, int FM, int HEAD>
__global__ void __launch_bounds__(96, 1) gru_fwd_kernel(GruArgs a) {
    pdl_enter();
    constexpr int F = FeatN<FM>::value, HP = Pad4<HT>::value, ROW = Row<HT, HEAD>::value;
    using SM = FwdSmem<HT, FM, HEAD>;
    constexpr int XPP = SM::XPP;
    constexpr bool LEAN = SM::LEAN;
    const GruLayout<FM, HEAD> L(a.H);
    const int H = a.H, T = a.T;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *zero = sp + Ppad;                 // [HP] zeros: h_{-1}
    float *sxp = zero + HP;                  // [2][CH+1][XPP]
    float *sft = sxp + 2 * SM::XP;           // [3][CH][8]
    float *sact = sft + 3 * SM::FT;          // [2][CH][ROW]
    float *sWi = sact + 2 * SM::ACT;         // [F][3*HP]   W_ih, k-major, gate-major columns g*HP + j
    float *sbi = sWi + SM::WI;               // [3*HP]      b_ih (+ b_hh for r, z)
    float *sWh = sbi + SM::BI;               // [HT][HP] + [HP]   fc_hid (DGRU)
    float *sWo = sWh + SM::WH;               // [2][HP]     fc_out columns of the hidden part
    float *spo = sWo + SM::WO;               // [2][CH][33] (lane-per-unit post warp only)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const FwdRange R = fwd_range(a);          // chunking.cuh: warm-up [t_lo, t_emit) from h = 0, emitted steps [t_emit, t_hi)
    const bool spec = R.spec;
    const int b = R.b, cc = R.cc, t_lo = R.t_lo, t_emit = R.t_emit, t_hi = R.t_hi;
    if (a.mode == 2 && fwd_verify_pass(a, b, HP, HP, H)) return;   // every chunk boundary of this sequence holds

    stage_params(sp, a.params, L.P, bars);
    if (threadIdx.x < HP) zero[threadIdx.x] = 0.f;
    if constexpr (SM::LEAN) {
    for (int i = threadIdx.x; i < SM::WI; i += blockDim.x) {
        const int k = i / (3 * HP), o = i - k * 3 * HP, g = o / HP, jj = o - g * HP;
        sWi[i] = jj < H ? sp[L.oWih + (g * H + jj) * F + k] : 0.f;
    }
    for (int o = threadIdx.x; o < 3 * HP; o += blockDim.x) {
        const int g = o / HP, jj = o - g * HP;
        sbi[o] = jj < H ? sp[L.obih + g * H + jj] + (g < 2 ? sp[L.obhh + g * H + jj] : 0.f) : 0.f;
    }
    if constexpr (HEAD) {
        for (int i = threadIdx.x; i < HT * HP; i += blockDim.x) {
            const int jj = i / HP, k = i - jj * HP;
            sWh[i] = (jj < H && k < H) ? sp[L.oWh + jj * H + k] : 0.f;
        }
        for (int i = threadIdx.x; i < HP; i += blockDim.x) sWh[HT * HP + i] = i < H ? sp[L.obh + i] : 0.f;
    }
    for (int i = threadIdx.x; i < 2 * HP; i += blockDim.x) {
        const int c = i / HP, jj = i - c * HP;
        sWo[i] = jj < H ? sp[L.oWo + c * L.O + jj] : 0.f;
    }
    }
    __syncthreads();

    const bool act = lane < H;
    const int j = act ? lane : 0;
    const int cb = t_lo / CH, nchunks = (t_hi + CH - 1) / CH - cb;   // pipeline stages q = 0..nchunks-1 cover 32-step blocks cb+q
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
    // fixed roles per warp (measured: rotating the roles of co-resident CTAs over the sub-partitions is 20-50 % slower)
    const int role = warp;

    if (role == 1) {
        if constexpr (LEAN) {
            // =============================== pre: features + input projection, one chunk ahead of the chain.  Lane = timestep: every lane
            // turns its sample into features and all 3*HP projections (weights broadcast from shared memory, 16 bytes at a time) — a
            // quarter of the instructions of a lane-per-unit sweep over the 32 steps, which matters because this warp shares an issue
            // port with the chain warps of the CTAs resident on the same SM.
            for (int s = 0; s < nchunks + 2; ++s) {
                if (s < nchunks) {
                    const int t0 = (cb + s) * CH, nt = min(CH, t_hi - t0);
                    float *ft = sft + (s % 3) * SM::FT, *xp = sxp + (s & 1) * SM::XP;
                    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (lane < nt) {
                        const float2 v = __ldg(x2 + t0 + lane);
                        features_fwd<FM>(v.x, v.y, 0.f, 0.f, f);
                    }
                    float4 *d = reinterpret_cast<float4 *>(ft + lane * 8);
                    d[0] = make_float4(f[0], f[1], f[2], f[3]);
                    d[1] = make_float4(f[4], f[5], f[6], f[7]);
                    float4 *o4 = reinterpret_cast<float4 *>(xp + lane * XPP);
                    const float4 *b4 = reinterpret_cast<const float4 *>(sbi);
    #pragma unroll
                    for (int q = 0; q < 3 * HP / 4; ++q) {
                        float4 acc = b4[q];
    #pragma unroll
                        for (int k = 0; k < F; ++k) {
                            const float4 w = *reinterpret_cast<const float4 *>(sWi + k * 3 * HP + 4 * q);
                            acc.x = fmaf(w.x, f[k], acc.x); acc.y = fmaf(w.y, f[k], acc.y); acc.z = fmaf(w.z, f[k], acc.z); acc.w = fmaf(w.w, f[k], acc.w);
                        }
                        o4[q] = acc;
                    }
                }
                __syncthreads();
            }
        } else {
            // =============================== pre: features + input projection, one chunk ahead of the chain
            float wir[F], wiz[F], win[F];
    #pragma unroll
            for (int f = 0; f < F; ++f) {
                wir[f] = act ? sp[L.oWih + (0 * H + j) * F + f] : 0.f;
                wiz[f] = act ? sp[L.oWih + (1 * H + j) * F + f] : 0.f;
                win[f] = act ? sp[L.oWih + (2 * H + j) * F + f] : 0.f;
            }
            const float b_r = act ? sp[L.obih + j] + sp[L.obhh + j] : 0.f;
            const float b_z = act ? sp[L.obih + H + j] + sp[L.obhh + H + j] : 0.f;
            const float b_in = act ? sp[L.obih + 2 * H + j] : 0.f;
            for (int s = 0; s < nchunks + 2; ++s) {
                if (s < nchunks) {
                    const int t0 = (cb + s) * CH, nt = min(CH, t_hi - t0);
                    float *ft = sft + (s % 3) * SM::FT, *xp = sxp + (s & 1) * SM::XP;
                    if (lane < nt) {
                        const float2 v = __ldg(x2 + t0 + lane);
                        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        features_fwd<FM>(v.x, v.y, 0.f, 0.f, f);
                        float4 *d = reinterpret_cast<float4 *>(ft + lane * 8);
                        d[0] = make_float4(f[0], f[1], f[2], f[3]);
                        d[1] = make_float4(f[4], f[5], f[6], f[7]);
                    }
                    __syncwarp();
                    if (lane < HP) {
    #pragma unroll 4
                        for (int tl = 0; tl < nt; ++tl) {
                            const float4 *fp = reinterpret_cast<const float4 *>(ft + tl * 8);
                            const float4 f0 = fp[0];
                            float feat[8] = {f0.x, f0.y, f0.z, f0.w, 0.f, 0.f, 0.f, 0.f};
                            if (F > 4) { const float4 f1 = fp[1]; feat[4] = f1.x; feat[5] = f1.y; feat[6] = f1.z; feat[7] = f1.w; }
                            float xr = b_r, xz = b_z, xn = b_in;
    #pragma unroll
                            for (int f = 0; f < F; ++f) { xr = fmaf(wir[f], feat[f], xr); xz = fmaf(wiz[f], feat[f], xz); xn = fmaf(win[f], feat[f], xn); }
                            float *o = xp + tl * XPP + lane;
                            o[0] = xr; o[HP] = xz; o[2 * HP] = xn;
                        }
                    }
                }
                __syncthreads();
            }
        }
    } else if (role == 0) {
        // =============================== chain: the serial recurrence
        if constexpr (HP <= 16) {
            // Two half-warps share one timestep: lanes 0..15 carry the r and n rows of unit u, lanes 16..31 the z row of the same
            // unit, so every lane issues 2H FMAs + one sigmoid instead of 3H + two; z comes down with one shuffle, off the r->n path.
            const int half = lane >> 4, u = lane & 15;
            const bool au = u < H;
            const int ju = au ? u : 0;
            float wA[HT], wB[HT];
#pragma unroll
            for (int k = 0; k < HT; ++k) {
                const bool ok = au && k < H;
                wA[k] = ok ? sp[L.oWhh + ((half ? 1 : 0) * H + ju) * H + k] : 0.f;      // r row (lower) / z row (upper)
                wB[k] = (ok && !half) ? sp[L.oWhh + (2 * H + ju) * H + k] : 0.f;        // n row (lower only)
            }
            const float b_hn = (au && !half) ? sp[L.obhh + 2 * H + ju] : 0.f;
            const int up = u < HP ? u : 0;
            float h = 0.f;
            for (int s = 0; s < nchunks + 2; ++s) {
                const int c = s - 1;
                if (c >= 0 && c < nchunks) {
                    const int t0 = (cb + c) * CH, nt = min(CH, t_hi - t0);
                    if (spec && cc > 0 && t0 == t_emit && lane < HP) a.sc_guess[(size_t)blockIdx.x * HP + lane] = h;
                    const float *xpA = sxp + (c & 1) * SM::XP + (half ? HP : 0) + up;     // xr (lower) / xz (upper)
                    const float *xpN = sxp + (c & 1) * SM::XP + 2 * HP + up;
                    float *ac = sact + (c & 1) * SM::ACT;
                    const float *hrow = (c == 0) ? zero : sact + ((c - 1) & 1) * SM::ACT + (CH - 1) * ROW + 4 * HP;
                    float xa = xpA[0], xn = xpN[0];
                    float pr_ = 0.f, pz_ = 0.f, pn_ = 0.f, phg_ = 0.f;
                    // pointer-increment form: the loop body carries no index arithmetic.  The last step's prefetch reads the spare
                    // row (CH) of this chunk's own projection buffer; the value is never used.
                    const bool wr = lane < HP;
                    float *row = ac + lane;                 // row[k*HP] = slot k of this lane's unit
                    const float *xa_p = xpA, *xn_p = xpN;
#pragma unroll 1
                    for (int tl = 0; tl < nt; ++tl) {
                        const float4 *hb4 = reinterpret_cast<const float4 *>(hrow);
                        float4 hv[HP / 4];
#pragma unroll
                        for (int k4 = 0; k4 < HP / 4; ++k4) hv[k4] = hb4[k4];
                        xa_p += XPP; xn_p += XPP;
                        const float nxa = *xa_p, nxn = *xn_p;
                        if (wr && tl > 0) { row[-ROW] = pr_; row[HP - ROW] = pz_; row[2 * HP - ROW] = pn_; row[3 * HP - ROW] = phg_; }
                        float a0 = xa, a1 = 0.f, b0 = b_hn, b1 = 0.f;
#pragma unroll
                        for (int k4 = 0; k4 < HP / 4; ++k4) {
                            const float hk[4] = {hv[k4].x, hv[k4].y, hv[k4].z, hv[k4].w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int k = k4 * 4 + e;
                                if (k < HT) {
                                    if (k & 1) { a1 = fmaf(wA[k], hk[e], a1); b1 = fmaf(wB[k], hk[e], b1); }
                                    else       { a0 = fmaf(wA[k], hk[e], a0); b0 = fmaf(wB[k], hk[e], b0); }
                                }
                            }
                        }
                        const float sA = sigmoidf_(a0 + a1);                       // r in lanes 0..15, z in lanes 16..31
                        const float hgn = b0 + b1;
                        const float z = __shfl_down_sync(ODPD_FULL, sA, 16);       // lower lanes: z of their unit
                        const float n = tanhf_(fmaf(sA, hgn, xn));
                        h = fmaf(h - n, z, n);
                        if (wr) row[4 * HP] = h;
                        hrow = row + 4 * HP - lane;
                        row += ROW;
                        pr_ = sA; pz_ = z; pn_ = n; phg_ = hgn;
                        xa = nxa; xn = nxn;
                        __syncwarp();
                    }
                    if (lane < HP) {
                        float *prow = ac + (nt - 1) * ROW;
                        prow[lane] = pr_; prow[HP + lane] = pz_; prow[2 * HP + lane] = pn_; prow[3 * HP + lane] = phg_;
                    }
                    __syncwarp();
                    fence_async_smem();
                }
                __syncthreads();
            }
            if (spec && lane < HP) a.sc_end[(size_t)blockIdx.x * HP + lane] = h;
        } else {
        float whr[HT], whz[HT], whn[HT];
#pragma unroll
            for (int k = 0; k < HT; ++k) {
                const bool ok = act && k < H;
                whr[k] = ok ? sp[L.oWhh + (0 * H + j) * H + k] : 0.f;
                whz[k] = ok ? sp[L.oWhh + (1 * H + j) * H + k] : 0.f;
                whn[k] = ok ? sp[L.oWhh + (2 * H + j) * H + k] : 0.f;
            }
            const float b_hn = act ? sp[L.obhh + 2 * H + j] : 0.f;
            const int lp = lane < HP ? lane : 0;  // clamp: lanes >= HP read lane 0's slot and never write
            float h = 0.f;
            for (int s = 0; s < nchunks + 2; ++s) {
                const int c = s - 1;
                if (c >= 0 && c < nchunks) {
                    const int t0 = (cb + c) * CH, nt = min(CH, t_hi - t0);
                    if (spec && cc > 0 && t0 == t_emit && lane < HP) a.sc_guess[(size_t)blockIdx.x * HP + lane] = h;
                    const float *xp = sxp + (c & 1) * SM::XP + lp;
                    float *ac = sact + (c & 1) * SM::ACT;
                    const float *hrow = (c == 0) ? zero : sact + ((c - 1) & 1) * SM::ACT + (CH - 1) * ROW + 4 * HP;
                    float xr = xp[0], xz = xp[HP], xn = xp[2 * HP];
                    float pr_ = 0.f, pz_ = 0.f, pn_ = 0.f, phg_ = 0.f;   // gate values of the previous step, stored one iteration late
                    for (int tl = 0; tl < nt; ++tl) {
                        // broadcast h_{t-1}: issue the loads first, everything below that does not need them fills the latency
                        const float4 *hb4 = reinterpret_cast<const float4 *>(hrow);
                        float4 hv[HP / 4];
#pragma unroll
                        for (int k4 = 0; k4 < HP / 4; ++k4) hv[k4] = hb4[k4];
                        // prefetch next step's input projection (independent of h)
                        const int tn = (tl + 1 < nt) ? tl + 1 : tl;
                        const float nxr = xp[tn * XPP], nxz = xp[tn * XPP + HP], nxn = xp[tn * XPP + 2 * HP];
                        // deferred activation stores of step tl-1 (off the dependent chain)
                        if (tl > 0 && lane < HP) {
                            float *prow = ac + (tl - 1) * ROW;
                            prow[lane] = pr_; prow[HP + lane] = pz_; prow[2 * HP + lane] = pn_; prow[3 * HP + lane] = phg_;
                        }
                        float ar0 = xr, ar1 = 0.f, az0 = xz, az1 = 0.f, an0 = b_hn, an1 = 0.f;
#pragma unroll
                        for (int k4 = 0; k4 < HP / 4; ++k4) {
                            const float hk[4] = {hv[k4].x, hv[k4].y, hv[k4].z, hv[k4].w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int k = k4 * 4 + e;
                                if (k < HT) {
                                    if (k & 1) { ar1 = fmaf(whr[k], hk[e], ar1); an1 = fmaf(whn[k], hk[e], an1); az1 = fmaf(whz[k], hk[e], az1); }
                                    else       { ar0 = fmaf(whr[k], hk[e], ar0); an0 = fmaf(whn[k], hk[e], an0); az0 = fmaf(whz[k], hk[e], az0); }
                                }
                            }
                        }
                        const float r = sigmoidf_(ar0 + ar1);
                        const float hgn = an0 + an1;
                        const float z = sigmoidf_(az0 + az1);
                        const float n = tanhf_(fmaf(r, hgn, xn));
                        h = fmaf(h - n, z, n);
                        float *row = ac + tl * ROW;
                        if (lane < HP) row[4 * HP + lane] = h;
                        hrow = row + 4 * HP;
                        pr_ = r; pz_ = z; pn_ = n; phg_ = hgn;
                        xr = nxr; xz = nxz; xn = nxn;
                        __syncwarp();
                    }
                    if (lane < HP) {
                        float *prow = ac + (nt - 1) * ROW;
                        prow[lane] = pr_; prow[HP + lane] = pz_; prow[2 * HP + lane] = pn_; prow[3 * HP + lane] = phg_;
                    }
                    __syncwarp();
                    fence_async_smem();   // rows of this chunk are bulk-stored by the post warp next stage
                }
                __syncthreads();
            }
            if (spec && lane < HP) a.sc_end[(size_t)blockIdx.x * HP + lane] = h;
        }
    } else {
        if constexpr (LEAN) {
            // =============================== post: head, output, squared error, activation store.  Lane = timestep: each lane reads the h
            // row of its step once, runs the head on it with broadcast weights and writes g back into the row for the bulk store.
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
                    const float *ft = sft + (c % 3) * SM::FT;
                    float *row = ac + lane * ROW;              // rows nt..31 hold stale data: computed on, never emitted
                    float hv[HP];
    #pragma unroll
                    for (int k4 = 0; k4 < HP / 4; ++k4) {
                        const float4 v = *reinterpret_cast<const float4 *>(row + 4 * HP + 4 * k4);
                        hv[4 * k4] = v.x; hv[4 * k4 + 1] = v.y; hv[4 * k4 + 2] = v.z; hv[4 * k4 + 3] = v.w;
                    }
                    float o0 = bo0, o1 = bo1;
                    if constexpr (HEAD) {
                        float g[HP];
    #pragma unroll
                        for (int jj = 0; jj < HP; ++jj) g[jj] = 0.f;
    #pragma unroll
                        for (int jj = 0; jj < HT; ++jj) {
                            float p0 = sWh[HT * HP + jj], p1 = 0.f;
    #pragma unroll
                            for (int k4 = 0; k4 < HP / 4; ++k4) {
                                const float4 w = *reinterpret_cast<const float4 *>(sWh + jj * HP + 4 * k4);
                                p0 = fmaf(w.x, hv[4 * k4], p0); p1 = fmaf(w.y, hv[4 * k4 + 1], p1);
                                p0 = fmaf(w.z, hv[4 * k4 + 2], p0); p1 = fmaf(w.w, hv[4 * k4 + 3], p1);
                            }
                            g[jj] = fmaxf(p0 + p1, 0.f);
                        }
    #pragma unroll
                        for (int k4 = 0; k4 < HP / 4; ++k4)
                            *reinterpret_cast<float4 *>(row + 5 * HP + 4 * k4) = make_float4(g[4 * k4], g[4 * k4 + 1], g[4 * k4 + 2], g[4 * k4 + 3]);
    #pragma unroll
                        for (int k4 = 0; k4 < HP / 4; ++k4) {
                            const float4 w0 = *reinterpret_cast<const float4 *>(sWo + 4 * k4), w1 = *reinterpret_cast<const float4 *>(sWo + HP + 4 * k4);
                            o0 = fmaf(w0.x, g[4 * k4], o0); o0 = fmaf(w0.y, g[4 * k4 + 1], o0); o0 = fmaf(w0.z, g[4 * k4 + 2], o0); o0 = fmaf(w0.w, g[4 * k4 + 3], o0);
                            o1 = fmaf(w1.x, g[4 * k4], o1); o1 = fmaf(w1.y, g[4 * k4 + 1], o1); o1 = fmaf(w1.z, g[4 * k4 + 2], o1); o1 = fmaf(w1.w, g[4 * k4 + 3], o1);
                        }
    #pragma unroll
                        for (int f = 0; f < F; ++f) {
                            const float fv = ft[lane * 8 + f];
                            o0 = fmaf(sp[L.oWo + H + f], fv, o0);
                            o1 = fmaf(sp[L.oWo + L.O + H + f], fv, o1);
                        }
                    } else {
    #pragma unroll
                        for (int k4 = 0; k4 < HP / 4; ++k4) {
                            const float4 w0 = *reinterpret_cast<const float4 *>(sWo + 4 * k4), w1 = *reinterpret_cast<const float4 *>(sWo + HP + 4 * k4);
                            o0 = fmaf(w0.x, hv[4 * k4], o0); o0 = fmaf(w0.y, hv[4 * k4 + 1], o0); o0 = fmaf(w0.z, hv[4 * k4 + 2], o0); o0 = fmaf(w0.w, hv[4 * k4 + 3], o0);
                            o1 = fmaf(w1.x, hv[4 * k4], o1); o1 = fmaf(w1.y, hv[4 * k4 + 1], o1); o1 = fmaf(w1.z, hv[4 * k4 + 2], o1); o1 = fmaf(w1.w, hv[4 * k4 + 3], o1);
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (svg && lane == 0) tma_store_1d(svg + (size_t)t0 * ROW, ac, (uint32_t)(nt * ROW * 4));
                    if (lane < nt) {
                        o2[t0 + lane] = make_float2(o0, o1);
                        if (y2) {
                            const float2 y = __ldg(y2 + t0 + lane);
                            const float d0 = o0 - y.x, d1 = o1 - y.y;
                            lsum = fmaf(d0, d0, fmaf(d1, d1, lsum));
                        }
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
        } else {
            // =============================== post: head, output, squared error, activation store
            float wh[HEAD ? HT : 1];
            if constexpr (HEAD) {
    #pragma unroll
                for (int k = 0; k < HT; ++k) wh[k] = (act && k < H) ? sp[L.oWh + j * H + k] : 0.f;
            }
            const float wo0 = act ? sp[L.oWo + j] : 0.f, wo1 = act ? sp[L.oWo + L.O + j] : 0.f;
            const float bh = (HEAD && act) ? sp[L.obh + j] : 0.f;
            const float bo0 = sp[L.obo], bo1 = sp[L.obo + 1];
            const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
            float2 *o2 = reinterpret_cast<float2 *>(a.out) + (size_t)b * T;
            float *svg = a.save ? a.saved + (size_t)b * T * ROW : nullptr;
            float *spo0 = spo, *spo1 = spo + SM::PO;
            float lsum = 0.f;
            for (int s = 0; s < nchunks + 2; ++s) {
                const int c = s - 2;
                if (c >= 0 && (cb + c) * CH >= t_emit) {       // warm-up blocks emit nothing
                    const int t0 = (cb + c) * CH, nt = min(CH, t_hi - t0);
                    float *ac = sact + (c & 1) * SM::ACT;
                    const float *ft = sft + (c % 3) * SM::FT;
    #pragma unroll 2
                    for (int tl = 0; tl < nt; ++tl) {
                        float *row = ac + tl * ROW;
                        float g;
                        if constexpr (HEAD) {
                            float p0 = bh, p1 = 0.f;
                            const float4 *hn4 = reinterpret_cast<const float4 *>(row + 4 * HP);
    #pragma unroll
                            for (int k4 = 0; k4 < HP / 4; ++k4) {
                                const float4 hv = hn4[k4];
                                const float hk[4] = {hv.x, hv.y, hv.z, hv.w};
    #pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int k = k4 * 4 + e;
                                    if (k < HT) { if (k & 1) p1 = fmaf(wh[k], hk[e], p1); else p0 = fmaf(wh[k], hk[e], p0); }
                                }
                            }
                            g = fmaxf(p0 + p1, 0.f);
                            if (lane < HP) row[5 * HP + lane] = g;
                        } else {
                            g = lane < HP ? row[4 * HP + lane] : 0.f;
                        }
                        spo0[tl * 33 + lane] = wo0 * g;
                        spo1[tl * 33 + lane] = wo1 * g;
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (svg && lane == 0) tma_store_1d(svg + (size_t)t0 * ROW, ac, (uint32_t)(nt * ROW * 4));
                    if (lane < nt) {
                        float o0 = bo0, o1 = bo1;
                        for (int k = 0; k < H; ++k) { o0 += spo0[lane * 33 + k]; o1 += spo1[lane * 33 + k]; }
                        if constexpr (HEAD) {
    #pragma unroll
                            for (int f = 0; f < F; ++f) {
                                const float fv = ft[lane * 8 + f];
                                o0 = fmaf(sp[L.oWo + H + f], fv, o0);
                                o1 = fmaf(sp[L.oWo + L.O + H + f], fv, o1);
                            }
                        }
                        o2[t0 + lane] = make_float2(o0, o1);
                        if (y2) {
                            const float2 y = __ldg(y2 + t0 + lane);
                            const float d0 = o0 - y.x, d1 = o1 - y.y;
                            lsum = fmaf(d0, d0, fmaf(d1, d1, lsum));
                        }
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
}

// ================================================================ backward
// SPLIT: the F "feature lanes" (which turn the gate gradients into dL/dfeatures) do not fit next to the H unit lanes in
// one warp (H+F>32); lanes 0..F-1 then serve the features.
// (capping registers for three CTAs per SM — __launch_bounds__(128, 3) — was measured and does not pay: the chunked backward is
// bound by per-SM throughput, 6 chunks x 3 CTAs/SM takes as long as 4 chunks x 2 CTAs/SM, and the serial kernel gets 15 % slower)
template <int HT, int FM, int HEAD, bool DW>
__global__ void __launch_bounds__(128, 1) gru_bwd_kernel(GruArgs a) {
    pdl_enter();
    constexpr int F = FeatN<FM>::value, HP = Pad4<HT>::value, ROW = Row<HT, HEAD>::value;
    constexpr bool SPLIT = (HT + F > 32);
    using SM = BwdSmem<HT, HEAD>;
    const GruLayout<FM, HEAD> L(a.H);
    const int H = a.H, T = a.T;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);   // [0] params, [1..3] activation slots
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *sact = sp + Ppad;                 // [3][(CH+1)][ROW]   row 0 = step t0-1
    float *spre = sact + 3 * SM::ACT;        // [3][CH][12]        feat(8) | go(2)
    float *sdh = spre + 3 * SM::PRE;         // [2][CH][HP]        dL/dh_t from the head
    float *sdp = sdh + 2 * SM::DH;           // [3][CH][HP]        dpre (DGRU head)
    float *sG = sdp + 3 * SM::DH;            // [2][CH][4*HP]
    float *sdf = sG + 2 * SM::G;             // [CH][8]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const BwdRange R = bwd_range(a);          // chunking.cuh: warm-up [t_ehi, t_hi) from dL/dh = 0, emitted steps [t_elo, t_ehi)
    const bool spec = R.spec;
    const int b = R.b, t_elo = R.t_elo, t_ehi = R.t_ehi, t_hi = R.t_hi;
    if (a.mode == 2 && bwd_verify_pass(a, b, HP, HP, H)) return;

    if (threadIdx.x == 0) { mbar_init(bars + 1, 1); mbar_init(bars + 2, 1); mbar_init(bars + 3, 1); }
    stage_params(sp, a.params, L.P, bars);   // contains the mbarrier-init fence + __syncthreads
    const int role = warp;

    const bool act = lane < H;
    const int j = act ? lane : 0;
    const int lp = lane < HP ? lane : 0;
    // 32-step blocks [cb, ce) are processed last to first; blocks >= ce_emit are warm-up
    const int cb = t_elo / CH, ce = (t_hi + CH - 1) / CH, nchunks = ce - cb, ce_emit = (t_ehi + CH - 1) / CH;
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
    const float *svg = a.saved + (size_t)b * T * ROW;

    // stage s: pre -> chunk index (nchunks-1-s), chain -> one stage later, post -> two stages later
    if (role == 1) {
        // =============================== pre
        float whT[HEAD ? HT : 1];
        if constexpr (HEAD) {
#pragma unroll
            for (int k = 0; k < HT; ++k) whT[k] = (act && k < H) ? sp[L.oWh + k * H + j] : 0.f;
        }
        const float wo0 = act ? sp[L.oWo + j] : 0.f, wo1 = act ? sp[L.oWo + L.O + j] : 0.f;
        const float2 *go2 = a.gout ? reinterpret_cast<const float2 *>(a.gout) + (size_t)b * T : nullptr;
        const float2 *oi2 = a.out_in ? reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * T : nullptr;
        const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
        const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
        for (int s = 0; s < nchunks + 2; ++s) {
            if (s < nchunks) {
                const int c = ce - 1 - s, t0 = c * CH, nt = min(CH, t_hi - t0);
                const int slot = s % 3;
                float *ac = sact + slot * SM::ACT;
                float *pr = spre + slot * SM::PRE;
                float *dh = sdh + (s & 1) * SM::DH, *dp = sdp + slot * SM::DH;
                uint64_t *bar = bars + 1 + slot;
                if (lane == 0) {
                    if (t0 > 0) {
                        const uint32_t bytes = (uint32_t)((nt + 1) * ROW * 4);
                        mbar_expect_tx(bar, bytes);
                        tma_load_1d(ac, svg + (size_t)(t0 - 1) * ROW, bytes, bar);
                    } else {
                        const uint32_t bytes = (uint32_t)(nt * ROW * 4);
                        mbar_expect_tx(bar, bytes);
                        tma_load_1d(ac + ROW, svg, bytes, bar);
                    }
                }
                if (t0 == 0) { for (int i = lane; i < ROW; i += 32) ac[i] = 0.f; }   // h_{-1} = 0
                if (lane < nt) {
                    const float2 v = __ldg(x2 + t0 + lane);
                    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    features_fwd<FM>(v.x, v.y, 0.f, 0.f, f);
                    float g0, g1;
                    if (go2) { const float2 g = __ldg(go2 + t0 + lane); g0 = g.x; g1 = g.y; }
                    else { const float2 o = __ldg(oi2 + t0 + lane), y = __ldg(y2 + t0 + lane); g0 = gs * (o.x - y.x); g1 = gs * (o.y - y.y); }
                    float4 *d = reinterpret_cast<float4 *>(pr + lane * 12);
                    d[0] = make_float4(f[0], f[1], f[2], f[3]);
                    d[1] = make_float4(f[4], f[5], f[6], f[7]);
                    d[2] = make_float4(g0, g1, 0.f, 0.f);
                }
                __syncwarp();
                mbar_wait(bar, (uint32_t)((s / 3) & 1));
                if constexpr (HEAD) {
                    if (lane < HP) {
                        for (int tl = 0; tl < nt; ++tl) {
                            const float2 go = *reinterpret_cast<const float2 *>(pr + tl * 12 + 8);
                            const float g = ac[(tl + 1) * ROW + 5 * HP + lane];
                            dp[tl * HP + lane] = g > 0.f ? fmaf(wo0, go.x, wo1 * go.y) : 0.f;
                        }
                    }
                    __syncwarp();
                    if (lane < HP) {
#pragma unroll 2
                        for (int tl = 0; tl < nt; ++tl) {
                            float d0 = 0.f, d1 = 0.f;
                            const float4 *dp4 = reinterpret_cast<const float4 *>(dp + tl * HP);
#pragma unroll
                            for (int k4 = 0; k4 < HP / 4; ++k4) {
                                const float4 dv = dp4[k4];
                                const float dk[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int k = k4 * 4 + e;
                                    if (k < HT) { if (k & 1) d1 = fmaf(whT[k], dk[e], d1); else d0 = fmaf(whT[k], dk[e], d0); }
                                }
                            }
                            dh[tl * HP + lane] = d0 + d1;
                        }
                    }
                } else {
                    if (lane < HP) {
                        for (int tl = 0; tl < nt; ++tl) {
                            const float2 go = *reinterpret_cast<const float2 *>(pr + tl * 12 + 8);
                            dh[tl * HP + lane] = fmaf(wo0, go.x, wo1 * go.y);
                        }
                    }
                }
            }
            __syncthreads();
        }
    } else if (role == 0) {
        // =============================== chain: reverse-time recurrence of dL/dh
        float wcol[3 * HT];   // column j of W_hh
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int k = 0; k < HT; ++k) wcol[g * HT + k] = (act && k < H) ? sp[L.oWhh + (g * H + k) * H + j] : 0.f;
        float gH = 0.f;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 1;
            if (sc >= 0 && sc < nchunks) {
                const int c = ce - 1 - sc, t0 = c * CH, nt = min(CH, t_hi - t0);
                if (spec && c == ce_emit - 1 && t_ehi < T && lane < HP) a.sc_guess[(size_t)blockIdx.x * HP + lane] = gH;
                const float *ac = sact + (sc % 3) * SM::ACT + lp;
                const float *dh = sdh + (sc & 1) * SM::DH + lp;
                float *Gb = sG + (sc & 1) * SM::G;
                // software-pipelined operand fetch (none of these depend on the recurrence)
                const float *row = ac + nt * ROW;   // step t0+nt-1
                float r = row[0], z = row[HP], n = row[2 * HP], hgn = row[3 * HP], hp = row[4 * HP - ROW], dht = dh[(nt - 1) * HP];
                for (int tl = nt - 1; tl >= 0; --tl) {
                    const int tp = tl > 0 ? tl - 1 : 0;
                    const float *rown = ac + (tp + 1) * ROW;
                    const float r_n = rown[0], z_n = rown[HP], n_n = rown[2 * HP], hgn_n = rown[3 * HP], hp_n = rown[4 * HP - ROW],
                                dh_n = dh[tp * HP];
                    gH += dht;
                    const float gz = gH * (hp - n), gn = gH * (1.f - z), ghp = gH * z;
                    const float an = gn * (1.f - n * n);
                    const float az = gz * z * (1.f - z);
                    const float anr = an * r;
                    const float ar = anr * hgn * (1.f - r);
                    float *G = Gb + tl * 4 * HP;
                    if (lane < HP) { G[lane] = ar; G[HP + lane] = az; G[2 * HP + lane] = anr; G[3 * HP + lane] = an; }
                    __syncwarp();
                    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
                    const float4 *G4 = reinterpret_cast<const float4 *>(G);
                    float4 gv[3 * (HP / 4)];
#pragma unroll
                    for (int q = 0; q < 3 * (HP / 4); ++q) gv[q] = G4[q];
#pragma unroll
                    for (int k4 = 0; k4 < HP / 4; ++k4) {
                        const float kr[4] = {gv[k4].x, gv[k4].y, gv[k4].z, gv[k4].w};
                        const float kz[4] = {gv[HP / 4 + k4].x, gv[HP / 4 + k4].y, gv[HP / 4 + k4].z, gv[HP / 4 + k4].w};
                        const float kn[4] = {gv[2 * (HP / 4) + k4].x, gv[2 * (HP / 4) + k4].y, gv[2 * (HP / 4) + k4].z, gv[2 * (HP / 4) + k4].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int k = k4 * 4 + e;
                            if (k < HT) {
                                if (k & 1) { acc1 = fmaf(wcol[k], kr[e], acc1); acc3 = fmaf(wcol[HT + k], kz[e], acc3); acc1 = fmaf(wcol[2 * HT + k], kn[e], acc1); }
                                else       { acc0 = fmaf(wcol[k], kr[e], acc0); acc2 = fmaf(wcol[HT + k], kz[e], acc2); acc0 = fmaf(wcol[2 * HT + k], kn[e], acc0); }
                            }
                        }
                    }
                    gH = ghp + ((acc0 + acc1) + (acc2 + acc3));
                    r = r_n; z = z_n; n = n_n; hgn = hgn_n; hp = hp_n; dht = dh_n;
                }
            }
            __syncthreads();
        }
        if (spec && lane < HP) a.sc_end[(size_t)blockIdx.x * HP + lane] = gH;
    } else {
        // =============================== post (two warps, the FFMA issue rate of one warp is the limit):
        //   warp 2 "post-A": dL/dW_hh (3H accumulators per lane)
        //   warp 3 "post-B": dL/dW_ih, biases, head gradients, dL/dfeatures -> dL/dx, time-parallel tail
        const bool roleA = (role == 2);
        const int fl = SPLIT ? lane : lane - H;   // feature column served by this lane (if 0<=fl<F)
        const bool isf = fl >= 0 && fl < F;
        float wicol[3 * HT];                      // column fl of W_ih (feature lanes, post-B)
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int k = 0; k < HT; ++k) wicol[g * HT + k] = (!roleA && isf && k < H) ? sp[L.oWih + (g * H + k) * F + fl] : 0.f;
        const float wof0 = (HEAD && isf) ? sp[L.oWo + H + fl] : 0.f, wof1 = (HEAD && isf) ? sp[L.oWo + L.O + H + fl] : 0.f;
        float gwhh[DW ? 3 * HT : 1], gwih[DW ? 3 * F : 1], gwh[(DW && HEAD) ? HT : 1], gwof[(DW && HEAD) ? 2 * F : 1];
        float gb_r = 0.f, gb_z = 0.f, gb_n = 0.f, gb_hn = 0.f, gwo0 = 0.f, gwo1 = 0.f, gbh = 0.f, gbo0 = 0.f, gbo1 = 0.f;
        if constexpr (DW) {
#pragma unroll
            for (int k = 0; k < 3 * HT; ++k) gwhh[k] = 0.f;
#pragma unroll
            for (int k = 0; k < 3 * F; ++k) gwih[k] = 0.f;
            if constexpr (HEAD) {
#pragma unroll
                for (int k = 0; k < HT; ++k) gwh[k] = 0.f;
#pragma unroll
                for (int k = 0; k < 2 * F; ++k) gwof[k] = 0.f;
            }
        }
        float2 *gx2 = (a.need_dx && a.gx) ? reinterpret_cast<float2 *>(a.gx) + (size_t)b * T : nullptr;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 2;
            if (sc >= 0 && ce - 1 - sc < ce_emit) {          // warm-up blocks emit nothing
                const int c = ce - 1 - sc, t0 = c * CH, nt = min(CH, t_hi - t0);
                const float *ac = sact + (sc % 3) * SM::ACT;
                const float *pr = spre + (sc % 3) * SM::PRE;
                const float *dp = sdp + (sc % 3) * SM::DH;
                const float *Gb = sG + (sc & 1) * SM::G;
                if (roleA) {
                    if constexpr (DW) {
#pragma unroll 2
                        for (int tl = 0; tl < nt; ++tl) {
                            const float4 *G4 = reinterpret_cast<const float4 *>(Gb + tl * 4 * HP);
                            const float hp = ac[tl * ROW + 4 * HP + lp];      // row tl = step t-1
#pragma unroll
                            for (int k4 = 0; k4 < HP / 4; ++k4) {
                                const float4 v_r = G4[k4], v_z = G4[HP / 4 + k4], v_nh = G4[2 * (HP / 4) + k4];
                                const float kr[4] = {v_r.x, v_r.y, v_r.z, v_r.w}, kz[4] = {v_z.x, v_z.y, v_z.z, v_z.w};
                                const float knh[4] = {v_nh.x, v_nh.y, v_nh.z, v_nh.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int k = k4 * 4 + e;
                                    if (k < HT) {
                                        gwhh[k] = fmaf(kr[e], hp, gwhh[k]);
                                        gwhh[HT + k] = fmaf(kz[e], hp, gwhh[HT + k]);
                                        gwhh[2 * HT + k] = fmaf(knh[e], hp, gwhh[2 * HT + k]);
                                    }
                                }
                            }
                        }
                    }
                } else {
                    for (int tl = 0; tl < nt; ++tl) {
                        const float *G = Gb + tl * 4 * HP;
                        const float4 *G4 = reinterpret_cast<const float4 *>(G);
                        const float *row = ac + (tl + 1) * ROW;
                        const float ht = row[4 * HP + lp];
                        const float2 go = *reinterpret_cast<const float2 *>(pr + tl * 12 + 8);
                        float feat[8];
                        {
                            const float4 *fp = reinterpret_cast<const float4 *>(pr + tl * 12);
                            const float4 f0 = fp[0];
                            feat[0] = f0.x; feat[1] = f0.y; feat[2] = f0.z; feat[3] = f0.w;
                            if (F > 4) { const float4 f1 = fp[1]; feat[4] = f1.x; feat[5] = f1.y; feat[6] = f1.z; feat[7] = f1.w; }
                        }
                        if constexpr (DW) {
                            const float ar = G[lp], az = G[HP + lp], anr = G[2 * HP + lp], an = G[3 * HP + lp];
#pragma unroll
                            for (int q = 0; q < F; ++q) {
                                gwih[q] = fmaf(ar, feat[q], gwih[q]);
                                gwih[F + q] = fmaf(az, feat[q], gwih[F + q]);
                                gwih[2 * F + q] = fmaf(an, feat[q], gwih[2 * F + q]);
                            }
                            gb_r += ar; gb_z += az; gb_n += an; gb_hn += anr;
                            if constexpr (HEAD) {
                                const float g = row[5 * HP + lp];
                                gwo0 = fmaf(go.x, g, gwo0); gwo1 = fmaf(go.y, g, gwo1);
                                gbh += dp[tl * HP + lp];
                                const float4 *dp4 = reinterpret_cast<const float4 *>(dp + tl * HP);
#pragma unroll
                                for (int k4 = 0; k4 < HP / 4; ++k4) {
                                    const float4 dv = dp4[k4];
                                    const float dk[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) { const int k = k4 * 4 + e; if (k < HT) gwh[k] = fmaf(dk[e], ht, gwh[k]); }
                                }
                            } else {
                                gwo0 = fmaf(go.x, ht, gwo0); gwo1 = fmaf(go.y, ht, gwo1);
                            }
                        }
                        if (a.need_dx) {
                            float fa0 = 0.f, fa1 = 0.f;
#pragma unroll
                            for (int k4 = 0; k4 < HP / 4; ++k4) {
                                const float4 v_r = G4[k4], v_z = G4[HP / 4 + k4], v_nx = G4[3 * (HP / 4) + k4];
                                const float kr[4] = {v_r.x, v_r.y, v_r.z, v_r.w}, kz[4] = {v_z.x, v_z.y, v_z.z, v_z.w};
                                const float knx[4] = {v_nx.x, v_nx.y, v_nx.z, v_nx.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int k = k4 * 4 + e;
                                    if (k < HT) {
                                        fa0 = fmaf(wicol[k], kr[e], fa0);
                                        fa1 = fmaf(wicol[HT + k], kz[e], fa1);
                                        fa0 = fmaf(wicol[2 * HT + k], knx[e], fa0);
                                    }
                                }
                            }
                            if (isf) sdf[tl * 8 + fl] = (fa0 + fa1) + (HEAD ? fmaf(wof0, go.x, wof1 * go.y) : 0.f);
                        }
                    }
                    __syncwarp();
                    // time-parallel tail of the chunk: one timestep per lane
                    if (lane < nt) {
                        const float *p = pr + lane * 12;
                        if constexpr (DW) {
                            gbo0 += p[8]; gbo1 += p[9];
                            if constexpr (HEAD) {
#pragma unroll
                                for (int q = 0; q < F; ++q) { gwof[q] = fmaf(p[8], p[q], gwof[q]); gwof[F + q] = fmaf(p[9], p[q], gwof[F + q]); }
                            }
                        }
                        if (gx2) {
                            const float2 v = __ldg(x2 + t0 + lane);
                            float gf[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) gf[q] = (q < F) ? sdf[lane * 8 + q] : 0.f;
                            float gi, gq;
                            features_bwd<FM>(v.x, v.y, gf, gi, gq);
                            gx2[t0 + lane] = make_float2(gi, gq);
                        }
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
        }
        if constexpr (DW) {
            if (a.partials) {
                float *prt = chunk_partial_row(a, spec, b, L.P, (role - 2) * 32 + lane, 64);   // both post warps
                if (roleA) {
                    if (act) {
#pragma unroll
                        for (int g = 0; g < 3; ++g)
#pragma unroll
                            for (int k = 0; k < HT; ++k)
                                if (k < H) prt[L.oWhh + (g * H + k) * H + lane] = gwhh[g * HT + k];
                    }
                } else {
                    if (act) {
#pragma unroll
                        for (int g = 0; g < 3; ++g)
#pragma unroll
                            for (int q = 0; q < F; ++q) prt[L.oWih + (g * H + lane) * F + q] = gwih[g * F + q];
                        prt[L.obih + lane] = gb_r; prt[L.obih + H + lane] = gb_z; prt[L.obih + 2 * H + lane] = gb_n;
                        prt[L.obhh + lane] = gb_r; prt[L.obhh + H + lane] = gb_z; prt[L.obhh + 2 * H + lane] = gb_hn;
                        prt[L.oWo + lane] = gwo0; prt[L.oWo + L.O + lane] = gwo1;
                        if constexpr (HEAD) {
#pragma unroll
                            for (int k = 0; k < HT; ++k)
                                if (k < H) prt[L.oWh + k * H + lane] = gwh[k];
                            prt[L.obh + lane] = gbh;
                        }
                    }
                    gbo0 = warp_sum(gbo0); gbo1 = warp_sum(gbo1);
                    if (lane == 0) { prt[L.obo] = gbo0; prt[L.obo + 1] = gbo1; }
                    if constexpr (HEAD) {
#pragma unroll
                        for (int q = 0; q < 2 * F; ++q) {
                            const float sum = warp_sum(gwof[q]);
                            if (lane == 0) prt[L.oWo + (q / F) * L.O + H + (q % F)] = sum;
                        }
                    }
                }
            }
        }
    }
}

// ================================================================ backward, split form (default)
// The fused backward above keeps four warps per sequence busy, but three of them do TIME-PARALLEL work (weight gradients,
// dL/dx) with lanes = hidden units inside the latency-critical chain kernel: 323 instructions per timestep and CTA, 219
// registers, 74 KB of shared memory => 2 CTAs per SM, 4 time chunks.  The split form separates the two kinds of work:
//   gru_bwdc_kernel  ("chain"): 2 working warps (+1 placement spacer) per (sequence, chunk).  Warp 1 loads the saved rows (TMA) and back-projects the output
//                    head with lanes = TIMESTEPS (32 steps at once instead of a 32-step loop); warp 0 runs the reverse
//                    recurrence and streams the per-step gate gradients G_t = (ar | az | an*r | an) to global memory with one
//                    TMA bulk store per 32-step block.  ~110 instructions per step, ~50 KB => 4 CTAs per SM, 8 chunks.
//   gru_bwdw_kernel  ("weights"): fully time-parallel, one CTA per (sequence, 256-step tile), no dependence between tiles:
//                    dL/dW_hh, dL/dW_ih, biases, head gradients (lanes = units, register accumulators) and dL/dx (lanes =
//                    timesteps) from G_t, the saved rows and the IQ samples; one gradient-partial row per tile.
// Chunk verification (mode 2) works on the chain kernel exactly as before; the weights kernel runs after it on the final G.
// floats per timestep of G in shared AND global memory: 4*HP + 4, so that (GS/4) is odd and one-timestep-per-lane LDS.128 reads of a
// block are bank-conflict free (the chain's own broadcast reads do not care)
template <int HT> struct GStride { static constexpr int value = 4 * Pad4<HT>::value + 4; };
template <int HT, int HEAD>
struct BwdcSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = Row<HT, HEAD>::value;
    static constexpr int ACT = (CH + 1) * ROW, DH = CH * HP, G = CH * GStride<HT>::value;
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + 2 * ACT + 2 * DH + 2 * G + DH + 2 * CH; }
};

template <int HT, int FM, int HEAD>
__global__ void __launch_bounds__(96, 1) gru_bwdc_kernel(GruArgs a) {
    pdl_enter();
    constexpr int HP = Pad4<HT>::value, ROW = Row<HT, HEAD>::value, GS = GStride<HT>::value;
    using SM = BwdcSmem<HT, HEAD>;
    const GruLayout<FM, HEAD> L(a.H);
    const int H = a.H, T = a.T;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);   // [0] params, [1..2] activation slots
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *sact = sp + Ppad;                 // [2][(CH+1)][ROW]   row 0 = step t0-1
    float *sdh = sact + 2 * SM::ACT;         // [2][CH][HP]        dL/dh_t from the head
    float *sG = sdh + 2 * SM::DH;            // [2][CH][GS]        ar | az | an*r | an | pad
    float *sdpre = sG + 2 * SM::G;           // [CH][HP]           dpre of the DGRU head (pre warp only)
    float2 *sgo = reinterpret_cast<float2 *>(sdpre + SM::DH);   // [CH] dLoss/dout (pre warp only)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const BwdRange R = bwd_range(a);
    const bool spec = R.spec;
    const int b = R.b, t_elo = R.t_elo, t_ehi = R.t_ehi, t_hi = R.t_hi;
    if (a.mode == 2 && bwd_verify_pass(a, b, HP, HP, H)) return;

    if (threadIdx.x == 0) { mbar_init(bars + 1, 1); mbar_init(bars + 2, 1); }
    stage_params(sp, a.params, L.P, bars);
    const bool act = lane < H;
    const int j = act ? lane : 0;
    const int lp = lane < HP ? lane : 0;
    const int cb = t_elo / CH, ce = (t_hi + CH - 1) / CH, nchunks = ce - cb, ce_emit = (t_ehi + CH - 1) / CH;
    const float *svg = a.saved + (size_t)b * T * ROW;
    float *gbuf = a.gbuf + (size_t)b * T * GS;
    // Warp 2 is a placement spacer and leaves here.  Warps go to sub-partition (warp slot % 4): with 2-warp CTAs the chain warps of the
    // four co-resident CTAs would pile up on two of the four sub-partitions; with 3-warp CTAs they land on four different ones.
    if (warp == 2) return;
    auto stage_sync = [] { asm volatile("bar.sync 1, 64;" ::: "memory"); };

    if (warp == 1) {
        // =============================== pre: saved rows (TMA) + head back-projection, one timestep per lane
        const float2 *go2 = a.gout ? reinterpret_cast<const float2 *>(a.gout) + (size_t)b * T : nullptr;
        const float2 *oi2 = a.out_in ? reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * T : nullptr;
        const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
        const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
        float whT[HEAD ? HT : 1];
        if constexpr (HEAD) {
#pragma unroll
            for (int k = 0; k < HT; ++k) whT[k] = (act && k < H) ? sp[L.oWh + k * H + j] : 0.f;
        }
        const float wo0 = act ? sp[L.oWo + j] : 0.f, wo1 = act ? sp[L.oWo + L.O + j] : 0.f;
        for (int s = 0; s < nchunks + 1; ++s) {
            if (s < nchunks) {
                const int c = ce - 1 - s, t0 = c * CH, nt = min(CH, t_hi - t0), slot = s & 1;
                float *ac = sact + slot * SM::ACT;
                float *dh = sdh + slot * SM::DH;
                uint64_t *bar = bars + 1 + slot;
                load_rows_with_prev(ac, svg, ROW, t0, nt, lane, bar);
                if (lane < nt) sgo[lane] = load_gout(go2, oi2, y2, t0 + lane, gs);     // one timestep per lane (coalesced)
                __syncwarp();
                mbar_wait(bar, (uint32_t)((s >> 1) & 1));
                // back-projection of the head with lane = unit: consecutive lanes read consecutive words (no bank conflicts — a
                // one-timestep-per-lane version needs 8x fewer instructions but its 32-way conflicted row reads stall the chain
                // warps of all co-resident CTAs, which share the SM's shared-memory pipe)
                if constexpr (HEAD) {
                    if (lane < HP) {
                        for (int tl = 0; tl < nt; ++tl) {
                            const float2 go = sgo[tl];
                            const float g = ac[(tl + 1) * ROW + 5 * HP + lane];
                            sdpre[tl * HP + lane] = g > 0.f ? fmaf(wo0, go.x, wo1 * go.y) : 0.f;
                        }
                    }
                    __syncwarp();
                    if (lane < HP) {
#pragma unroll 2
                        for (int tl = 0; tl < nt; ++tl) {
                            float d0 = 0.f, d1 = 0.f;
                            const float4 *dp4 = reinterpret_cast<const float4 *>(sdpre + tl * HP);
#pragma unroll
                            for (int k4 = 0; k4 < HP / 4; ++k4) {
                                const float4 dv = dp4[k4];
                                const float dk[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int k = k4 * 4 + e;
                                    if (k < HT) { if (k & 1) d1 = fmaf(whT[k], dk[e], d1); else d0 = fmaf(whT[k], dk[e], d0); }
                                }
                            }
                            dh[tl * HP + lane] = d0 + d1;
                        }
                    }
                } else {
                    if (lane < HP) {
                        for (int tl = 0; tl < nt; ++tl) {
                            const float2 go = sgo[tl];
                            dh[tl * HP + lane] = fmaf(wo0, go.x, wo1 * go.y);
                        }
                    }
                }
            }
            stage_sync();
        }
    } else {
        // =============================== chain: reverse-time recurrence of dL/dh (identical arithmetic to the fused kernel)
        float wcol[3 * HT];   // column j of W_hh
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int k = 0; k < HT; ++k) wcol[g * HT + k] = (act && k < H) ? sp[L.oWhh + (g * H + k) * H + j] : 0.f;
        float gH = 0.f;
        int stores = 0;
        for (int s = 0; s < nchunks + 1; ++s) {
            const int sc = s - 1;
            if (sc >= 0) {
                const int c = ce - 1 - sc, t0 = c * CH, nt = min(CH, t_hi - t0);
                if (spec && c == ce_emit - 1 && t_ehi < T && lane < HP) a.sc_guess[(size_t)blockIdx.x * HP + lane] = gH;
                const float *ac = sact + (sc & 1) * SM::ACT + lp;
                const float *dh = sdh + (sc & 1) * SM::DH + lp;
                float *Gb = sG + (sc & 1) * SM::G;
                // the bulk store issued two blocks ago read this buffer: it must have finished reading before we overwrite it
                if (stores >= 2) { if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); __syncwarp(); }
                const float *row = ac + nt * ROW;   // step t0+nt-1
                float r = row[0], z = row[HP], n = row[2 * HP], hgn = row[3 * HP], hp = row[4 * HP - ROW], dht = dh[(nt - 1) * HP];
                for (int tl = nt - 1; tl >= 0; --tl) {
                    const int tp = tl > 0 ? tl - 1 : 0;
                    const float *rown = ac + (tp + 1) * ROW;
                    const float r_n = rown[0], z_n = rown[HP], n_n = rown[2 * HP], hgn_n = rown[3 * HP], hp_n = rown[4 * HP - ROW],
                                dh_n = dh[tp * HP];
                    gH += dht;
                    const float gz = gH * (hp - n), gn = gH * (1.f - z), ghp = gH * z;
                    const float an = gn * (1.f - n * n);
                    const float az = gz * z * (1.f - z);
                    const float anr = an * r;
                    const float ar = anr * hgn * (1.f - r);
                    float *G = Gb + tl * GS;
                    if (lane < HP) { G[lane] = ar; G[HP + lane] = az; G[2 * HP + lane] = anr; G[3 * HP + lane] = an; }
                    __syncwarp();
                    // six accumulators: three 3H-term dot products as chains of depth ~H/2 (the fused kernel ran depth ~H)
                    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f, acc4 = 0.f, acc5 = 0.f;
                    const float4 *G4 = reinterpret_cast<const float4 *>(G);
                    float4 gv[3 * (HP / 4)];
#pragma unroll
                    for (int q = 0; q < 3 * (HP / 4); ++q) gv[q] = G4[q];
#pragma unroll
                    for (int k4 = 0; k4 < HP / 4; ++k4) {
                        const float kr[4] = {gv[k4].x, gv[k4].y, gv[k4].z, gv[k4].w};
                        const float kz[4] = {gv[HP / 4 + k4].x, gv[HP / 4 + k4].y, gv[HP / 4 + k4].z, gv[HP / 4 + k4].w};
                        const float kn[4] = {gv[2 * (HP / 4) + k4].x, gv[2 * (HP / 4) + k4].y, gv[2 * (HP / 4) + k4].z, gv[2 * (HP / 4) + k4].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int k = k4 * 4 + e;
                            if (k < HT) {
                                if (k & 1) { acc1 = fmaf(wcol[k], kr[e], acc1); acc3 = fmaf(wcol[HT + k], kz[e], acc3); acc5 = fmaf(wcol[2 * HT + k], kn[e], acc5); }
                                else       { acc0 = fmaf(wcol[k], kr[e], acc0); acc2 = fmaf(wcol[HT + k], kz[e], acc2); acc4 = fmaf(wcol[2 * HT + k], kn[e], acc4); }
                            }
                        }
                    }
                    gH = (ghp + (acc0 + acc1)) + ((acc2 + acc3) + (acc4 + acc5));
                    r = r_n; z = z_n; n = n_n; hgn = hgn_n; hp = hp_n; dht = dh_n;
                }
                if (c < ce_emit) {       // warm-up blocks emit nothing
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) tma_store_1d(gbuf + (size_t)t0 * GS, Gb, (uint32_t)(nt * GS * 4));
                    ++stores;
                }
            }
            stage_sync();
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (spec && lane < HP) a.sc_end[(size_t)blockIdx.x * HP + lane] = gH;
    }
}

template <int HT, int FM, int HEAD>
struct BwdwSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = Row<HT, HEAD>::value;
    static constexpr int ACT = (CH + 1) * ROW, G = CH * GStride<HT>::value, PRE = CH * 12, DP = CH * HP;
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + 3 * ACT + 3 * G + 2 * PRE + 2 * DP + 16; }
};

// N consecutive floats from shared memory with the widest loads the tile geometry guarantees (p is 4*gcd(N,4)-byte aligned)
template <int N>
__device__ __forceinline__ void lds_vec(const float *p, float (&v)[N]) {
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) { const float4 t = reinterpret_cast<const float4 *>(p)[q]; v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
    } else if constexpr (N % 2 == 0) {
#pragma unroll
        for (int q = 0; q < N / 2; ++q) { const float2 t = reinterpret_cast<const float2 *>(p)[q]; v[2 * q] = t.x; v[2 * q + 1] = t.y; }
    } else {
#pragma unroll
        for (int q = 0; q < N; ++q) v[q] = p[q];
    }
}
// acc[i][c] += a[i] * b[c]
template <int TR, int TC>
__device__ __forceinline__ void outer_acc(float (&acc)[TR * TC], const float (&av)[TR], const float (&bv)[TC]) {
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int c = 0; c < TC; ++c) acc[i * TC + c] = fmaf(av[i], bv[c], acc[i * TC + c]);
}

// One CTA = one (sequence, tile of a.wt_blocks 32-step blocks); tiles are independent.  Every weight gradient is a sum over time of
// an outer product of two per-timestep vectors, so the kernel is organised as small register-tiled GEMMs whose K dimension is time:
// each lane owns a TR x TC tile of the output, reads TR + TC operands per timestep (vector LDS, mostly broadcast) and issues TR*TC
// FMAs — all 32 lanes busy, instead of lanes = hidden units (13 of 32 busy for the headline model).
//   warp 0 "prep" : TMA loads of the saved rows and of G one block ahead; per timestep (one per lane) features (+ a ones column),
//                   dLoss/dout and, for the DGRU head, dpre = relu'(.) * (W_o^T dLoss/dout)
//   warp 1 "hh"   : dL/dW_hh            = sum_t [ar|az|an*r]_t (x) h_{t-1}            lanes 8 x 4, tile ceil(3HP/8) x HP/4
//   warp 2 "ih"   : dL/dW_ih, all biases = sum_t [ar|az|an*r|an]_t (x) [feat_t, 1]     lanes 16 x 2 (32 x 1 for 2 features), tile x 4
//                   head: dL/dfc_hid.W  = sum_t dpre_t (x) h_t (lanes 8 x 4);  dL/dfc_out.W[:, :H] = sum_t dLoss/dout_t (x) g_t (lane = column)
//   warp 3 "dx"   : one timestep per lane: dL/dfeatures = W_ih^T [ar|az|an] (+ head) -> dL/dx; per-timestep sums (fc_out bias and
//                   feature columns, fc_hid bias)
template <int HT, int FM, int HEAD, bool DW>
__global__ void __launch_bounds__(128, 1) gru_bwdw_kernel(GruArgs a) {
    pdl_enter();
    constexpr int F = FeatN<FM>::value, HP = Pad4<HT>::value, ROW = Row<HT, HEAD>::value, GS = GStride<HT>::value;
    using SM = BwdwSmem<HT, FM, HEAD>;
    const GruLayout<FM, HEAD> L(a.H);
    const int H = a.H, T = a.T;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);   // [0] params, [1..3] block slots
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *sact = sp + Ppad;                 // [3][(CH+1)][ROW]
    float *sG = sact + 3 * SM::ACT;          // [3][CH][GS]
    float *spre = sG + 3 * SM::G;            // [2][CH][12]   feat(F) | 1 | 0.. (8) | go(2) | -
    float *sdp = spre + 2 * SM::PRE;         // [2][CH][HP]   dpre (DGRU head)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, tile = blockIdx.x;
    const int WTs = a.wt_blocks * CH;
    const int t_lo = tile * WTs, t_hi = min(T, t_lo + WTs);
    const int nblk = (t_hi - t_lo + CH - 1) / CH;
    if (threadIdx.x == 0) { mbar_init(bars + 1, 1); mbar_init(bars + 2, 1); mbar_init(bars + 3, 1); }
    stage_params(sp, a.params, L.P, bars);
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
    const float *svg = a.saved + (size_t)b * T * ROW;
    const float *gbuf = a.gbuf + (size_t)b * T * GS;
    float *prt = (DW && a.partials) ? a.partials + ((size_t)b * gridDim.x + tile) * L.P : nullptr;

    if (warp == 0) {
        // =============================== prep
        const float2 *go2 = a.gout ? reinterpret_cast<const float2 *>(a.gout) + (size_t)b * T : nullptr;
        const float2 *oi2 = a.out_in ? reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * T : nullptr;
        const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
        const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
        auto issue = [&](int q) {      // TMA loads of block q into slot q % 3
            const int t0 = t_lo + q * CH, nt = min(CH, t_hi - t0), slot = q % 3;
            float *ac = sact + slot * SM::ACT;
            uint64_t *bar = bars + 1 + slot;
            if (lane == 0) {
                const uint32_t gbytes = (uint32_t)(nt * GS * 4);
                if (t0 > 0) {
                    const uint32_t bytes = (uint32_t)((nt + 1) * ROW * 4);
                    mbar_expect_tx(bar, bytes + gbytes);
                    tma_load_1d(ac, svg + (size_t)(t0 - 1) * ROW, bytes, bar);
                } else {
                    const uint32_t bytes = (uint32_t)(nt * ROW * 4);
                    mbar_expect_tx(bar, bytes + gbytes);
                    tma_load_1d(ac + ROW, svg, bytes, bar);
                }
                tma_load_1d(sG + slot * SM::G, gbuf + (size_t)t0 * GS, gbytes, bar);
            }
            if (t0 == 0) { for (int i = lane; i < ROW; i += 32) ac[i] = 0.f; }   // h_{-1} = 0
        };
        // global loads of a block (x, and out/target or dLoss/dout) are issued one block ahead and held in registers, so that their
        // latency overlaps the previous block's work instead of heading every stage
        auto fetch = [&](int q, float2 &v, float2 &go) {
            const int t0 = t_lo + q * CH, nt = min(CH, t_hi - t0);
            v = make_float2(0.f, 0.f); go = make_float2(0.f, 0.f);
            if (q < nblk && lane < nt) { v = __ldg(x2 + t0 + lane); go = load_gout(go2, oi2, y2, t0 + lane, gs); }
        };
        const float wo0 = (HEAD && lane < H) ? sp[L.oWo + lane] : 0.f, wo1 = (HEAD && lane < H) ? sp[L.oWo + L.O + lane] : 0.f;
        issue(0);
        float2 v_n, go_n;
        fetch(0, v_n, go_n);
        for (int s = 0; s < nblk + 1; ++s) {
            if (s < nblk) {
                if (s + 1 < nblk) issue(s + 1);
                const float2 v = v_n, go = go_n;
                fetch(s + 1, v_n, go_n);
                const int t0 = t_lo + s * CH, nt = min(CH, t_hi - t0), slot = s % 3;
                float *pr = spre + (s & 1) * SM::PRE, *dp = sdp + (s & 1) * SM::DP;
                float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (lane < nt) {
                    features_fwd<FM>(v.x, v.y, 0.f, 0.f, f);
                    f[F] = 1.0f;                                                  // the ones column: bias gradients fall out of the same GEMM
                }
                // lanes >= nt (ragged last block) contribute zeros: the GEMM warps always run all 32 timesteps of a block
                float4 *d = reinterpret_cast<float4 *>(pr + lane * 12);
                d[0] = make_float4(f[0], f[1], f[2], f[3]);
                d[1] = make_float4(f[4], f[5], f[6], f[7]);
                d[2] = make_float4(go.x, go.y, 0.f, 0.f);
                __syncwarp();
                mbar_wait(bars + 1 + slot, (uint32_t)((s / 3) & 1));
                if (nt < CH) {      // ragged block: zero the G rows and the h rows of the missing timesteps so that they add nothing
                    float *Gz = sG + slot * SM::G + nt * GS;
                    for (int i = lane; i < (CH - nt) * GS; i += 32) Gz[i] = 0.f;
                    float *az = sact + slot * SM::ACT + (nt + 1) * ROW;
                    for (int i = lane; i < (CH - nt) * ROW; i += 32) az[i] = 0.f;
                    __syncwarp();
                }
                if constexpr (HEAD) {
                    // dpre with lane = unit (conflict-free): relu'(fc_hid) * (W_o^T dLoss/dout); zero rows for a ragged tail
                    if (lane < HP) {
                        const float *grow = sact + slot * SM::ACT + ROW + 5 * HP + lane;
#pragma unroll 4
                        for (int tl = 0; tl < CH; ++tl) {
                            const float2 g2 = *reinterpret_cast<const float2 *>(pr + tl * 12 + 8);
                            dp[tl * HP + lane] = (tl < nt && grow[tl * ROW] > 0.f) ? fmaf(wo0, g2.x, wo1 * g2.y) : 0.f;
                        }
                    }
                }
            }
            __syncthreads();
        }
    } else if (warp == 1) {
        // =============================== dL/dW_hh
        constexpr int TR = (3 * HP + 7) / 8, TC = HP / 4;
        const int rg = lane >> 2, cg = lane & 3;
        float acc[DW ? TR * TC : 1];
        if constexpr (DW) {
#pragma unroll
            for (int q = 0; q < TR * TC; ++q) acc[q] = 0.f;
        }
        for (int s = 0; s < nblk + 1; ++s) {
            const int sc = s - 1;
            if (sc >= 0) {
                if constexpr (DW) {
                    const float *Ap = sG + (sc % 3) * SM::G + TR * rg;
                    const float *Bp = sact + (sc % 3) * SM::ACT + 4 * HP + TC * cg;      // row tl = step t-1: h_{t-1}
#pragma unroll 4
                    for (int tl = 0; tl < CH; ++tl) {
                        float av[TR], bv[TC];
                        lds_vec<TR>(Ap + tl * GS, av);
                        lds_vec<TC>(Bp + tl * ROW, bv);
                        outer_acc<TR, TC>(acc, av, bv);
                    }
                }
            }
            __syncthreads();
        }
        if constexpr (DW) {
            if (prt) {
#pragma unroll
                for (int i = 0; i < TR; ++i) {
                    const int r = TR * rg + i, g = r / HP, k = r - g * HP;
#pragma unroll
                    for (int c = 0; c < TC; ++c) {
                        const int jj = TC * cg + c;
                        if (r < 3 * HP && k < H && jj < H) prt[L.oWhh + (g * H + k) * H + jj] = acc[i * TC + c];
                    }
                }
            }
        }
    } else if (warp == 2) {
        // =============================== dL/dW_ih + every bias (ones column), head weights
        constexpr int NC = (F + 1 + 3) & ~3;                 // feature columns + the ones column, padded to 4
        constexpr int LC = NC / 4, LR = 32 / LC, TR = (4 * HP + LR - 1) / LR, TC = 4;
        constexpr int TR3 = (HP + 7) / 8, TC3 = HP / 4;       // head: dpre (x) h_t
        const int rg = lane / LC, cg = lane % LC;
        const int rg3 = lane >> 2, cg3 = lane & 3;
        const int lp = lane < HP ? lane : 0;
        float acc[DW ? TR * TC : 1], acc3[(DW && HEAD) ? TR3 * TC3 : 1];
        float gwo0 = 0.f, gwo1 = 0.f;
        if constexpr (DW) {
#pragma unroll
            for (int q = 0; q < TR * TC; ++q) acc[q] = 0.f;
            if constexpr (HEAD) {
#pragma unroll
                for (int q = 0; q < TR3 * TC3; ++q) acc3[q] = 0.f;
            }
        }
        for (int s = 0; s < nblk + 1; ++s) {
            const int sc = s - 1;
            if (sc >= 0) {
                if constexpr (DW) {
                    const float *Ap = sG + (sc % 3) * SM::G + (TR * rg < 4 * HP ? TR * rg : 0);      // lanes whose rows lie past the tile (their sums are discarded below) re-read row group 0 instead of running off the row
                    const float *Bp = spre + (sc & 1) * SM::PRE + TC * cg;
                    const float *ac = sact + (sc % 3) * SM::ACT;
                    const float *pr = spre + (sc & 1) * SM::PRE;
                    const float *A3 = sdp + (sc & 1) * SM::DP + (TR3 * rg3 < HP ? TR3 * rg3 : 0);
                    const float *B3 = ac + ROW + 4 * HP + TC3 * cg3;                      // row tl+1 = step t: h_t
                    const float *hcol = ac + ROW + (HEAD ? 5 : 4) * HP + lp;             // g_t (DGRU) or h_t feeds fc_out
#pragma unroll 4
                    for (int tl = 0; tl < CH; ++tl) {
                        float av[TR], bv[TC];
                        lds_vec<TR>(Ap + tl * GS, av);
                        lds_vec<TC>(Bp + tl * 12, bv);
                        outer_acc<TR, TC>(acc, av, bv);
                        if constexpr (HEAD) {
                            float a3[TR3], b3[TC3];
                            lds_vec<TR3>(A3 + tl * HP, a3);
                            lds_vec<TC3>(B3 + tl * ROW, b3);
                            outer_acc<TR3, TC3>(acc3, a3, b3);
                        }
                        const float2 go = *reinterpret_cast<const float2 *>(pr + tl * 12 + 8);
                        const float hv = hcol[tl * ROW];
                        gwo0 = fmaf(go.x, hv, gwo0); gwo1 = fmaf(go.y, hv, gwo1);
                    }
                }
            }
            __syncthreads();
        }
        if constexpr (DW) {
            if (prt) {
#pragma unroll
                for (int i = 0; i < TR; ++i) {
                    const int r = TR * rg + i, blk = r / HP, k = r - blk * HP;     // blk: 0 ar, 1 az, 2 an*r, 3 an
                    if (r < 4 * HP && k < H) {
#pragma unroll
                        for (int c = 0; c < TC; ++c) {
                            const int col = TC * cg + c;
                            const float v = acc[i * TC + c];
                            if (col < F && blk != 2) prt[L.oWih + ((blk == 3 ? 2 : blk) * H + k) * F + col] = v;
                            if (col == F) {
                                if (blk < 2) { prt[L.obih + blk * H + k] = v; prt[L.obhh + blk * H + k] = v; }
                                else if (blk == 3) prt[L.obih + 2 * H + k] = v;
                                else prt[L.obhh + 2 * H + k] = v;
                            }
                        }
                    }
                }
                if constexpr (HEAD) {
#pragma unroll
                    for (int i = 0; i < TR3; ++i) {
                        const int k = TR3 * rg3 + i;
#pragma unroll
                        for (int c = 0; c < TC3; ++c) {
                            const int jj = TC3 * cg3 + c;
                            if (k < H && jj < H) prt[L.oWh + k * H + jj] = acc3[i * TC3 + c];
                        }
                    }
                }
                if (lane < H) { prt[L.oWo + lane] = gwo0; prt[L.oWo + L.O + lane] = gwo1; }
            }
        }
    } else {
        // =============================== dL/dfeatures -> dL/dx and the per-timestep sums: one timestep per lane
        float gbo0 = 0.f, gbo1 = 0.f;
        float gwof[(DW && HEAD) ? 2 * F : 1], gbh[(DW && HEAD) ? HP : 1];
        if constexpr (DW && HEAD) {
#pragma unroll
            for (int q = 0; q < 2 * F; ++q) gwof[q] = 0.f;
#pragma unroll
            for (int q = 0; q < HP; ++q) gbh[q] = 0.f;
        }
        float2 *gx2 = (a.need_dx && a.gx) ? reinterpret_cast<float2 *>(a.gx) + (size_t)b * T : nullptr;
        for (int s = 0; s < nblk + 1; ++s) {
            const int sc = s - 1;
            if (sc >= 0) {
                const int t0 = t_lo + sc * CH, nt = min(CH, t_hi - t0);
                const float *Gb = sG + (sc % 3) * SM::G;
                const float *p = spre + (sc & 1) * SM::PRE + lane * 12;
                const float g0 = p[8], g1 = p[9];          // zero for lanes >= nt
                if constexpr (DW) {
                    gbo0 += g0; gbo1 += g1;
                    if constexpr (HEAD) {
                        float fv[8];
                        lds_vec<8>(p, fv);
#pragma unroll
                        for (int q = 0; q < F; ++q) { gwof[q] = fmaf(g0, fv[q], gwof[q]); gwof[F + q] = fmaf(g1, fv[q], gwof[F + q]); }
                        float dv[HP];
                        lds_vec<HP>(sdp + (sc & 1) * SM::DP + lane * HP, dv);
#pragma unroll
                        for (int q = 0; q < HP; ++q) gbh[q] += dv[q];
                    }
                }
                if (gx2 && lane < nt) {
                    // dL/dfeat[q] = sum_k W_ih[r,k][q] ar_k + W_ih[z,k][q] az_k + W_ih[n,k][q] an_k  (+ head: wof[.][q] . go)
                    float gf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    float gv[4 * HP];          // ar | az | an*r | an of this lane's timestep (GS/4 odd: conflict-free LDS.128)
                    lds_vec<4 * HP>(Gb + lane * GS, gv);
#pragma unroll
                    for (int k = 0; k < HT; ++k) {
                        if (k < H) {
                            const float ar = gv[k], az = gv[HP + k], an = gv[3 * HP + k];
                            const float *wr = sp + L.oWih + (0 * H + k) * F, *wz = sp + L.oWih + (1 * H + k) * F, *wn = sp + L.oWih + (2 * H + k) * F;
#pragma unroll
                            for (int q = 0; q < F; ++q) gf[q] = fmaf(wn[q], an, fmaf(wz[q], az, fmaf(wr[q], ar, gf[q])));
                        }
                    }
                    if constexpr (HEAD) {
#pragma unroll
                        for (int q = 0; q < F; ++q) gf[q] += fmaf(sp[L.oWo + H + q], g0, sp[L.oWo + L.O + H + q] * g1);
                    }
                    float gi, gq;
                    features_bwd<FM>(p[0], p[1], gf, gi, gq);          // features 0,1 are the raw (I,Q) sample in every feature mode
                    gx2[t0 + lane] = make_float2(gi, gq);
                }
            }
            __syncthreads();
        }
        if constexpr (DW) {
            if (prt) {
                gbo0 = warp_sum(gbo0); gbo1 = warp_sum(gbo1);
                if (lane == 0) { prt[L.obo] = gbo0; prt[L.obo + 1] = gbo1; }
                if constexpr (HEAD) {
#pragma unroll
                    for (int q = 0; q < 2 * F; ++q) {
                        const float sum = warp_sum(gwof[q]);
                        if (lane == 0) prt[L.oWo + (q / F) * L.O + H + (q % F)] = sum;
                    }
#pragma unroll
                    for (int q = 0; q < HP; ++q) {
                        const float sum = warp_sum(gbh[q]);
                        if (lane == 0 && q < H) prt[L.obh + q] = sum;
                    }
                }
            }
        }
    }
}

// ================================================================ backward, lean fused form (default)
// Measured (B200, C2a): the split form above does not pay — its chain kernel runs 8 chunks at 54 us, but the weights kernel needs
// ~45 us to stream G and the saved rows back in (170 MB of extra traffic; 21 KB bulk copies with one block of look-ahead).  The lean
// fused form keeps everything in one kernel like the original, so G never leaves shared memory, but borrows the two things that
// made the split attractive: the register-tiled GEMM warps (every lane busy) instead of lane = hidden unit, and a footprint small
// enough (5 warps, <= 136 registers, ~75 KB) for 3 CTAs per SM => 6 time chunks instead of 4.
//   warp 0 chain | warp 1 pre (TMA rows, features, dLoss/dout, head back-projection) | warp 2 dW_hh | warp 3 dW_ih + biases + head
//   weights | warp 4 dL/dx + per-timestep sums.   Stage s: pre works on block s, chain on block s-1, the three GEMM warps on block s-2.
template <int HT, int FM, int HEAD>
struct BwdfSmem {
    static constexpr int HP = Pad4<HT>::value, ROW = Row<HT, HEAD>::value;
    static constexpr int ACT = (CH + 1) * ROW, DH = CH * HP, G = CH * GStride<HT>::value, PRE = CH * 12;
    __host__ __device__ static constexpr int total(int Ppad) { return 16 + Ppad + 3 * ACT + 2 * DH + 2 * G + 3 * PRE + 3 * DH; }
};

template <int HT, int FM, int HEAD, bool DW>
__global__ void __launch_bounds__(160, 1) gru_bwdf_kernel(GruArgs a) {
    pdl_enter();
    constexpr int F = FeatN<FM>::value, HP = Pad4<HT>::value, ROW = Row<HT, HEAD>::value, GS = GStride<HT>::value;
    using SM = BwdfSmem<HT, FM, HEAD>;
    const GruLayout<FM, HEAD> L(a.H);
    const int H = a.H, T = a.T;
    extern __shared__ __align__(128) float smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);   // [0] params, [1..3] activation slots
    float *sp = smem + 16;
    const int Ppad = (L.P + 3) & ~3;
    float *sact = sp + Ppad;                 // [3][(CH+1)][ROW]   row 0 = step t0-1
    float *sdh = sact + 3 * SM::ACT;         // [2][CH][HP]        dL/dh_t from the head
    float *sG = sdh + 2 * SM::DH;            // [2][CH][GS]        ar | az | an*r | an | pad
    float *spre = sG + 2 * SM::G;            // [3][CH][12]        feat(F) | 1 | 0.. (8) | go(2) | -
    float *sdp = spre + 3 * SM::PRE;         // [3][CH][HP]        dpre (DGRU head)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const BwdRange R = bwd_range(a);
    const bool spec = R.spec;
    const int b = R.b, t_elo = R.t_elo, t_ehi = R.t_ehi, t_hi = R.t_hi;
    if (a.mode == 2 && bwd_verify_pass(a, b, HP, HP, H)) return;

    if (threadIdx.x == 0) { mbar_init(bars + 1, 1); mbar_init(bars + 2, 1); mbar_init(bars + 3, 1); }
    stage_params(sp, a.params, L.P, bars);
    const bool act = lane < H;
    const int j = act ? lane : 0;
    const int lp = lane < HP ? lane : 0;
    const int cb = t_elo / CH, ce = (t_hi + CH - 1) / CH, nchunks = ce - cb, ce_emit = (t_ehi + CH - 1) / CH;
    const IqRow x2 = iq_row(a.x, a.x_bf16, a.x_starts, b, T);
    const float *svg = a.saved + (size_t)b * T * ROW;
    float *prt = nullptr;
    if constexpr (DW) {
        if (a.partials && warp >= 2) prt = chunk_partial_row(a, spec, b, L.P, (warp - 2) * 32 + lane, 96);
    }

    if (warp == 1) {
        // =============================== pre
        const float2 *go2 = a.gout ? reinterpret_cast<const float2 *>(a.gout) + (size_t)b * T : nullptr;
        const float2 *oi2 = a.out_in ? reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * T : nullptr;
        const IqRow y2 = iq_row(a.target, a.target_bf16, a.target_starts, b, T);
        const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);
        float whT[HEAD ? HT : 1];
        if constexpr (HEAD) {
#pragma unroll
            for (int k = 0; k < HT; ++k) whT[k] = (act && k < H) ? sp[L.oWh + k * H + j] : 0.f;
        }
        const float wo0 = act ? sp[L.oWo + j] : 0.f, wo1 = act ? sp[L.oWo + L.O + j] : 0.f;
        for (int s = 0; s < nchunks + 2; ++s) {
            if (s < nchunks) {
                const int c = ce - 1 - s, t0 = c * CH, nt = min(CH, t_hi - t0), slot = s % 3;
                float *ac = sact + slot * SM::ACT;
                float *pr = spre + slot * SM::PRE, *dp = sdp + slot * SM::DH;
                float *dh = sdh + (s & 1) * SM::DH;
                uint64_t *bar = bars + 1 + slot;
                load_rows_with_prev(ac, svg, ROW, t0, nt, lane, bar);
                {   // one timestep per lane: features (+ the ones column) and dLoss/dout; zeros for a ragged tail
                    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    float2 go = make_float2(0.f, 0.f);
                    if (lane < nt) {
                        const float2 v = __ldg(x2 + t0 + lane);
                        features_fwd<FM>(v.x, v.y, 0.f, 0.f, f);
                        f[F] = 1.0f;
                        go = load_gout(go2, oi2, y2, t0 + lane, gs);
                    }
                    float4 *d = reinterpret_cast<float4 *>(pr + lane * 12);
                    d[0] = make_float4(f[0], f[1], f[2], f[3]);
                    d[1] = make_float4(f[4], f[5], f[6], f[7]);
                    d[2] = make_float4(go.x, go.y, 0.f, 0.f);
                }
                __syncwarp();
                mbar_wait(bar, (uint32_t)((s / 3) & 1));
                if (nt < CH) {      // ragged block: the GEMM warps run all 32 timesteps, the missing ones must add nothing
                    float *az = ac + (nt + 1) * ROW;
                    for (int i = lane; i < (CH - nt) * ROW; i += 32) az[i] = 0.f;
                    __syncwarp();
                }
                if constexpr (HEAD) {
                    if (lane < HP) {
                        const float *grow = ac + ROW + 5 * HP + lane;
#pragma unroll 4
                        for (int tl = 0; tl < CH; ++tl) {
                            const float2 g2 = *reinterpret_cast<const float2 *>(pr + tl * 12 + 8);
                            dp[tl * HP + lane] = (tl < nt && grow[tl * ROW] > 0.f) ? fmaf(wo0, g2.x, wo1 * g2.y) : 0.f;
                        }
                    }
                    __syncwarp();
                    if (lane < HP) {
#pragma unroll 2
                        for (int tl = 0; tl < nt; ++tl) {
                            float d0 = 0.f, d1 = 0.f;
                            const float4 *dp4 = reinterpret_cast<const float4 *>(dp + tl * HP);
#pragma unroll
                            for (int k4 = 0; k4 < HP / 4; ++k4) {
                                const float4 dv = dp4[k4];
                                const float dk[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int k = k4 * 4 + e;
                                    if (k < HT) { if (k & 1) d1 = fmaf(whT[k], dk[e], d1); else d0 = fmaf(whT[k], dk[e], d0); }
                                }
                            }
                            dh[tl * HP + lane] = d0 + d1;
                        }
                    }
                } else {
                    if (lane < HP) {
                        for (int tl = 0; tl < nt; ++tl) {
                            const float2 g2 = *reinterpret_cast<const float2 *>(pr + tl * 12 + 8);
                            dh[tl * HP + lane] = fmaf(wo0, g2.x, wo1 * g2.y);
                        }
                    }
                }
            }
            __syncthreads();
        }
    } else if (warp == 0) {
        // =============================== chain: reverse-time recurrence of dL/dh
        float wcol[3 * HT];   // column j of W_hh
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int k = 0; k < HT; ++k) wcol[g * HT + k] = (act && k < H) ? sp[L.oWhh + (g * H + k) * H + j] : 0.f;
        float gH = 0.f;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 1;
            if (sc >= 0 && sc < nchunks) {
                const int c = ce - 1 - sc, t0 = c * CH, nt = min(CH, t_hi - t0);
                if (spec && c == ce_emit - 1 && t_ehi < T && lane < HP) a.sc_guess[(size_t)blockIdx.x * HP + lane] = gH;
                const float *ac = sact + (sc % 3) * SM::ACT + lp;
                const float *dh = sdh + (sc & 1) * SM::DH + lp;
                float *Gb = sG + (sc & 1) * SM::G;
                if (nt < CH) { for (int i = lane; i < (CH - nt) * GS; i += 32) Gb[nt * GS + i] = 0.f; }
                const float *row = ac + nt * ROW;   // step t0+nt-1
                float r = row[0], z = row[HP], n = row[2 * HP], hgn = row[3 * HP], hp = row[4 * HP - ROW], dht = dh[(nt - 1) * HP];
                for (int tl = nt - 1; tl >= 0; --tl) {
                    const int tp = tl > 0 ? tl - 1 : 0;
                    const float *rown = ac + (tp + 1) * ROW;
                    const float r_n = rown[0], z_n = rown[HP], n_n = rown[2 * HP], hgn_n = rown[3 * HP], hp_n = rown[4 * HP - ROW],
                                dh_n = dh[tp * HP];
                    gH += dht;
                    const float gz = gH * (hp - n), gn = gH * (1.f - z), ghp = gH * z;
                    const float an = gn * (1.f - n * n);
                    const float az = gz * z * (1.f - z);
                    const float anr = an * r;
                    const float ar = anr * hgn * (1.f - r);
                    float *G = Gb + tl * GS;
                    if (lane < HP) { G[lane] = ar; G[HP + lane] = az; G[2 * HP + lane] = anr; G[3 * HP + lane] = an; }
                    __syncwarp();
                    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f, acc4 = 0.f, acc5 = 0.f;
                    const float4 *G4 = reinterpret_cast<const float4 *>(G);
                    float4 gv[3 * (HP / 4)];
#pragma unroll
                    for (int q = 0; q < 3 * (HP / 4); ++q) gv[q] = G4[q];
#pragma unroll
                    for (int k4 = 0; k4 < HP / 4; ++k4) {
                        const float kr[4] = {gv[k4].x, gv[k4].y, gv[k4].z, gv[k4].w};
                        const float kz[4] = {gv[HP / 4 + k4].x, gv[HP / 4 + k4].y, gv[HP / 4 + k4].z, gv[HP / 4 + k4].w};
                        const float kn[4] = {gv[2 * (HP / 4) + k4].x, gv[2 * (HP / 4) + k4].y, gv[2 * (HP / 4) + k4].z, gv[2 * (HP / 4) + k4].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int k = k4 * 4 + e;
                            if (k < HT) {
                                if (k & 1) { acc1 = fmaf(wcol[k], kr[e], acc1); acc3 = fmaf(wcol[HT + k], kz[e], acc3); acc5 = fmaf(wcol[2 * HT + k], kn[e], acc5); }
                                else       { acc0 = fmaf(wcol[k], kr[e], acc0); acc2 = fmaf(wcol[HT + k], kz[e], acc2); acc4 = fmaf(wcol[2 * HT + k], kn[e], acc4); }
                            }
                        }
                    }
                    gH = (ghp + (acc0 + acc1)) + ((acc2 + acc3) + (acc4 + acc5));
                    r = r_n; z = z_n; n = n_n; hgn = hgn_n; hp = hp_n; dht = dh_n;
                }
            }
            __syncthreads();
        }
        if (spec && lane < HP) a.sc_end[(size_t)blockIdx.x * HP + lane] = gH;
    } else if (warp == 2) {
        // =============================== dL/dW_hh = sum_t [ar|az|an*r]_t (x) h_{t-1}: lanes 8 x 4, register tile TR x TC
        constexpr int TR = (3 * HP + 7) / 8, TC = HP / 4;
        const int rg = lane >> 2, cg = lane & 3;
        float acc[DW ? TR * TC : 1];
        if constexpr (DW) {
#pragma unroll
            for (int q = 0; q < TR * TC; ++q) acc[q] = 0.f;
        }
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 2;
            if (sc >= 0 && ce - 1 - sc < ce_emit) {
                if constexpr (DW) {
                    const float *Ap = sG + (sc & 1) * SM::G + TR * rg;
                    const float *Bp = sact + (sc % 3) * SM::ACT + 4 * HP + TC * cg;      // row tl = step t-1: h_{t-1}
#pragma unroll 4
                    for (int tl = 0; tl < CH; ++tl) {
                        float av[TR], bv[TC];
                        lds_vec<TR>(Ap + tl * GS, av);
                        lds_vec<TC>(Bp + tl * ROW, bv);
                        outer_acc<TR, TC>(acc, av, bv);
                    }
                }
            }
            __syncthreads();
        }
        if constexpr (DW) {
            if (prt) {
#pragma unroll
                for (int i = 0; i < TR; ++i) {
                    const int r = TR * rg + i, g = r / HP, k = r - g * HP;
#pragma unroll
                    for (int c = 0; c < TC; ++c) {
                        const int jj = TC * cg + c;
                        if (r < 3 * HP && k < H && jj < H) prt[L.oWhh + (g * H + k) * H + jj] = acc[i * TC + c];
                    }
                }
            }
        }
    } else if (warp == 3) {
        // =============================== dL/dW_ih + every bias (ones column) = sum_t [ar|az|an*r|an]_t (x) [feat_t, 1]; head weights
        constexpr int NC = (F + 1 + 3) & ~3;
        constexpr int LC = NC / 4, LR = 32 / LC, TR = (4 * HP + LR - 1) / LR, TC = 4;
        constexpr int TR3 = (HP + 7) / 8, TC3 = HP / 4;
        static_assert((4 * HP) % TR == 0 || TR <= 4, "a lane's rows must end inside the padded G row (GS = 4*HP + 4)");
        static_assert(HP % TR3 == 0, "a lane's rows of the head tile must end inside the row");
        const int rg = lane / LC, cg = lane % LC;
        const int rg3 = lane >> 2, cg3 = lane & 3;
        float acc[DW ? TR * TC : 1], acc3[(DW && HEAD) ? TR3 * TC3 : 1];
        float gwo0 = 0.f, gwo1 = 0.f;
        if constexpr (DW) {
#pragma unroll
            for (int q = 0; q < TR * TC; ++q) acc[q] = 0.f;
            if constexpr (HEAD) {
#pragma unroll
                for (int q = 0; q < TR3 * TC3; ++q) acc3[q] = 0.f;
            }
        }
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 2;
            if (sc >= 0 && ce - 1 - sc < ce_emit) {
                if constexpr (DW) {
                    const float *Ap = sG + (sc & 1) * SM::G + (TR * rg < 4 * HP ? TR * rg : 0);      // lanes whose rows lie past the tile (their sums are discarded below) re-read row group 0 instead of running off the row
                    const float *pr = spre + (sc % 3) * SM::PRE;
                    const float *Bp = pr + TC * cg;
                    const float *ac = sact + (sc % 3) * SM::ACT;
                    const float *A3 = sdp + (sc % 3) * SM::DH + (TR3 * rg3 < HP ? TR3 * rg3 : 0);
                    const float *B3 = ac + ROW + 4 * HP + TC3 * cg3;                      // row tl+1 = step t: h_t
                    const float *hcol = ac + ROW + (HEAD ? 5 : 4) * HP + lp;             // g_t (DGRU) or h_t feeds fc_out
#pragma unroll 4
                    for (int tl = 0; tl < CH; ++tl) {
                        float av[TR], bv[TC];
                        lds_vec<TR>(Ap + tl * GS, av);
                        lds_vec<TC>(Bp + tl * 12, bv);
                        outer_acc<TR, TC>(acc, av, bv);
                        if constexpr (HEAD) {
                            float a3[TR3], b3[TC3];
                            lds_vec<TR3>(A3 + tl * HP, a3);
                            lds_vec<TC3>(B3 + tl * ROW, b3);
                            outer_acc<TR3, TC3>(acc3, a3, b3);
                        }
                        const float2 go = *reinterpret_cast<const float2 *>(pr + tl * 12 + 8);
                        const float hv = hcol[tl * ROW];
                        gwo0 = fmaf(go.x, hv, gwo0); gwo1 = fmaf(go.y, hv, gwo1);
                    }
                }
            }
            __syncthreads();
        }
        if constexpr (DW) {
            if (prt) {
#pragma unroll
                for (int i = 0; i < TR; ++i) {
                    const int r = TR * rg + i, blk = r / HP, k = r - blk * HP;     // blk: 0 ar, 1 az, 2 an*r, 3 an
                    if (r < 4 * HP && k < H) {
#pragma unroll
                        for (int c = 0; c < TC; ++c) {
                            const int col = TC * cg + c;
                            const float v = acc[i * TC + c];
                            if (col < F && blk != 2) prt[L.oWih + ((blk == 3 ? 2 : blk) * H + k) * F + col] = v;
                            if (col == F) {
                                if (blk < 2) { prt[L.obih + blk * H + k] = v; prt[L.obhh + blk * H + k] = v; }
                                else if (blk == 3) prt[L.obih + 2 * H + k] = v;
                                else prt[L.obhh + 2 * H + k] = v;
                            }
                        }
                    }
                }
                if constexpr (HEAD) {
#pragma unroll
                    for (int i = 0; i < TR3; ++i) {
                        const int k = TR3 * rg3 + i;
#pragma unroll
                        for (int c = 0; c < TC3; ++c) {
                            const int jj = TC3 * cg3 + c;
                            if (k < H && jj < H) prt[L.oWh + k * H + jj] = acc3[i * TC3 + c];
                        }
                    }
                }
                if (lane < H) { prt[L.oWo + lane] = gwo0; prt[L.oWo + L.O + lane] = gwo1; }
            }
        }
    } else {
        // =============================== dL/dfeatures -> dL/dx and the per-timestep sums: one timestep per lane
        float gbo0 = 0.f, gbo1 = 0.f;
        float gwof[(DW && HEAD) ? 2 * F : 1], gbh[(DW && HEAD) ? HP : 1];
        if constexpr (DW && HEAD) {
#pragma unroll
            for (int q = 0; q < 2 * F; ++q) gwof[q] = 0.f;
#pragma unroll
            for (int q = 0; q < HP; ++q) gbh[q] = 0.f;
        }
        float2 *gx2 = (a.need_dx && a.gx) ? reinterpret_cast<float2 *>(a.gx) + (size_t)b * T : nullptr;
        for (int s = 0; s < nchunks + 2; ++s) {
            const int sc = s - 2;
            if (sc >= 0 && ce - 1 - sc < ce_emit) {
                const int c = ce - 1 - sc, t0 = c * CH, nt = min(CH, t_hi - t0);
                const float *Gb = sG + (sc & 1) * SM::G;
                const float *p = spre + (sc % 3) * SM::PRE + lane * 12;
                const float g0 = p[8], g1 = p[9];          // zero for lanes >= nt
                if constexpr (DW) {
                    gbo0 += g0; gbo1 += g1;
                    if constexpr (HEAD) {
                        float fv[8];
                        lds_vec<8>(p, fv);
#pragma unroll
                        for (int q = 0; q < F; ++q) { gwof[q] = fmaf(g0, fv[q], gwof[q]); gwof[F + q] = fmaf(g1, fv[q], gwof[F + q]); }
                        float dv[HP];
                        lds_vec<HP>(sdp + (sc % 3) * SM::DH + lane * HP, dv);
#pragma unroll
                        for (int q = 0; q < HP; ++q) gbh[q] += dv[q];
                    }
                }
                if (gx2 && lane < nt) {
                    float gf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    float gv[4 * HP];          // ar | az | an*r | an of this lane's timestep (GS/4 odd: conflict-free LDS.128)
                    lds_vec<4 * HP>(Gb + lane * GS, gv);
#pragma unroll
                    for (int k = 0; k < HT; ++k) {
                        if (k < H) {
                            const float ar = gv[k], az = gv[HP + k], an = gv[3 * HP + k];
                            const float *wr = sp + L.oWih + (0 * H + k) * F, *wz = sp + L.oWih + (1 * H + k) * F, *wn = sp + L.oWih + (2 * H + k) * F;
#pragma unroll
                            for (int q = 0; q < F; ++q) gf[q] = fmaf(wn[q], an, fmaf(wz[q], az, fmaf(wr[q], ar, gf[q])));
                        }
                    }
                    if constexpr (HEAD) {
#pragma unroll
                        for (int q = 0; q < F; ++q) gf[q] += fmaf(sp[L.oWo + H + q], g0, sp[L.oWo + L.O + H + q] * g1);
                    }
                    float gi, gq;
                    features_bwd<FM>(p[0], p[1], gf, gi, gq);          // features 0,1 are the raw (I,Q) sample in every feature mode
                    gx2[t0 + lane] = make_float2(gi, gq);
                }
            }
            __syncthreads();
        }
        if constexpr (DW) {
            if (prt) {
                gbo0 = warp_sum(gbo0); gbo1 = warp_sum(gbo1);
                if (lane == 0) { prt[L.obo] = gbo0; prt[L.obo + 1] = gbo1; }
                if constexpr (HEAD) {
#pragma unroll
                    for (int q = 0; q < 2 * F; ++q) {
                        const float sum = warp_sum(gwof[q]);
                        if (lane == 0) prt[L.oWo + (q / F) * L.O + H + (q % F)] = sum;
                    }
#pragma unroll
                    for (int q = 0; q < HP; ++q) {
                        const float sum = warp_sum(gbh[q]);
                        if (lane == 0 && q < H) prt[L.obh + q] = sum;
                    }
                }
            }
        }
    }
}

// ================================================================ host dispatch
template <int FM, int HEAD> static int64_t gru_nparams(int H) { return GruLayout<FM, HEAD>(H).P; }

// compiled hidden-size tiers: exact for the sizes the reference's scripts use, next-larger tier otherwise
#define ODPD_GRU_TIERS(X, FM, HEAD) X(8, FM, HEAD) X(10, FM, HEAD) X(13, FM, HEAD) X(16, FM, HEAD) X(23, FM, HEAD) X(32, FM, HEAD)
static int gru_tier(int H) {
#define X(HTV, FMV, HEADV) if (H <= HTV) return HTV;
    ODPD_GRU_TIERS(X, 0, 0)
#undef X
    return -1;
}

// dir: 0 fwd, 1 bwd, 2 fwd plan only, 3 bwd plan only (plan -> info[0..3] = chunks, steps per chunk, warm-up, index of the re-run counter)
template <int HT, int FM, int HEAD>
static int launch_fwd(GruArgs a, cudaStream_t st, bool plan_only, int *info) {
    constexpr int HP = Pad4<HT>::value, ROW = Row<HT, HEAD>::value;
    const GruLayout<FM, HEAD> L(a.H);
    const size_t smem = (size_t)FwdSmem<HT, FM, HEAD>::total((L.P + 3) & ~3) * sizeof(float);
    static OccCache occ{};
    const int64_t soff = a.save ? (int64_t)a.B * a.T * ROW : 0;     // `saved` = [B][T][ROW] rows (when saving) | chunk scratch
    return chunk_launch(gru_fwd_kernel<HT, FM, HEAD>, 96, smem, &occ, a, 0, a.saved ? a.saved + soff : nullptr, soff, HP, st, plan_only, info,
                        "gru_fwd_kernel");
}
// backward workspace (floats):  [rowsP][P] gradient partials (4-aligned) | chunk scratch | G buffer [B][T][GS]
//   rowsP = max(rows of the fused kernel (one per (sequence, chunk)), rows of the weights kernel (one per (sequence, WT-step tile)))
static constexpr int WTILES_MAX = 16;   // weights-kernel tiles per sequence (upper bound: sizes the gradient-partial rows)
static inline int64_t bwd_partial_rows(int B, int T, int tchunks_req) {
    const int64_t nblk = (T + CH - 1) / CH;
    const int64_t a = chunk_rows(B, tchunks_req), b = (int64_t)(B > 0 ? B : 1) * (nblk < WTILES_MAX ? (nblk > 0 ? nblk : 1) : WTILES_MAX);
    return a > b ? a : b;
}
// backward form: 2 = lean fused (default), 1 = split (chain kernel + weights kernel), 0 = original fused.  env ODPD_BWD_FORM
static inline int bwd_form() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("ODPD_BWD_FORM"); v = e ? atoi(e) : 2; if (v < 0 || v > 2) v = 2; }
    return v;
}

template <int HT, int FM, int HEAD, bool DW>
static int launch_bwd_t(GruArgs a, cudaStream_t st, bool plan_only, int *info) {
    constexpr int HP = Pad4<HT>::value, GS = GStride<HT>::value;
    const GruLayout<FM, HEAD> L(a.H);
    const int Ppad = (L.P + 3) & ~3;
    const int64_t rows = chunk_rows(a.B, a.tchunks_req);
    const int64_t woff = (bwd_partial_rows(a.B, a.T, a.tchunks_req) * L.P + 3) & ~(int64_t)3;
    float *scr = a.partials ? a.partials + woff : nullptr;
    if (bwd_form() == 2) {
        const size_t smem = (size_t)BwdfSmem<HT, FM, HEAD>::total(Ppad) * sizeof(float);
        static OccCache occ{};
        const int rc = chunk_launch(gru_bwdf_kernel<HT, FM, HEAD, DW>, 160, smem, &occ, a, 1, scr, woff, HP, st, plan_only, info, "gru_bwdf_kernel");
        if (info) info[4] = a.B * info[0];
        return rc;
    }
    if (bwd_form() == 0 || (!plan_only && !a.partials)) {
        // fused form: one kernel (kept for A/B comparison, ODPD_BWD_SPLIT=0, and for a dX-only backward without workspace)
        const size_t smem = (size_t)BwdSmem<HT, HEAD>::total(Ppad) * sizeof(float);
        static OccCache occ{};
        const int rc = chunk_launch(gru_bwd_kernel<HT, FM, HEAD, DW>, 128, smem, &occ, a, 1, scr, woff, HP, st, plan_only, info, "gru_bwd_kernel");
        if (info) info[4] = a.B * info[0];
        return rc;
    }
    a.gbuf = scr ? scr + chunk_bwd_scratch_floats(rows, HP) : nullptr;
    const size_t smem_c = (size_t)BwdcSmem<HT, HEAD>::total(Ppad) * sizeof(float);
    static OccCache occ_c{};
    int rc = chunk_launch(gru_bwdc_kernel<HT, FM, HEAD>, 96, smem_c, &occ_c, a, 1, scr, woff, HP, st, plan_only, info, "gru_bwdc_kernel");
    // weights kernel: tiles per sequence chosen so that all CTAs are resident at once (one wave), at most WTILES_MAX
    const size_t smem_w = (size_t)BwdwSmem<HT, FM, HEAD>::total(Ppad) * sizeof(float);
    static OccCache occ_w{};
    auto kw = gru_bwdw_kernel<HT, FM, HEAD, DW>;
    int *ow = &occ_w.v[cur_dev_slot()];
    if (!*ow || smem_w > occ_w.smem[cur_dev_slot()]) {
        occ_w.smem[cur_dev_slot()] = smem_w;
        cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kw, 128, smem_w) != cudaSuccess || o <= 0) o = 1;
        *ow = o;
    }
    const int nblk = (a.T + CH - 1) / CH;
    int tiles = (*ow * num_sms()) / (a.B > 0 ? a.B : 1);
    tiles = tiles < 1 ? 1 : (tiles > WTILES_MAX ? WTILES_MAX : tiles);
    if (tiles > nblk) tiles = nblk;
    a.wt_blocks = (nblk + tiles - 1) / tiles;
    tiles = (nblk + a.wt_blocks - 1) / a.wt_blocks;
    if (info) info[4] = a.B * tiles;
    if (rc || plan_only) return rc;
    launch_pdl(kw, dim3((unsigned)tiles, (unsigned)a.B), dim3(128), smem_w, st, a);
    (void)GS;
    return check_launch("gru_bwdw_kernel");
}

template <int FM, int HEAD>
static int dispatch(const GruArgs &a, int dir, bool dw, cudaStream_t st, int *info) {
    const bool plan_only = dir >= 2;
    const int d = dir & 1;
#define X(HTV, FMV, HEADV)                                                                         \
    if (a.H <= HTV)                                                                                \
        return d == 0 ? launch_fwd<HTV, FMV, HEADV>(a, st, plan_only, info)                        \
                      : (dw ? launch_bwd_t<HTV, FMV, HEADV, true>(a, st, plan_only, info)          \
                            : launch_bwd_t<HTV, FMV, HEADV, false>(a, st, plan_only, info));
    ODPD_GRU_TIERS(X, FM, HEAD)
#undef X
    set_error("GRU-family kernels support hidden_size <= 32 (got %d)", a.H);
    return -1;
}

int64_t gru_family_nparams(int cell, int H) {
    switch (cell) {
    case ODPD_CELL_GRU: return gru_nparams<FM_RAW2, 0>(H);
    case ODPD_CELL_DGRU: return gru_nparams<FM_DGRU6, 1>(H);
    default: return gru_nparams<FM_QGRU4, 0>(H);
    }
}
// `saved` = [B][T][ROW] activation rows (only when the forward saves) followed by the chunk scratch
int64_t gru_family_saved_floats(int cell, int B, int T, int H, bool save, int tchunks_req) {
    const int ht = gru_tier(H);
    if (ht < 0) return -1;
    const int HP = (ht + 3) & ~3;
    const int64_t rowsz = save ? (int64_t)B * T * (cell == ODPD_CELL_DGRU ? 6 : 5) * HP : 0;
    return rowsz + chunk_fwd_scratch_floats(chunk_rows(B, tchunks_req), HP);
}
int64_t gru_family_workspace_floats(int cell, int B, int T, int H, int tchunks_req) {
    const int ht = gru_tier(H);
    if (ht < 0) return -1;
    const int HP = (ht + 3) & ~3;
    const int64_t partials = (bwd_partial_rows(B, T, tchunks_req) * gru_family_nparams(cell, H) + 3) & ~(int64_t)3;
    return partials + chunk_bwd_scratch_floats(chunk_rows(B, tchunks_req), HP) + (int64_t)(B > 0 ? B : 1) * (T > 0 ? T : 1) * (4 * HP + 4);
}

static int run_or_plan(int cell, const GruArgs &a, int dir, bool dw, cudaStream_t st, int *info) {
    switch (cell) {
    case ODPD_CELL_GRU: return dispatch<FM_RAW2, 0>(a, dir, dw, st, info);
    case ODPD_CELL_DGRU: return dispatch<FM_DGRU6, 1>(a, dir, dw, st, info);
    case ODPD_CELL_QGRU: return dispatch<FM_QGRU4, 0>(a, dir, dw, st, info);
    case ODPD_CELL_QGRU_AMP1: return dispatch<FM_AMP4, 0>(a, dir, dw, st, info);
    }
    set_error("gru_family_run: bad cell %d", cell);
    return -1;
}

// rows_out: number of gradient-partial rows the backward wrote (B * chunks), for the ordered reduction that follows
int gru_family_run(int cell, const GruArgs &a, int dir, bool dw, cudaStream_t st, int *rows_out) {
    int info[5] = {1, 0, 0, 0, 0};
    const int rc = run_or_plan(cell, a, dir, dw, st, info);
    if (rows_out) *rows_out = info[4] > 0 ? info[4] : a.B * info[0];
    return rc;
}

int gru_family_plan(int cell, int B, int T, int H, int tchunks_req, int twarm_req, int dir, bool dw, bool save, int out[4]) {
    GruArgs a{};
    a.B = B; a.T = T; a.H = H; a.tchunks_req = tchunks_req; a.twarm_req = twarm_req; a.save = save;
    int info[5] = {1, 0, 0, -1, 0};
    const int rc = run_or_plan(cell, a, 2 + (dir & 1), dw, nullptr, info);
    for (int i = 0; i < 4; ++i) out[i] = info[i];
    return rc;
}

}  // namespace odpd

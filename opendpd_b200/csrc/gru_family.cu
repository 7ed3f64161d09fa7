// gru_family.cu — fused forward / backward for the nn.GRU-based backbones: GRU, DGRU, QGRU, QGRU_AMP1.
//
// Replaces (reference, file:line): backbones/gru.py:45-48, backbones/dgru.py:59-74, backbones/qgru.py:59-71,
// backbones/qgru_amp1.py:59-76 and the ATen GRU cell behind torch.nn.GRU (gate order r,z,n;
// n = tanh(W_in x + b_in + r*(W_hn h + b_hn)); h' = (h-n)*z + n), plus nn.MSELoss (project.py:262-272).
//
// Mapping (B200): one warp per sequence, lane j owns hidden unit j.  The flat parameter block is staged once
// into shared memory by a TMA bulk copy, each lane then keeps its weight rows in registers for the whole frame;
// h_t lives in a register, is broadcast through a double-buffered shared-memory line (1 STS + H/4 LDS.128) and
// the recurrent matvec is 3H register FMAs per lane.  Work that is not part of the serial chain (feature
// extraction with IEEE sqrt/div, the 2-wide output reduction, the MSE, dX) is done time-parallel, 32 steps
// at a time, one step per lane, with coalesced float2 loads/stores of the IQ stream.
#include "cells.h"

namespace odpd {


template <int FM, int HEAD>
struct GruLayout {
    static constexpr int F = FeatN<FM>::value;
    int H, O, oWih, oWhh, obih, obhh, oWo, obo, oWh, obh, P, NS;
    __host__ __device__ explicit GruLayout(int h) {
        H = h; O = HEAD ? h + F : h;
        oWih = 0; oWhh = 3 * h * F; obih = oWhh + 3 * h * h; obhh = obih + 3 * h; oWo = obhh + 3 * h;
        obo = oWo + 2 * O; oWh = obo + 2; obh = oWh + h * h; P = HEAD ? obh + h : obo + 2;
        NS = HEAD ? 6 : 5;
    }
};

template <int HT> struct Pad4 { static constexpr int value = (HT + 3) & ~3; };

// shared memory: [mbarrier 16 B][params, padded to 4 floats][per warp scratch]
template <int HT> __host__ __device__ constexpr int fwd_warp_floats() { return 32 * 8 + 2 * Pad4<HT>::value + 2 * 32 * 33; }
template <int HT> __host__ __device__ constexpr int bwd_warp_floats() { return 32 * 8 + 32 * 2 + 32 * 8 + 2 * 4 * Pad4<HT>::value + 2 * Pad4<HT>::value; }

// ================================================================ forward
template <int HT, int FM, int HEAD>
__global__ void __launch_bounds__(128) gru_fwd_kernel(GruArgs a) {
    constexpr int F = FeatN<FM>::value, HP = Pad4<HT>::value;
    const GruLayout<FM, HEAD> L(a.H);
    const int H = a.H;
    extern __shared__ __align__(16) float smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 4;
    const int Ppad = (L.P + 3) & ~3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    float *ws = sp + Ppad + warp * fwd_warp_floats<HT>();
    float *sfeat = ws;                  // [32][8]
    float *shb = sfeat + 32 * 8;        // [2][HP]
    float *spo0 = shb + 2 * HP;         // [32][33]
    float *spo1 = spo0 + 32 * 33;       // [32][33]

    stage_params(sp, a.params, L.P, bar);
    const int b = blockIdx.x * wpc + warp;
    if (b >= a.B) return;

    const bool act = lane < H;
    const int j = act ? lane : 0;
    float whr[HT], whz[HT], whn[HT], wir[F], wiz[F], win[F], wh[HEAD ? HT : 1];
#pragma unroll
    for (int k = 0; k < HT; ++k) {
        const bool ok = act && k < H;
        whr[k] = ok ? sp[L.oWhh + (0 * H + j) * H + k] : 0.f;
        whz[k] = ok ? sp[L.oWhh + (1 * H + j) * H + k] : 0.f;
        whn[k] = ok ? sp[L.oWhh + (2 * H + j) * H + k] : 0.f;
        if constexpr (HEAD) wh[k] = ok ? sp[L.oWh + j * H + k] : 0.f;
    }
#pragma unroll
    for (int f = 0; f < F; ++f) {
        wir[f] = act ? sp[L.oWih + (0 * H + j) * F + f] : 0.f;
        wiz[f] = act ? sp[L.oWih + (1 * H + j) * F + f] : 0.f;
        win[f] = act ? sp[L.oWih + (2 * H + j) * F + f] : 0.f;
    }
    const float b_r = act ? sp[L.obih + j] + sp[L.obhh + j] : 0.f;
    const float b_z = act ? sp[L.obih + H + j] + sp[L.obhh + H + j] : 0.f;
    const float b_in = act ? sp[L.obih + 2 * H + j] : 0.f;
    const float b_hn = act ? sp[L.obhh + 2 * H + j] : 0.f;
    const float wo0 = act ? sp[L.oWo + j] : 0.f, wo1 = act ? sp[L.oWo + L.O + j] : 0.f;
    const float bh = (HEAD && act) ? sp[L.obh + j] : 0.f;
    const float bo0 = sp[L.obo], bo1 = sp[L.obo + 1];

    const float2 *x2 = reinterpret_cast<const float2 *>(a.x) + (size_t)b * a.T;
    const float2 *y2 = a.target ? reinterpret_cast<const float2 *>(a.target) + (size_t)b * a.T : nullptr;
    float2 *o2 = reinterpret_cast<float2 *>(a.out) + (size_t)b * a.T;
    float *sv = a.save ? a.saved + (size_t)b * a.T * L.NS * H : nullptr;

    float h = 0.f, lsum = 0.f;
    int cur = 0;
    if (lane < HP) { shb[lane] = 0.f; shb[HP + lane] = 0.f; }
    __syncwarp();

    for (int t0 = 0; t0 < a.T; t0 += ODPD_CHUNK) {
        const int nt = min(ODPD_CHUNK, a.T - t0);
        // ---- phase A: one timestep per lane — coalesced IQ load + feature extraction
        if (lane < nt) {
            const float2 v = __ldg(x2 + t0 + lane);
            float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            features_fwd<FM>(v.x, v.y, 0.f, 0.f, f);
            float4 *d = reinterpret_cast<float4 *>(sfeat + lane * 8);
            d[0] = make_float4(f[0], f[1], f[2], f[3]);
            d[1] = make_float4(f[4], f[5], f[6], f[7]);
        }
        __syncwarp();
        // ---- phase B: the serial recurrence
        for (int tl = 0; tl < nt; ++tl) {
            float feat[8];
            {
                const float4 *fp = reinterpret_cast<const float4 *>(sfeat + tl * 8);
                const float4 f0 = fp[0];
                feat[0] = f0.x; feat[1] = f0.y; feat[2] = f0.z; feat[3] = f0.w;
                if (F > 4) { const float4 f1 = fp[1]; feat[4] = f1.x; feat[5] = f1.y; feat[6] = f1.z; feat[7] = f1.w; }
            }
            float xr = b_r, xz = b_z, xn = b_in;
#pragma unroll
            for (int f = 0; f < F; ++f) { xr = fmaf(wir[f], feat[f], xr); xz = fmaf(wiz[f], feat[f], xz); xn = fmaf(win[f], feat[f], xn); }
            float ar0 = xr, ar1 = 0.f, az0 = xz, az1 = 0.f, an0 = b_hn, an1 = 0.f;
            const float4 *hb4 = reinterpret_cast<const float4 *>(shb + cur * HP);
#pragma unroll
            for (int k4 = 0; k4 < HP / 4; ++k4) {
                const float4 hv = hb4[k4];
                const float hk[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = k4 * 4 + e;
                    if (k < HT) {
                        if (k & 1) { ar1 = fmaf(whr[k], hk[e], ar1); az1 = fmaf(whz[k], hk[e], az1); an1 = fmaf(whn[k], hk[e], an1); }
                        else       { ar0 = fmaf(whr[k], hk[e], ar0); az0 = fmaf(whz[k], hk[e], az0); an0 = fmaf(whn[k], hk[e], an0); }
                    }
                }
            }
            const float r = sigmoidf_(ar0 + ar1);
            const float z = sigmoidf_(az0 + az1);
            const float hgn = an0 + an1;
            const float n = tanhf_(fmaf(r, hgn, xn));
            h = fmaf(h - n, z, n);
            cur ^= 1;
            if (lane < HP) shb[cur * HP + lane] = h;
            __syncwarp();
            float g = h;
            if constexpr (HEAD) {
                float p0 = bh, p1 = 0.f;
                const float4 *hn4 = reinterpret_cast<const float4 *>(shb + cur * HP);
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
            }
            spo0[tl * 33 + lane] = wo0 * g;
            spo1[tl * 33 + lane] = wo1 * g;
            if (sv && act) {
                float *s = sv + (size_t)(t0 + tl) * L.NS * H + lane;
                s[0] = r; s[H] = z; s[2 * H] = n; s[3 * H] = hgn; s[4 * H] = h;
                if constexpr (HEAD) s[5 * H] = g;
            }
        }
        __syncwarp();
        // ---- phase C: one timestep per lane — output reduction, store, squared error
        if (lane < nt) {
            float o0 = bo0, o1 = bo1;
            for (int k = 0; k < H; ++k) { o0 += spo0[lane * 33 + k]; o1 += spo1[lane * 33 + k]; }
            if constexpr (HEAD) {
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    const float fv = sfeat[lane * 8 + f];
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
        __syncwarp();
    }
    if (a.loss && y2) {
        lsum = warp_sum(lsum);
        if (lane == 0) atomicAdd(a.loss, (double)lsum * (double)a.loss_scale);
    }
}

// ================================================================ backward
// SPLIT = the F "feature lanes" (which turn the broadcast gate gradients into dL/dfeatures) do not fit next to
// the H unit lanes in one warp (H+F>32): lanes 0..F-1 then carry a second weight column.
template <int HT, int FM, int HEAD, bool DW>
__global__ void __launch_bounds__(128) gru_bwd_kernel(GruArgs a) {
    constexpr int F = FeatN<FM>::value, HP = Pad4<HT>::value;
    constexpr bool SPLIT = (HT + F > 32);
    const GruLayout<FM, HEAD> L(a.H);
    const int H = a.H;
    extern __shared__ __align__(16) float smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    float *sp = smem + 4;
    const int Ppad = (L.P + 3) & ~3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    float *ws = sp + Ppad + warp * bwd_warp_floats<HT>();
    float *sfeat = ws;               // [32][8]
    float *sgo = sfeat + 32 * 8;     // [32][2]
    float *sdf = sgo + 32 * 2;       // [32][8]
    float *sG = sdf + 32 * 8;        // [2][4*HP]   ar | az | an*r | an
    float *sdp = sG + 2 * 4 * HP;    // [2][HP]

    stage_params(sp, a.params, L.P, bar);
    const int b = blockIdx.x * wpc + warp;
    if (b >= a.B) return;

    const bool act = lane < H;
    const int j = act ? lane : 0;
    const int fl = SPLIT ? lane : lane - H;               // feature index this lane serves (if 0<=fl<F)
    const bool isf = fl >= 0 && fl < F;
    // weight columns
    float wcol[3 * HT], wicol[SPLIT ? 3 * HT : 1], whT[HEAD ? HT : 1];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int k = 0; k < HT; ++k) {
            float w = 0.f;
            if (k < H) {
                if (act) w = sp[L.oWhh + (g * H + k) * H + j];
                else if (!SPLIT && isf) w = sp[L.oWih + (g * H + k) * F + fl];
            }
            wcol[g * HT + k] = w;
            if constexpr (SPLIT) wicol[g * HT + k] = (k < H && isf) ? sp[L.oWih + (g * H + k) * F + fl] : 0.f;
        }
    if constexpr (HEAD) {
#pragma unroll
        for (int k = 0; k < HT; ++k) whT[k] = (act && k < H) ? sp[L.oWh + k * H + j] : 0.f;
    }
    const float wo0 = act ? sp[L.oWo + j] : 0.f, wo1 = act ? sp[L.oWo + L.O + j] : 0.f;
    const float wof0 = (HEAD && isf) ? sp[L.oWo + H + fl] : 0.f, wof1 = (HEAD && isf) ? sp[L.oWo + L.O + H + fl] : 0.f;

    // gradient accumulators (registers)
    float gwhh[DW ? 3 * HT : 1], gwih[DW ? 3 * F : 1], gwh[(DW && HEAD) ? HT : 1];
    float gb_r = 0.f, gb_z = 0.f, gb_n = 0.f, gb_hn = 0.f, gwo0 = 0.f, gwo1 = 0.f, gbh = 0.f, gbo0 = 0.f, gbo1 = 0.f;
    float gwof[(DW && HEAD) ? 2 * F : 1];
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

    const float2 *x2 = reinterpret_cast<const float2 *>(a.x) + (size_t)b * a.T;
    const float2 *go2 = a.gout ? reinterpret_cast<const float2 *>(a.gout) + (size_t)b * a.T : nullptr;
    const float2 *oi2 = a.out_in ? reinterpret_cast<const float2 *>(a.out_in) + (size_t)b * a.T : nullptr;
    const float2 *y2 = a.target ? reinterpret_cast<const float2 *>(a.target) + (size_t)b * a.T : nullptr;
    float2 *gx2 = (a.need_dx && a.gx) ? reinterpret_cast<float2 *>(a.gx) + (size_t)b * a.T : nullptr;
    const float *sv = a.saved + (size_t)b * a.T * L.NS * H;
    const float gs = a.gscale * (a.gscale_dev ? __ldg(a.gscale_dev) : 1.0f);

    float gH = 0.f;
    int cur = 0;
    const int nchunks = (a.T + ODPD_CHUNK - 1) / ODPD_CHUNK;
    for (int c = nchunks - 1; c >= 0; --c) {
        const int t0 = c * ODPD_CHUNK, nt = min(ODPD_CHUNK, a.T - t0);
        // ---- phase A': per-lane timestep: features (recomputed) and dLoss/dout
        float my_go0 = 0.f, my_go1 = 0.f;
        if (lane < nt) {
            const float2 v = __ldg(x2 + t0 + lane);
            float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            features_fwd<FM>(v.x, v.y, 0.f, 0.f, f);
            float4 *d = reinterpret_cast<float4 *>(sfeat + lane * 8);
            d[0] = make_float4(f[0], f[1], f[2], f[3]);
            d[1] = make_float4(f[4], f[5], f[6], f[7]);
            if (go2) { const float2 g = __ldg(go2 + t0 + lane); my_go0 = g.x; my_go1 = g.y; }
            else { const float2 o = __ldg(oi2 + t0 + lane), y = __ldg(y2 + t0 + lane); my_go0 = gs * (o.x - y.x); my_go1 = gs * (o.y - y.y); }
            *reinterpret_cast<float2 *>(sgo + lane * 2) = make_float2(my_go0, my_go1);
            if constexpr (DW) {
                gbo0 += my_go0; gbo1 += my_go1;
                if constexpr (HEAD) {
#pragma unroll
                    for (int q = 0; q < F; ++q) { gwof[q] = fmaf(my_go0, f[q], gwof[q]); gwof[F + q] = fmaf(my_go1, f[q], gwof[F + q]); }
                }
            }
        }
        __syncwarp();
        // ---- phase B': serial reverse-time recurrence
        for (int tl = nt - 1; tl >= 0; --tl) {
            const int t = t0 + tl;
            float r = 0.f, z = 0.f, n = 0.f, hgn = 0.f, ht = 0.f, g = 0.f, hp = 0.f;
            if (act) {
                const float *s = sv + (size_t)t * L.NS * H + lane;
                r = __ldg(s); z = __ldg(s + H); n = __ldg(s + 2 * H); hgn = __ldg(s + 3 * H); ht = __ldg(s + 4 * H);
                if constexpr (HEAD) g = __ldg(s + 5 * H);
                if (t > 0) hp = __ldg(s - (size_t)L.NS * H + 4 * H);
            }
            const float2 go = *reinterpret_cast<const float2 *>(sgo + tl * 2);
            float feat[8];
            {
                const float4 *fp = reinterpret_cast<const float4 *>(sfeat + tl * 8);
                const float4 f0 = fp[0];
                feat[0] = f0.x; feat[1] = f0.y; feat[2] = f0.z; feat[3] = f0.w;
                if (F > 4) { const float4 f1 = fp[1]; feat[4] = f1.x; feat[5] = f1.y; feat[6] = f1.z; feat[7] = f1.w; }
            }
            // head backward
            if constexpr (HEAD) {
                const float dg = fmaf(wo0, go.x, wo1 * go.y);
                const float dpre = g > 0.f ? dg : 0.f;
                if constexpr (DW) { gwo0 = fmaf(go.x, g, gwo0); gwo1 = fmaf(go.y, g, gwo1); gbh += dpre; }
                if (lane < HP) sdp[cur * HP + lane] = dpre;
                __syncwarp();
                float dh0 = 0.f, dh1 = 0.f;
                const float4 *dp4 = reinterpret_cast<const float4 *>(sdp + cur * HP);
#pragma unroll
                for (int k4 = 0; k4 < HP / 4; ++k4) {
                    const float4 dv = dp4[k4];
                    const float dk[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int k = k4 * 4 + e;
                        if (k < HT) {
                            if (k & 1) dh1 = fmaf(whT[k], dk[e], dh1); else dh0 = fmaf(whT[k], dk[e], dh0);
                            if constexpr (DW) gwh[k] = fmaf(dk[e], ht, gwh[k]);
                        }
                    }
                }
                gH += dh0 + dh1;
            } else {
                gH += fmaf(wo0, go.x, wo1 * go.y);
                if constexpr (DW) { gwo0 = fmaf(go.x, ht, gwo0); gwo1 = fmaf(go.y, ht, gwo1); }
            }
            // cell backward
            const float gz = gH * (hp - n), gn = gH * (1.f - z), ghp = gH * z;
            const float an = gn * (1.f - n * n);
            const float az = gz * z * (1.f - z);
            const float anr = an * r;
            const float ar = anr * hgn * (1.f - r);
            if constexpr (DW) {
#pragma unroll
                for (int q = 0; q < F; ++q) {
                    gwih[q] = fmaf(ar, feat[q], gwih[q]);
                    gwih[F + q] = fmaf(az, feat[q], gwih[F + q]);
                    gwih[2 * F + q] = fmaf(an, feat[q], gwih[2 * F + q]);
                }
                gb_r += ar; gb_z += az; gb_n += an; gb_hn += anr;
            }
            if (lane < HP) {
                float *G = sG + cur * 4 * HP + lane;
                G[0] = ar; G[HP] = az; G[2 * HP] = anr; G[3 * HP] = an;
            }
            __syncwarp();
            float acc0 = 0.f, acc1 = 0.f, fa0 = 0.f, fa1 = 0.f;
            const float4 *G4 = reinterpret_cast<const float4 *>(sG + cur * 4 * HP);
#pragma unroll
            for (int k4 = 0; k4 < HP / 4; ++k4) {
                const float4 v_r = G4[k4], v_z = G4[HP / 4 + k4], v_nh = G4[2 * (HP / 4) + k4], v_nx = G4[3 * (HP / 4) + k4];
                const float kr[4] = {v_r.x, v_r.y, v_r.z, v_r.w}, kz[4] = {v_z.x, v_z.y, v_z.z, v_z.w};
                const float knh[4] = {v_nh.x, v_nh.y, v_nh.z, v_nh.w}, knx[4] = {v_nx.x, v_nx.y, v_nx.z, v_nx.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = k4 * 4 + e;
                    if (k < HT) {
                        const float vn = SPLIT ? knh[e] : (act ? knh[e] : knx[e]);
                        acc0 = fmaf(wcol[k], kr[e], acc0);
                        acc1 = fmaf(wcol[HT + k], kz[e], acc1);
                        acc0 = fmaf(wcol[2 * HT + k], vn, acc0);
                        if constexpr (SPLIT) {
                            fa0 = fmaf(wicol[k], kr[e], fa0);
                            fa1 = fmaf(wicol[HT + k], kz[e], fa1);
                            fa0 = fmaf(wicol[2 * HT + k], knx[e], fa0);
                        }
                        if constexpr (DW) {
                            gwhh[k] = fmaf(kr[e], hp, gwhh[k]);
                            gwhh[HT + k] = fmaf(kz[e], hp, gwhh[HT + k]);
                            gwhh[2 * HT + k] = fmaf(knh[e], hp, gwhh[2 * HT + k]);
                        }
                    }
                }
            }
            const float acc = acc0 + acc1;
            if (a.need_dx && isf) {
                const float df = (SPLIT ? (fa0 + fa1) : acc) + (HEAD ? fmaf(wof0, go.x, wof1 * go.y) : 0.f);
                sdf[tl * 8 + fl] = df;
            }
            gH = act ? ghp + acc : 0.f;
            cur ^= 1;
        }
        __syncwarp();
        // ---- phase C': per-lane timestep: dL/dfeatures -> dL/d(I,Q), coalesced store
        if (gx2 && lane < nt) {
            const float2 v = __ldg(x2 + t0 + lane);
            float gf[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) gf[q] = (q < F) ? sdf[lane * 8 + q] : 0.f;
            float gi, gq;
            features_bwd<FM>(v.x, v.y, gf, gi, gq);
            gx2[t0 + lane] = make_float2(gi, gq);
        }
        __syncwarp();
    }

    if constexpr (DW) if (a.partials) {
        float *pr = a.partials + (size_t)b * L.P;
        if (act) {
#pragma unroll
            for (int g = 0; g < 3; ++g) {
#pragma unroll
                for (int q = 0; q < F; ++q) pr[L.oWih + (g * H + lane) * F + q] = gwih[g * F + q];
#pragma unroll
                for (int k = 0; k < HT; ++k)
                    if (k < H) pr[L.oWhh + (g * H + k) * H + lane] = gwhh[g * HT + k];
            }
            pr[L.obih + lane] = gb_r; pr[L.obih + H + lane] = gb_z; pr[L.obih + 2 * H + lane] = gb_n;
            pr[L.obhh + lane] = gb_r; pr[L.obhh + H + lane] = gb_z; pr[L.obhh + 2 * H + lane] = gb_hn;
            pr[L.oWo + lane] = gwo0; pr[L.oWo + L.O + lane] = gwo1;
            if constexpr (HEAD) {
#pragma unroll
                for (int k = 0; k < HT; ++k)
                    if (k < H) pr[L.oWh + k * H + lane] = gwh[k];
                pr[L.obh + lane] = gbh;
            }
        }
        gbo0 = warp_sum(gbo0); gbo1 = warp_sum(gbo1);
        if (lane == 0) { pr[L.obo] = gbo0; pr[L.obo + 1] = gbo1; }
        if constexpr (HEAD) {
#pragma unroll
            for (int q = 0; q < 2 * F; ++q) {
                const float s = warp_sum(gwof[q]);
                if (lane == 0) pr[L.oWo + (q / F) * L.O + H + (q % F)] = s;
            }
        }
    }
}

// ================================================================ host dispatch
template <int FM, int HEAD> static int64_t gru_nparams(int H) { return GruLayout<FM, HEAD>(H).P; }

static int pick_wpc(int B) { return B > 148 * 16 ? 4 : 1; }

template <int HT, int FM, int HEAD>
static int launch_fwd(const GruArgs &a, cudaStream_t st) {
    const GruLayout<FM, HEAD> L(a.H);
    const int wpc = pick_wpc(a.B);
    const size_t smem = (4 + ((L.P + 3) & ~3) + wpc * fwd_warp_floats<HT>()) * sizeof(float);
    auto k = gru_fwd_kernel<HT, FM, HEAD>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<(a.B + wpc - 1) / wpc, wpc * 32, smem, st>>>(a);
    return check_launch("gru_fwd_kernel");
}
template <int HT, int FM, int HEAD>
static int launch_bwd(const GruArgs &a, bool dw, cudaStream_t st) {
    const GruLayout<FM, HEAD> L(a.H);
    const int wpc = pick_wpc(a.B);
    const size_t smem = (4 + ((L.P + 3) & ~3) + wpc * bwd_warp_floats<HT>()) * sizeof(float);
    if (dw) {
        auto k = gru_bwd_kernel<HT, FM, HEAD, true>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<(a.B + wpc - 1) / wpc, wpc * 32, smem, st>>>(a);
    } else {
        auto k = gru_bwd_kernel<HT, FM, HEAD, false>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<(a.B + wpc - 1) / wpc, wpc * 32, smem, st>>>(a);
    }
    return check_launch("gru_bwd_kernel");
}

// compiled hidden-size tiers: exact for the sizes the reference's scripts use, next-larger tier otherwise
#define ODPD_GRU_TIERS(X, FM, HEAD) X(8, FM, HEAD) X(10, FM, HEAD) X(13, FM, HEAD) X(16, FM, HEAD) X(23, FM, HEAD) X(32, FM, HEAD)

template <int FM, int HEAD>
static int dispatch(const GruArgs &a, int dir, bool dw, cudaStream_t st) {
#define X(HTV, FMV, HEADV)                                                                  \
    if (a.H <= HTV) return dir == 0 ? launch_fwd<HTV, FMV, HEADV>(a, st) : launch_bwd<HTV, FMV, HEADV>(a, dw, st);
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
int64_t gru_family_saved_floats(int cell, int B, int T, int H) { return (int64_t)B * T * (cell == ODPD_CELL_DGRU ? 6 : 5) * H; }

int gru_family_run(int cell, const GruArgs &a, int dir, bool dw, cudaStream_t st) {
    switch (cell) {
    case ODPD_CELL_GRU: return dispatch<FM_RAW2, 0>(a, dir, dw, st);
    case ODPD_CELL_DGRU: return dispatch<FM_DGRU6, 1>(a, dir, dw, st);
    case ODPD_CELL_QGRU: return dispatch<FM_QGRU4, 0>(a, dir, dw, st);
    case ODPD_CELL_QGRU_AMP1: return dispatch<FM_AMP4, 0>(a, dir, dw, st);
    }
    set_error("gru_family_run: bad cell %d", cell);
    return -1;
}

}  // namespace odpd

// pipeline.cuh — pieces shared by the warp-specialised chunk pipelines (pre / chain / post warps).
#pragma once
#include "common.cuh"

namespace odpd {

static constexpr int CH = ODPD_CHUNK;   // timesteps per pipeline stage
template <int HT> struct Pad4 { static constexpr int value = (HT + 3) & ~3; };

// ---------------------------------------------------------------- mbarrier + TMA 1-D bulk copies (SASS: SYNCS.*, UBLKCP.S.G / UBLKCP.G.S)
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_1d(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// every thread that wrote shared memory the bulk store will read calls this before the barrier that precedes the store
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_1d(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// acc += sum_k w[k] * line[k]  — `line` is an HP-float shared-memory row read as broadcast LDS.128; two accumulators
template <int HT>
__device__ __forceinline__ void bcast_dot(const float *line, const float (&w)[HT], float &a0, float &a1) {
    constexpr int HP = Pad4<HT>::value;
    const float4 *l4 = reinterpret_cast<const float4 *>(line);
    float4 v[HP / 4];
#pragma unroll
    for (int q = 0; q < HP / 4; ++q) v[q] = l4[q];
#pragma unroll
    for (int q = 0; q < HP / 4; ++q) {
        const float e[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = q * 4 + i;
            if (k < HT) { if (k & 1) a1 = fmaf(w[k], e[i], a1); else a0 = fmaf(w[k], e[i], a0); }
        }
    }
}

// Linear output head + squared error for one chunk (post warp):  out_t = W_o h_t (+ b_o) (+ skip_t)
//   rows  : chunk activation rows in shared memory, h_t at rows[tl*ROW + hoff + k]
//   spo   : scratch [2][CH][33]
//   skip  : optional per-step additive term (float2 per step, shared memory) or nullptr
//   hoff1 : offset of the vector feeding the second output (== hoff for every cell but DVRJANET: y_I from h_I, y_Q from h_Q)
__device__ __forceinline__ void linear_head_chunk2(const float *rows, int ROW, int hoff, int hoff1, int HP, int H, int nt, int lane, float wo0,
                                                   float wo1, float bo0, float bo1, float *spo, const float2 *skip, float2 *out,
                                                   IqRow tgt, float &lsum) {
    float *spo0 = spo, *spo1 = spo + CH * 33;
#pragma unroll 4
    for (int tl = 0; tl < nt; ++tl) {
        const float h0 = lane < HP ? rows[tl * ROW + hoff + lane] : 0.f;
        const float h1 = lane < HP ? rows[tl * ROW + hoff1 + lane] : 0.f;
        spo0[tl * 33 + lane] = wo0 * h0;
        spo1[tl * 33 + lane] = wo1 * h1;
    }
    __syncwarp();
    if (lane < nt) {
        float o0 = bo0, o1 = bo1;
        for (int k = 0; k < H; ++k) { o0 += spo0[lane * 33 + k]; o1 += spo1[lane * 33 + k]; }
        if (skip) { const float2 sk = skip[lane]; o0 += sk.x; o1 += sk.y; }
        out[lane] = make_float2(o0, o1);
        if (tgt) {
            const float2 y = __ldg(tgt + lane);
            const float d0 = o0 - y.x, d1 = o1 - y.y;
            lsum = fmaf(d0, d0, fmaf(d1, d1, lsum));
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void linear_head_chunk(const float *rows, int ROW, int hoff, int HP, int H, int nt, int lane, float wo0, float wo1,
                                                  float bo0, float bo1, float *spo, const float2 *skip, float2 *out, IqRow tgt,
                                                  float &lsum) {
    linear_head_chunk2(rows, ROW, hoff, hoff, HP, H, nt, lane, wo0, wo1, bo0, bo1, spo, skip, out, tgt, lsum);
}

// dLoss/dout for one chunk, one timestep per lane: explicit gout tensor or fused MSE gradient gs*(out-target)
__device__ __forceinline__ float2 load_gout(const float2 *go2, const float2 *oi2, IqRow y2, int t, float gs) {
    if (go2) return __ldg(go2 + t);
    const float2 o = __ldg(oi2 + t), y = __ldg(y2 + t);
    return make_float2(gs * (o.x - y.x), gs * (o.y - y.y));
}

// TMA bulk load of the saved rows of chunk [t0, t0+nt) plus the row of step t0-1 into `ac` (row 0 = step t0-1; zero when t0==0)
__device__ __forceinline__ void load_rows_with_prev(float *ac, const float *svg, int ROW, int t0, int nt, int lane, uint64_t *bar) {
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
    if (t0 == 0) { for (int i = lane; i < ROW; i += 32) ac[i] = 0.f; }
}

}  // namespace odpd

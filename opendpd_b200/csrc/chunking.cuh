// chunking.cuh — time-chunked execution of the contractive recurrences (GRU family, LSTM, PGJANET, DVRJANET); see
// include/odpd.h "Time-chunked execution" and DESIGN.md §4.
//
// A gated cell started from the wrong state forgets it geometrically, so a T-step sequence can be cut into C chunks that run
// CONCURRENTLY (one CTA each), every chunk preceded by Wu warm-up steps started from the zero state.  A verify pass then compares
// the state every chunk was started from with the state its predecessor really ended with and re-runs, serially, every sequence
// with a failing boundary — results never depend on the forgetting assumption, only the speed does.  The backward pass is the
// same construction on the (linear) adjoint recurrence in reverse time.  The delta cells are excluded: their running delta
// memories are fp32 prefix sums whose rounding depends on the whole history, and their masks must match the reference bit for bit.
#pragma once
#include "cells.h"

namespace odpd {

static constexpr int SPEC_ROWS_AUTO = 2048, SPEC_CMAX = 32, SPEC_WARM_DEFAULT = 128;
// Boundary tolerance, relative to max|state| (forward) / max|adjoint state| (backward) at the boundary: 2^-18 = 3.8e-6.
// It cannot be much tighter: two fp32 evaluations of the SAME converged trajectory differ by rounding noise amplified by
// 1/(1 - forgetting rate) — measured floor (max over ~5e4 boundaries, warm-up 256 and 512 alike): 0.7e-6 DGRU, 1.3e-6 PGJANET,
// 1.1e-6 DVRJANET forward, 1.6e-6 backward — the same size as the GPU-vs-CPU-reference difference of the serial kernels.
static constexpr float SPEC_TOL_FWD = 3.8146973e-06f;
static constexpr float SPEC_TOL_BWD = 3.8146973e-06f;

// ---------------------------------------------------------------- host: plan + scratch layout
// rows of per-(sequence,chunk) scratch / gradient partials a call of B sequences may use
inline int64_t chunk_rows(int B, int tchunks_req) {
    if (B <= 0) return 1;
    if (tchunks_req == 1) return B;
    if (tchunks_req > 1) return (int64_t)B * (tchunks_req < SPEC_CMAX ? tchunks_req : SPEC_CMAX);
    return B > SPEC_ROWS_AUTO ? B : SPEC_ROWS_AUTO;
}
// forward scratch (tail of `saved`):  guess[rows][ss] | end[rows][ss] | loss[rows] | int32 re-run counter | float worst boundary
// mismatch seen so far in units of the tolerance (<= 1 passes) | pad
inline int64_t chunk_fwd_scratch_floats(int64_t rows, int ss) { return rows * (2 * ss + 1) + 4; }
// backward scratch (tail of the workspace, after the 4-float-aligned [rows][P] partials):  guess | end | re-run counter (+pad)
inline int64_t chunk_bwd_scratch_floats(int64_t rows, int ss) { return rows * 2 * ss + 4; }
inline int64_t chunk_workspace_floats(int64_t rows, int64_t P, int ss) { return ((rows * P + 3) & ~(int64_t)3) + chunk_bwd_scratch_floats(rows, ss); }

// Per-process caches are keyed by the CURRENT device: cudaFuncSetAttribute and occupancy are per device, and one process may
// touch several GPUs (tests do).
static constexpr int ODPD_MAX_DEV = 32;
struct OccCache { int v[ODPD_MAX_DEV]; size_t smem[ODPD_MAX_DEV]; };   // occupancy and the dynamic-smem limit it was taken at, per device
inline int cur_dev_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= ODPD_MAX_DEV) dev = 0;
    return dev;
}
inline int num_sms() {
    static int n[ODPD_MAX_DEV] = {0};
    const int dev = cur_dev_slot();
    if (!n[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n[dev] = v;
    }
    return n[dev];
}

// fills a.C / a.Lc / a.Wu;  slots = CTAs of this kernel the device holds at once
inline void chunk_make_plan(GruArgs &a, int slots, bool have_scratch) {
    a.C = 1; a.Lc = a.T; a.Wu = 0;
    const int req = a.tchunks_req;
    if (req == 1 || !have_scratch || a.T < 2 * ODPD_CHUNK) return;
    constexpr int CHK = ODPD_CHUNK;
    const int Wu = a.twarm_req > 0 ? ((a.twarm_req + CHK - 1) / CHK) * CHK : (a.twarm_default > 0 ? a.twarm_default : SPEC_WARM_DEFAULT);
    const int nblk = (a.T + CHK - 1) / CHK;
    auto lc_of = [&](int C) { return ((nblk + C - 1) / C) * CHK; };
    auto valid = [&](int C) { return (int64_t)(C - 1) * lc_of(C) < a.T; };
    int C = 1;
    if (req > 1) {
        C = req < SPEC_CMAX ? req : SPEC_CMAX;
        while (C > 1 && !valid(C)) --C;
    } else {
        // cost model: every CTA walks Lc + Wu steps; CTAs beyond what the device holds at once wait for a second wave
        if (slots > SPEC_ROWS_AUTO) slots = SPEC_ROWS_AUTO;
        if (slots < 1) slots = 1;
        int64_t best = (int64_t)((a.B + slots - 1) / slots) * a.T;
        for (int c = 2; c <= SPEC_CMAX; ++c) {
            if (!valid(c) || lc_of(c) < Wu || (int64_t)a.B * c > SPEC_ROWS_AUTO) continue;
            const int64_t cost = (int64_t)(((int64_t)a.B * c + slots - 1) / slots) * (lc_of(c) + Wu);
            if (cost < best) { best = cost; C = c; }
        }
    }
    if (C > 1 && (int64_t)a.B * C <= chunk_rows(a.B, req)) { a.C = C; a.Lc = lc_of(C); a.Wu = Wu; }
}

// Plan, bind the scratch and launch:  [B*C chunk CTAs] + [B verify CTAs]  or, with one chunk,  [B serial CTAs].
//   scr      start of the scratch inside the caller's buffer (nullptr -> serial), scr_off its float offset (for info[3])
//   ss       floats of recurrent state per (sequence, chunk)
//   info     optional out: chunks, steps per chunk, warm-up steps, float index of the re-run counter
//   warm     default warm-up of the cell (steps; the caller's OdpdDims.twarm overrides it)
template <typename K>
inline int chunk_launch(K k, int nthreads, size_t smem, OccCache *occ_all, GruArgs a, int dir, float *scr, int64_t scr_off, int ss,
                        cudaStream_t st, bool plan_only, int *info, const char *what, int warm = SPEC_WARM_DEFAULT) {
    a.twarm_default = warm;
    const int dev_slot = cur_dev_slot();
    int *occ_cache = &occ_all->v[dev_slot];
    // the footprint depends on the runtime hidden size inside a compiled tier (the parameter block): re-arm when a larger one shows up
    if (!*occ_cache || smem > occ_all->smem[dev_slot]) {
        occ_all->smem[dev_slot] = smem;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, nthreads, smem) != cudaSuccess || occ <= 0) occ = 1;
        *occ_cache = occ;
    }
    const int64_t rows = chunk_rows(a.B, a.tchunks_req);
    chunk_make_plan(a, *occ_cache * num_sms(), scr != nullptr || plan_only);
    const int64_t fail_off = scr_off + (dir == 0 ? rows * (2 * ss + 1) : rows * 2 * ss);
    if (info) { info[0] = a.C; info[1] = a.Lc; info[2] = a.Wu; info[3] = (plan_only && fail_off <= 0x7fffffff) ? (int)fail_off : -1; }
    if (plan_only) return 0;
    if (a.C > 1) {
        a.sc_guess = scr; a.sc_end = scr + rows * ss;
        a.sc_loss = dir == 0 ? scr + 2 * rows * ss : nullptr;
        a.sc_fail = reinterpret_cast<int *>(scr + (dir == 0 ? rows * (2 * ss + 1) : rows * 2 * ss));
        a.tol = dir == 0 ? SPEC_TOL_FWD : SPEC_TOL_BWD;
        a.mode = 0;
        launch_pdl(k, dim3(a.B * a.C), dim3(nthreads), smem, st, a);
        a.mode = 2;
        launch_pdl(k, dim3(a.B), dim3(nthreads), smem, st, a);
    } else {
        a.mode = 0;
        launch_pdl(k, dim3(a.B), dim3(nthreads), smem, st, a);
    }
    return check_launch(what);
}

// ---------------------------------------------------------------- device: the time range of one CTA
struct FwdRange {   // warm-up [t_lo, t_emit) from the zero state (nothing emitted), then the emitted steps [t_emit, t_hi)
    int b, cc, t_lo, t_emit, t_hi;
    bool spec;
};
__device__ __forceinline__ FwdRange fwd_range(const GruArgs &a) {
    FwdRange r;
    r.spec = (a.C > 1 && a.mode == 0);
    r.b = blockIdx.x; r.cc = 0; r.t_lo = 0; r.t_emit = 0; r.t_hi = a.T;
    if (r.spec) {
        r.b = blockIdx.x / a.C; r.cc = blockIdx.x - r.b * a.C;
        r.t_emit = r.cc * a.Lc; r.t_lo = max(0, r.t_emit - a.Wu); r.t_hi = min(a.T, r.t_emit + a.Lc);
    }
    return r;
}
struct BwdRange {   // reverse time: warm-up steps [t_ehi, t_hi) from a zero adjoint (nothing emitted), then [t_elo, t_ehi)
    int b, cc, t_elo, t_ehi, t_hi;
    bool spec;
};
__device__ __forceinline__ BwdRange bwd_range(const GruArgs &a) {
    BwdRange r;
    r.spec = (a.C > 1 && a.mode == 0);
    r.b = blockIdx.x; r.cc = 0; r.t_elo = 0; r.t_ehi = a.T; r.t_hi = a.T;
    if (r.spec) {
        r.b = blockIdx.x / a.C; r.cc = blockIdx.x - r.b * a.C;
        r.t_elo = r.cc * a.Lc; r.t_ehi = min(a.T, r.t_elo + a.Lc); r.t_hi = min(a.T, r.t_ehi + a.Wu);
    }
    return r;
}

// Verify pass of the forward (mode 2, one CTA per sequence): the state chunk c was started from (after its warm-up) against the
// state chunk c-1 ended with.  Returns true when every boundary of sequence b holds — the CTA then adds the sequence's squared
// error to the loss and exits; false -> the caller recomputes the sequence serially.   State layout: ns vectors of HP floats
// (stride ss), components >= H are padding.
// records the worst mismatch/tolerance ratio (non-negative floats order like their bit patterns)
__device__ __forceinline__ void chunk_note_ratio(const GruArgs &a, float ratio) {
    if (ratio == ratio) atomicMax(a.sc_fail + 1, __float_as_int(fminf(ratio, 3.0e38f)));
}
// One boundary: `g` = state the chunk was started from, `e` = state its neighbour really produced (ss floats, components with
// k % HP >= H are padding).  Passes when max|g-e| <= tol * max|e|.
__device__ __forceinline__ bool chunk_boundary_ok(const GruArgs &a, const float *g, const float *e, int ss, int HP, int H) {
    float m = 0.f, dmax = 0.f;
    for (int k = 0; k < ss; ++k)
        if (k % HP < H) { m = fmaxf(m, fabsf(e[k])); dmax = fmaxf(dmax, fabsf(g[k] - e[k])); }
    const float lim = a.tol * m + 1e-37f;
    const bool ok = (dmax <= lim) && (m == m) && (dmax == dmax);
    chunk_note_ratio(a, dmax / lim);
    return ok;
}
__device__ __forceinline__ bool fwd_verify_pass(const GruArgs &a, int b, int ss, int HP, int H) {
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < a.C - 1; i += blockDim.x)   // boundary between chunks i and i+1
        if (!chunk_boundary_ok(a, a.sc_guess + ((size_t)b * a.C + i + 1) * ss, a.sc_end + ((size_t)b * a.C + i) * ss, ss, HP, H)) s_bad = 1;
    __syncthreads();
    if (s_bad) {
        if (threadIdx.x == 0) atomicAdd(a.sc_fail, 1);
        return false;
    }
    if (threadIdx.x == 0 && a.loss && a.target) {
        float sl = 0.f;
        for (int c1 = 0; c1 < a.C; ++c1) sl += a.sc_loss[(size_t)b * a.C + c1];
        atomicAdd(a.loss, (double)sl * (double)a.loss_scale);
    }
    return true;
}
// Verify pass of the backward: the adjoint state chunk c was started from (after its warm-up over the steps that follow it)
// against what chunk c+1 really handed down.
__device__ __forceinline__ bool bwd_verify_pass(const GruArgs &a, int b, int ss, int HP, int H) {
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < a.C - 1; i += blockDim.x)
        if (!chunk_boundary_ok(a, a.sc_guess + ((size_t)b * a.C + i) * ss, a.sc_end + ((size_t)b * a.C + i + 1) * ss, ss, HP, H)) s_bad = 1;
    __syncthreads();
    if (s_bad) {
        if (threadIdx.x == 0) atomicAdd(a.sc_fail, 1);
        return false;
    }
    return true;
}

// loss of one forward CTA: chunk CTAs park it for the verify pass, serial CTAs add it to the caller's accumulator
__device__ __forceinline__ void chunk_store_loss(const GruArgs &a, bool spec, float lsum_warp_total) {
    if (spec) a.sc_loss[blockIdx.x] = lsum_warp_total;
    else if (a.loss) atomicAdd(a.loss, (double)lsum_warp_total * (double)a.loss_scale);
}

// gradient-partial row of one backward CTA: (sequence, chunk) when chunked, else the sequence's first row; the serial re-run of
// the verify pass clears the sequence's other rows (call from the `nthr` threads that write partials, tid = 0..nthr-1)
__device__ __forceinline__ float *chunk_partial_row(const GruArgs &a, bool spec, int b, int64_t P, int tid, int nthr) {
    float *prt = a.partials + (size_t)(spec ? blockIdx.x : b * a.C) * P;
    if (a.mode == 2)
        for (int64_t i = tid; i < (int64_t)(a.C - 1) * P; i += nthr) prt[P + i] = 0.f;
    return prt;
}

}  // namespace odpd

// common.cuh — device helpers shared by the sm_100a backbone kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/odpd.h"

#define ODPD_CHUNK 32           // timesteps per time-parallel phase (one per lane)
#define ODPD_FULL 0xffffffffu

namespace odpd {

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char *fmt, ...);
#define ODPD_CHECK(cond, ...)                \
    do {                                     \
        if (!(cond)) {                       \
            ::odpd::set_error(__VA_ARGS__);  \
            return -1;                       \
        }                                    \
    } while (0)
int check_launch(const char *what);

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// A train step is 6 small dependent kernels replayed from one CUDA graph; between two dependent kernels the GPU idles for the
// launch latency (~2 us each, ~7 % of the 160 us step).  With the programmatic-stream-serialization launch attribute the next
// kernel is launched while its predecessor still runs: its CTAs become resident as SMs drain and park in griddepcontrol.wait until
// the predecessor grid has completed and its memory is visible.  Every kernel of the step path calls pdl_enter() first thing, so
// nothing is ever read before the predecessor's results are final; kernels launched the ordinary way execute both as no-ops.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool pdl_enabled();   // env ODPD_PDL != 0 (api.cu)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------- math
// Gate nonlinearities: 2 MUFU ops each (ex2 + rcp), ~2e-7 relative error — far inside the 1e-5 parity budget and
// ~3x shorter dependent chain than expf()+IEEE divide; they sit on the serial critical path of every timestep.
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoidf_(float x) { return fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanhf_(float x) {
    // |x| >= 0.55: 1 - 2/(exp(2|x|)+1) (2 MUFU);  |x| < 0.55: x + x^3 Q(x^2), Q least-squares fitted (max rel err 6.4e-8 in
    // fp32).  The small-argument branch keeps RELATIVE accuracy near 0 — PGJANET multiplies three small tanh outputs, and the
    // exp form alone has a 2e-7 ABSOLUTE error there.  Both paths are evaluated and selected (no divergence).
    const float ax = fabsf(x);
    const float e = fast_ex2(2.8853900817779268f * ax);
    const float big = copysignf(fmaf(-2.0f, fast_rcp(e + 1.0f), 1.0f), x);
    const float x2 = x * x;
    float q = fmaf(x2, -6.296012253e-03f, 2.108818408e-02f);
    q = fmaf(x2, q, -5.385666501e-02f);
    q = fmaf(x2, q, 1.333263216e-01f);
    q = fmaf(x2, q, -3.333331912e-01f);
    const float small = fmaf(x * x2, q, x);
    return ax < 0.55f ? small : big;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(ODPD_FULL, v, o);
    return v;
}

// ---------------------------------------------------------------- features (bit-exact w.r.t. the reference's unfused ATen ops)
// gru.py:45 | dgru.py:61-68 / deltagru.py:61-72 | qgru.py:61-66 | qgru_amp1.py:61-70 | deltagru_tcnskip.py:91-100
enum { FM_RAW2 = 0, FM_DGRU6 = 1, FM_QGRU4 = 2, FM_AMP4 = 3, FM_TRES6 = 4 };
template <int FM> struct FeatN { static constexpr int value = (FM == FM_RAW2) ? 2 : ((FM == FM_QGRU4 || FM == FM_AMP4) ? 4 : 6); };

// f[0..F): features of sample (i,q); (in,qn) = next sample (TRES only).  No FMA contraction, IEEE sqrt/div:
// the delta-x mask must match the CPU reference bit for bit (SURVEY.md §7 hard part 1).
template <int FM>
__device__ __forceinline__ void features_fwd(float i, float q, float in, float qn, float *f) {
    f[0] = i;
    f[1] = q;
    if constexpr (FM == FM_RAW2) return;
    const float a2 = __fadd_rn(__fmul_rn(i, i), __fmul_rn(q, q));
    if constexpr (FM == FM_QGRU4) {
        f[2] = a2;
        f[3] = __fmul_rn(a2, a2);
        return;
    }
    const float a = __fsqrt_rn(a2);
    f[2] = a;
    f[3] = __fmul_rn(__fmul_rn(a, a), a);
    if constexpr (FM == FM_DGRU6) {
        f[4] = __fdiv_rn(q, a);  // sin
        f[5] = __fdiv_rn(i, a);  // cos
    }
    if constexpr (FM == FM_TRES6) {
        f[4] = in;
        f[5] = qn;
    }
}

// d(loss)/d(i,q) from d(loss)/d(features).  For TRES the (I_next,Q_next) grads are returned separately.
template <int FM>
__device__ __forceinline__ void features_bwd(float i, float q, const float *gf, float &gi, float &gq) {
    gi = gf[0];
    gq = gf[1];
    if constexpr (FM == FM_RAW2) return;
    const float a2 = i * i + q * q;
    if constexpr (FM == FM_QGRU4) {
        const float ga2 = gf[2] + 2.0f * a2 * gf[3];
        gi += 2.0f * i * ga2;
        gq += 2.0f * q * ga2;
        return;
    }
    const float a = sqrtf(a2);
    const float ga = gf[2] + 3.0f * a2 * gf[3];
    gi += ga * i / a;
    gq += ga * q / a;
    if constexpr (FM == FM_DGRU6) {
        // sin = q/a, cos = i/a.  Autograd's form (g_cos/a - (q g_sin + i g_cos) i / a^3) cancels catastrophically for small a;
        // the algebraically identical  d/di = q*w, d/dq = -i*w  with  w = (g_cos q - g_sin i)/a^3  does not.
        const float w = (gf[5] * q - gf[4] * i) / (a2 * a);
        gi = fmaf(q, w, gi);
        gq = fmaf(-i, w, gq);
    }
}

// ---------------------------------------------------------------- IQ sample rows
// One sequence's (T,2) IQ samples as the kernels see them: fp32 pairs, or bf16 pairs (ODPD_F_X_BF16 / ODPD_F_TARGET_BF16: storage only,
// every value is widened exactly to fp32), starting either at frame b of a framed (B,T,2) tensor or at sample starts[b] of a raw
// (N,2) stream (OdpdDims.x_starts / .target_starts: the framing of data_collector.py:233-252 done by address arithmetic).
// Pointer-like on purpose: `row + t` and `__ldg(row + t)` keep the kernels' load sites unchanged.
struct IqRow {
    const void *p;
    int bf16;
    __device__ __forceinline__ float2 ld(int t) const {
        if (bf16) {
            const unsigned v = ::__ldg(reinterpret_cast<const unsigned *>(p) + t);        // [I | Q] as two bf16: I in the low half
            return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
        }
        return ::__ldg(reinterpret_cast<const float2 *>(p) + t);
    }
    __device__ __forceinline__ IqRow operator+(int t) const {
        IqRow r;
        r.bf16 = bf16;
        r.p = bf16 ? static_cast<const void *>(reinterpret_cast<const unsigned *>(p) + t) : static_cast<const void *>(reinterpret_cast<const float2 *>(p) + t);
        return r;
    }
    __device__ __forceinline__ explicit operator bool() const { return p != nullptr; }
};
using ::__ldg;   // keep the built-in overloads visible next to the IqRow one
__device__ __forceinline__ float2 __ldg(const IqRow &r) { return r.ld(0); }
__device__ __forceinline__ IqRow iq_row(const void *base, int bf16, const int *starts, int b, int T) {
    IqRow r;
    r.bf16 = bf16;
    r.p = nullptr;
    if (base) {
        const size_t off = starts ? (size_t)::__ldg(starts + b) : (size_t)b * T;
        r.p = bf16 ? static_cast<const void *>(reinterpret_cast<const unsigned *>(base) + off) : static_cast<const void *>(reinterpret_cast<const float2 *>(base) + off);
    }
    return r;
}
__device__ __forceinline__ IqRow iq_none() { IqRow r; r.p = nullptr; r.bf16 = 0; return r; }

// ---------------------------------------------------------------- parameter staging: global -> shared via the TMA bulk-copy engine
// One elected thread issues cp.async.bulk (SASS: UBLKCP) on an mbarrier; everybody waits on the barrier phase.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_params(float *sdst, const float *gsrc, int n, uint64_t *bar) {
    const int n16 = (((uintptr_t)gsrc & 15) == 0) ? (n & ~3) : 0;  // bulk part: 16-byte granules
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (n16 > 0) {
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)n16 * 4u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                : "memory");
        }
    }
    for (int i = n16 + threadIdx.x; i < n; i += blockDim.x) sdst[i] = gsrc[i];
    if (n16 > 0) {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(smem_u32(bar)), "r"(0u)
                : "memory");
        }
    }
    __syncthreads();
}

// Data-parallel publish fused into the gradient reduction (csrc/dp.cu has the receive side): when world > 1 every reduced gradient
// element (and the loss) is also pushed as an 8-byte {value, step tag} word into slot [parity][rank] of every rank's receive buffer
// (own buffer included), straight from the reduction's epilogue — the NVLink flight overlaps the launch of the optimiser kernel.
static constexpr int ODPD_DP_MAX_WORLD = 8;
struct DpPushArgs {
    uint2 *buf[ODPD_DP_MAX_WORLD];
    int world, rank;
    int64_t stride;
    const int64_t *step_dev;
    const double *loss_local;
};
// Optimiser fused into the gradient reduction (single GPU): the last CTA of reduce_partials_kernel to finish (atomic ticket) runs
// clip_grad_norm_ + AdamW on the flat buffers, saving one launch per train step.  p == nullptr disables it.
struct AdamFuseArgs {
    float *p, *m, *v;
    const float *lr_dev;
    int64_t *step_dev;
    float *gnorm_out;
    int *ticket;            // device int, zero before the first use; the last CTA resets it
    float b1, b2, eps, wd, max_norm;
};
// Deterministic second-stage reduction of per-sequence gradient partials:  g[p] += sum_b part[b][p]
__global__ void reduce_partials_kernel(const float *__restrict__ part, int nrows, int64_t P, float *__restrict__ g, int overwrite, DpPushArgs push,
                                       AdamFuseArgs adam);

}  // namespace odpd

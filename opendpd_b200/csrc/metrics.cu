// metrics.cu — the evaluation metrics of modules/train_funcs.py:93-105 (calculate_metrics) on the device, SURVEY.md §8 row f-1:
// NMSE (utils/metrics.py:42-53) as per-row double sums, and the spectra behind EVM (:56-111, np.fft.fft(x, n=nperseg)) and ACLR
// (:114-190, scipy.signal.welch: periodic Hann window, constant detrend, 50 % overlap) as one windowed-DFT kernel that writes
// fftshift-ed bin magnitudes.  The band bookkeeping (index_left/right, sub-channels, dB) is a few dozen scalars and stays on the
// host side (opendpd_b200/metrics.py) exactly as the reference spells it.
//
// Sizes are tiny (the reference evaluates S <= 8..256 segments of nperseg = 2560 samples), so the DFT is evaluated directly:
// one thread per output bin, the segment and a double-precision twiddle table in shared memory, double accumulation — the side
// channels ACLR looks at sit 40-60 dB below the carrier and must not drown in fp32 rounding noise of the carrier bins.
#include "cells.h"

namespace odpd {

// out[row] = { sum |truth - pred|^2 , sum |truth|^2 }
__global__ void __launch_bounds__(256) nmse_sums_kernel(const float2 *__restrict__ pred, const float2 *__restrict__ truth, int N,
                                                       double *__restrict__ out) {
    __shared__ double red[2][8];
    const float2 *p = pred + (size_t)blockIdx.x * N, *g = truth + (size_t)blockIdx.x * N;
    double e = 0.0, r = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float2 a = __ldg(p + n), b = __ldg(g + n);
        const double d0 = (double)b.x - (double)a.x, d1 = (double)b.y - (double)a.y;
        e += d0 * d0 + d1 * d1;
        r += (double)b.x * b.x + (double)b.y * b.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { e += __shfl_xor_sync(ODPD_FULL, e, o); r += __shfl_xor_sync(ODPD_FULL, r, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = e; red[1][threadIdx.x >> 5] = r; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double se = 0.0, sr = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { se += red[0][w]; sr += red[1][w]; }
        out[2 * blockIdx.x] = se; out[2 * blockIdx.x + 1] = sr;
    }
}

// out[row][seg][i] = | sum_n w[n] * (x[row][seg*hop + n] - mean) * exp(-2 pi i k n / nfft) |,  k = (i + nfft/2) mod nfft  (fftshift-ed),
// x = a - b (b optional), samples beyond N are zero (np.fft.fft(x, n) zero-padding), mean = the segment's mean when `detrend`,
// w = periodic Hann when `hann` else 1.      grid (ceil(nfft/256), nseg, S), dynamic smem = 2 * nfft * sizeof(double2)
__global__ void __launch_bounds__(256) dft_mag_kernel(const float2 *__restrict__ a, const float2 *__restrict__ b, int N, int nfft, int nseg,
                                                     int hop, int hann, int detrend, double *__restrict__ out) {
    extern __shared__ double2 sm[];
    double2 *seg = sm, *tw = sm + nfft;
    __shared__ double red[2][8];
    const int row = blockIdx.z, s = blockIdx.y, t0 = s * hop;
    const float2 *pa = a + (size_t)row * N, *pb = b ? b + (size_t)row * N : nullptr;
    double mr = 0.0, mi = 0.0;
    for (int n = threadIdx.x; n < nfft; n += blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        if (t0 + n < N) {
            const float2 x = __ldg(pa + t0 + n);
            v.x = x.x; v.y = x.y;
            if (pb) { const float2 y = __ldg(pb + t0 + n); v.x -= (double)y.x; v.y -= (double)y.y; }
        }
        seg[n] = v;
        mr += v.x; mi += v.y;
        double sn, cs;
        sincospi(2.0 * (double)n / (double)nfft, &sn, &cs);
        tw[n] = make_double2(cs, -sn);
    }
    if (detrend) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mr += __shfl_xor_sync(ODPD_FULL, mr, o); mi += __shfl_xor_sync(ODPD_FULL, mi, o); }
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = mr; red[1][threadIdx.x >> 5] = mi; }
    }
    __syncthreads();
    if (detrend || hann) {
        double sr = 0.0, si = 0.0;
        if (detrend) {
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { sr += red[0][w]; si += red[1][w]; }
            sr /= (double)nfft; si /= (double)nfft;
        }
        for (int n = threadIdx.x; n < nfft; n += blockDim.x) {
            const double w = hann ? 0.5 - 0.5 * tw[n].x : 1.0;      // periodic Hann: 0.5 - 0.5 cos(2 pi n / nfft)
            seg[n] = make_double2((seg[n].x - sr) * w, (seg[n].y - si) * w);
        }
        __syncthreads();
    }
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nfft) return;
    double re0 = 0.0, im0 = 0.0, re1 = 0.0, im1 = 0.0;
    int idx = 0;
    for (int n = 0; n + 1 < nfft; n += 2) {
        const double2 x0 = seg[n], t0w = tw[idx];
        idx += k; if (idx >= nfft) idx -= nfft;
        const double2 x1 = seg[n + 1], t1w = tw[idx];
        idx += k; if (idx >= nfft) idx -= nfft;
        re0 = fma(x0.x, t0w.x, fma(-x0.y, t0w.y, re0)); im0 = fma(x0.x, t0w.y, fma(x0.y, t0w.x, im0));
        re1 = fma(x1.x, t1w.x, fma(-x1.y, t1w.y, re1)); im1 = fma(x1.x, t1w.y, fma(x1.y, t1w.x, im1));
    }
    if (nfft & 1) {
        const double2 x0 = seg[nfft - 1], t0w = tw[idx];
        re0 = fma(x0.x, t0w.x, fma(-x0.y, t0w.y, re0)); im0 = fma(x0.x, t0w.y, fma(x0.y, t0w.x, im0));
    }
    const double re = re0 + re1, im = im0 + im1;
    const int half = (nfft + 1) / 2;                    // np.fft.fftshift: shifted[i] = natural[(i + ceil(n/2)) % n]
    const int i = (k - half + nfft) % nfft;
    out[((size_t)row * nseg + s) * nfft + i] = sqrt(re * re + im * im);
}

}  // namespace odpd

using namespace odpd;

extern "C" {

int odpd_nmse_sums(const float *pred, const float *truth, int32_t S, int32_t N, double *out, void *stream) {
    ODPD_CHECK(pred && truth && out, "odpd_nmse_sums: NULL buffer");
    ODPD_CHECK(S >= 0 && N >= 1, "odpd_nmse_sums: bad sizes (%d,%d)", S, N);
    if (S == 0) return 0;
    nmse_sums_kernel<<<S, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(pred), reinterpret_cast<const float2 *>(truth), N, out);
    return check_launch("nmse_sums_kernel");
}

int odpd_dft_magnitude(const float *a, const float *b, int32_t S, int32_t N, int32_t nfft, int32_t nseg, int32_t hop, int32_t hann,
                       int32_t detrend, double *out, void *stream) {
    ODPD_CHECK(a && out, "odpd_dft_magnitude: NULL buffer");
    ODPD_CHECK(S >= 0 && N >= 1 && nfft >= 1 && nfft <= 6144 && nseg >= 1 && hop >= 0, "odpd_dft_magnitude: bad sizes (S=%d N=%d nfft=%d nseg=%d hop=%d)",
               S, N, nfft, nseg, hop);
    if (S == 0) return 0;
    const size_t smem = (size_t)2 * nfft * sizeof(double2);
    static bool attr[32] = {false};        // cudaFuncSetAttribute is per device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) dev = 0;
    if (!attr[dev]) { cudaFuncSetAttribute(dft_mag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6144 * 2 * (int)sizeof(double2)); attr[dev] = true; }
    dim3 grid((unsigned)((nfft + 255) / 256), (unsigned)nseg, (unsigned)S);
    dft_mag_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(a), reinterpret_cast<const float2 *>(b), N, nfft, nseg, hop,
                                                             hann, detrend, out);
    return check_launch("dft_mag_kernel");
}

}  // extern "C"

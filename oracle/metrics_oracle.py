"""CPU restatement (numpy only, float64) of the reference's evaluation metrics — TEST INFRASTRUCTURE, never imported by opendpd_b200.

Follows /root/reference/utils/metrics.py: NMSE :42-53, EVM :56-111 (magnitude_spectrum :8-40), ACLR :114-155, power_spectrum
:158-190 (scipy.signal.welch with its defaults: periodic Hann window, constant detrend, noverlap = nperseg // 2,
scaling='spectrum', two-sided).  scipy is deliberately NOT used here: Welch is restated from its definition so that the CUDA path
is checked against an independent implementation; tests/golden/metrics_*.npz (made by oracle/make_metrics_golden.py from the
unmodified reference, which does call scipy) pin this file."""
import numpy as np


def _band_indices(n_freq, fs, bw_main_ch, n_sub_ch):
    """index_left / index_right / sub-channel length on the fftshift-ed frequency axis (metrics.py:82-86, :138-142)."""
    freq = np.fft.fftshift(np.fft.fftfreq(n_freq, d=1.0 / fs))
    index_left = int(np.min(np.where(freq >= -bw_main_ch / 2)))
    index_right = int(np.max(np.where(freq <= bw_main_ch / 2)))
    return index_left, index_right, int((index_right - index_left) / n_sub_ch)


def nmse(prediction, ground_truth):
    p, g = np.asarray(prediction, np.float64), np.asarray(ground_truth, np.float64)
    mse = np.mean((g[..., 0] - p[..., 0]) ** 2 + (g[..., 1] - p[..., 1]) ** 2, axis=-1)
    energy = np.mean(g[..., 0] ** 2 + g[..., 1] ** 2, axis=-1)
    return float(np.mean(10 * np.log10(mse / energy)))


def _dft(x, n):
    """n-point DFT of the first n samples (zero-padded if shorter) of every row, fftshift-ed — np.fft.fft(x, n=n) semantics."""
    x = np.asarray(x, np.complex128)
    if x.shape[-1] < n:
        x = np.concatenate([x, np.zeros(x.shape[:-1] + (n - x.shape[-1],), np.complex128)], -1)
    k = np.arange(n)
    w = np.exp(-2j * np.pi * np.outer(k, k) / n)              # restated from the definition (O(n^2), test sizes only)
    return np.fft.fftshift(x[..., :n] @ w.T, axes=-1)


def evm(prediction, ground_truth, sample_rate=int(800e6), bw_main_ch=200e6, n_sub_ch=10, nperseg=2560):
    p = np.asarray(prediction, np.float64)
    g = np.asarray(ground_truth, np.float64)
    pc, gc = p[..., 0] + 1j * p[..., 1], g[..., 0] + 1j * g[..., 1]
    sp, sg = _dft(pc, nperseg), _dft(gc, nperseg)
    il, ir, ln = _band_indices(pc.shape[1], sample_rate, bw_main_ch, n_sub_ch)    # frequency axis has the SIGNAL length (:36)
    err = np.zeros((p.shape[0], n_sub_ch))
    for c in range(n_sub_ch):
        sl = slice(il + c * ln, il + (c + 1) * ln)
        err[:, c] = np.mean(np.abs(sp[:, sl] - sg[:, sl]), axis=-1) / np.mean(np.abs(sg[:, sl]), axis=-1)
    return float(20 * np.log10(np.mean(err.mean(axis=-1))))


def welch_spectrum(x, nperseg):
    """Two-sided Welch 'spectrum' of every row, scipy defaults (hann, constant detrend, 50 % overlap), natural bin order."""
    x = np.asarray(x, np.complex128)
    n = np.arange(nperseg)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / nperseg)                      # get_window('hann', nperseg) (periodic)
    step = nperseg - nperseg // 2
    starts = range(0, x.shape[-1] - nperseg + 1, step)
    k = np.arange(nperseg)
    w = np.exp(-2j * np.pi * np.outer(k, k) / nperseg)
    acc = np.zeros(x.shape[:-1] + (nperseg,))
    for s in starts:
        seg = x[..., s:s + nperseg]
        seg = (seg - seg.mean(axis=-1, keepdims=True)) * win
        acc += np.abs(seg @ w.T) ** 2
    return acc / len(starts) / win.sum() ** 2


def aclr(prediction, fs=800e6, nperseg=2560, bw_main_ch=200e6, n_sub_ch=10):
    p = np.asarray(prediction, np.float64)
    ps = welch_spectrum(p[..., 0] + 1j * p[..., 1], nperseg)
    half = int(nperseg / 2)
    ps = np.concatenate((ps[..., half:], ps[..., :half]), axis=-1).mean(axis=0)
    f = np.fft.fftfreq(nperseg, d=1.0 / fs)
    freq = np.concatenate((f[half:], f[:half]))
    il = int(np.min(np.where(freq >= -bw_main_ch / 2)))
    ir = int(np.max(np.where(freq <= bw_main_ch / 2)))
    ln = int((ir - il) / n_sub_ch)
    sub = np.array([ps[il + c * ln:il + (c + 1) * ln].sum() for c in range(n_sub_ch)])
    mx = sub.max()
    return float(10 * np.log10(ps[il - ln:il].sum() / mx)), float(10 * np.log10(ps[ir:ir + ln].sum() / mx))

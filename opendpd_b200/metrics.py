"""Evaluation metrics on the device — drop-in for the reference's utils/metrics.py NMSE / EVM / ACLR and
modules/train_funcs.py:93-105 `calculate_metrics`.  Same signatures, defaults and quirks (EVM's frequency axis
has the signal length and its sample rate defaults to 800e6 whatever the dataset says, metrics.py:36,56; ACLR averages the Welch
spectrum over the segments before the band sums, :187).  The reductions over samples (NMSE) and the windowed DFTs (EVM, ACLR) run in
libodpd.so (csrc/metrics.cu); the band bookkeeping — a few dozen scalars — is spelled exactly as the reference spells it."""
import ctypes
import math
import numpy as np
import torch

from . import _ffi
from .functional import _ptr, _stream


def _dev(t):
    """(S,N,2) float32 CUDA tensor from a numpy array / CPU tensor / CUDA tensor (net_eval hands numpy arrays over, train_funcs.py:86-87)."""
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    if not torch.cuda.is_available():
        raise _ffi.OdpdError("opendpd_b200.metrics runs on a CUDA device only (no CPU fallback); use the reference's utils/metrics.py on CPU")
    t = t.to(device="cuda", dtype=torch.float32).contiguous()
    if t.dim() != 3 or t.size(-1) != 2:
        raise _ffi.OdpdError(f"expected (segments, samples, 2) I/Q data, got {tuple(t.shape)}")
    return t


def _dft_mag(a, b, nfft, nseg=1, hop=0, hann=False, detrend=False):
    S, N = a.shape[0], a.shape[1]
    out = torch.empty((S, nseg, nfft), dtype=torch.float64, device=a.device)
    _ffi.check(_ffi.lib().odpd_dft_magnitude(_ptr(a), _ptr(b), S, N, int(nfft), int(nseg), int(hop), int(bool(hann)), int(bool(detrend)),
                                             _ptr(out), _stream()))
    return out


def _band(n_freq, fs, bw_main_ch, n_sub_ch):
    freq = np.fft.fftshift(np.fft.fftfreq(n_freq, d=1 / fs))
    index_left = int(np.min(np.where(freq >= -bw_main_ch / 2)))
    index_right = int(np.max(np.where(freq <= bw_main_ch / 2)))
    return index_left, index_right, int((index_right - index_left) / n_sub_ch)


def NMSE(prediction, ground_truth):
    """utils/metrics.py:42-53."""
    p, g = _dev(prediction), _dev(ground_truth)
    sums = torch.empty((p.shape[0], 2), dtype=torch.float64, device=p.device)
    _ffi.check(_ffi.lib().odpd_nmse_sums(_ptr(p), _ptr(g), p.shape[0], p.shape[1], _ptr(sums), _stream()))
    return float((10 * torch.log10(sums[:, 0] / sums[:, 1])).mean().item())


def EVM(prediction, ground_truth, sample_rate=int(800e6), bw_main_ch=200e6, n_sub_ch=10, nperseg=2560):
    """utils/metrics.py:56-111: |FFT_nperseg(pred) - FFT_nperseg(truth)| = |FFT_nperseg(pred - truth)| by linearity."""
    p, g = _dev(prediction), _dev(ground_truth)
    d_mag = _dft_mag(p, g, nperseg)[:, 0]
    g_mag = _dft_mag(g, None, nperseg)[:, 0]
    il, ir, ln = _band(p.shape[1], sample_rate, bw_main_ch, n_sub_ch)
    err = torch.stack([d_mag[:, il + c * ln:il + (c + 1) * ln].mean(-1) / g_mag[:, il + c * ln:il + (c + 1) * ln].mean(-1)
                       for c in range(n_sub_ch)], dim=-1)
    return float(20 * math.log10(err.mean(dim=-1).mean().item()))


def ACLR(prediction, fs=800e6, nperseg=2560, bw_main_ch=200e6, n_sub_ch=10):
    """utils/metrics.py:114-190 (scipy.signal.welch defaults: periodic Hann, constant detrend, noverlap = nperseg // 2, 'spectrum')."""
    p = _dev(prediction)
    if nperseg % 2:
        raise _ffi.OdpdError("ACLR: nperseg must be even (the reference's half-spectrum swap, metrics.py:181-185)")
    N = p.shape[1]
    if N < nperseg:
        raise _ffi.OdpdError(f"ACLR: segments of {N} samples are shorter than nperseg={nperseg}")
    step = nperseg - nperseg // 2
    nseg = (N - nperseg) // step + 1
    mag = _dft_mag(p, None, nperseg, nseg=nseg, hop=step, hann=True, detrend=True)
    ps = (mag * mag).mean(dim=1) / (nperseg / 2.0) ** 2          # sum of the periodic Hann window = nperseg / 2
    ps = ps.mean(dim=0)
    il, ir, ln = _band(nperseg, fs, bw_main_ch, n_sub_ch)
    sub = torch.stack([ps[il + c * ln:il + (c + 1) * ln].sum() for c in range(n_sub_ch)])
    mx = sub.max()
    left, right = ps[il - ln:il].sum(), ps[ir:ir + ln].sum()
    return float(10 * torch.log10(left / mx).item()), float(10 * torch.log10(right / mx).item())


def calculate_metrics(args, stat, prediction, ground_truth):
    """Drop-in for modules/train_funcs.py:93-105."""
    stat["NMSE"] = NMSE(prediction, ground_truth)
    stat["EVM"] = EVM(prediction, ground_truth, bw_main_ch=args.bw_main_ch, n_sub_ch=args.n_sub_ch, nperseg=args.nperseg)
    left, right = ACLR(prediction, fs=args.input_signal_fs, nperseg=args.nperseg, bw_main_ch=args.bw_main_ch, n_sub_ch=args.n_sub_ch)
    stat["ACLR_L"], stat["ACLR_R"] = float(np.mean([left])), float(np.mean([right]))
    stat["ACLR_AVG"] = (stat["ACLR_L"] + stat["ACLR_R"]) / 2
    return stat

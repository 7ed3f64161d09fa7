"""Golden vectors for the evaluation metrics, produced by the UNMODIFIED reference (utils/metrics.py) in this container:
    PYTHONDONTWRITEBYTECODE=1 python oracle/make_metrics_golden.py
writes tests/golden/metrics_<case>.npz (inputs as float32 — what net_eval hands to calculate_metrics — and the reference's outputs)."""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
from utils import metrics as ref      # noqa: E402


def ofdm_like(rng, S, N, fs, bw, regrowth, noise):
    """Band-limited multicarrier 'PA output' with third-order spectral regrowth, plus the linear target."""
    k = np.fft.fftfreq(N, d=1.0 / fs)
    occ = np.abs(k) <= 0.45 * bw
    X = np.zeros((S, N), np.complex128)
    X[:, occ] = (rng.standard_normal((S, occ.sum())) + 1j * rng.standard_normal((S, occ.sum())))
    x = np.fft.ifft(X, axis=-1)
    x /= np.abs(x).max()
    y = x * (1 - regrowth * np.abs(x) ** 2) + noise * (rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape))
    to_iq = lambda z: np.stack([z.real, z.imag], -1).astype(np.float32)
    return to_iq(y), to_iq(x)


def main():
    rng = np.random.default_rng(7)
    cases = {
        "apa": dict(S=8, N=2560, fs=983.04e6, bw=200e6, n_sub=1, regrowth=0.15, noise=1e-4),      # datasets/APA_200MHz/spec.json
        "dpa": dict(S=3, N=2560, fs=800e6, bw=200e6, n_sub=10, regrowth=0.05, noise=1e-5),        # the functions' defaults
        "small": dict(S=2, N=512, fs=800e6, bw=160e6, n_sub=4, regrowth=0.3, noise=1e-3),
    }
    for name, c in cases.items():
        pred, truth = ofdm_like(rng, c["S"], c["N"], c["fs"], c["bw"], c["regrowth"], c["noise"])
        out = dict(pred=pred, truth=truth, fs=c["fs"], bw=c["bw"], n_sub=c["n_sub"], nperseg=c["N"],
                   nmse=ref.NMSE(pred, truth),
                   evm=ref.EVM(pred, truth, bw_main_ch=c["bw"], n_sub_ch=c["n_sub"], nperseg=c["N"]),        # sample_rate stays the default (train_funcs.py:95)
                   aclr=np.array(ref.ACLR(pred, fs=c["fs"], nperseg=c["N"], bw_main_ch=c["bw"], n_sub_ch=c["n_sub"])))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"metrics_{name}.npz"), **out)
        print(name, out["nmse"], out["evm"], out["aclr"])


if __name__ == "__main__":
    main()

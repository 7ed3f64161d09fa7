"""CPU: the numpy restatement of the reference's evaluation metrics (oracle/metrics_oracle.py) against golden vectors produced by the
unmodified reference (oracle/make_metrics_golden.py -> tests/golden/metrics_*.npz) and against scipy.signal.welch itself."""
import glob
import os
import numpy as np
import pytest

from oracle import metrics_oracle as mo
from tests.util import GOLDEN

CASES = sorted(os.path.basename(f)[8:-4] for f in glob.glob(os.path.join(GOLDEN, "metrics_*.npz")))


def load(name):
    g = np.load(os.path.join(GOLDEN, f"metrics_{name}.npz"))
    kw = dict(bw_main_ch=float(g["bw"]), n_sub_ch=int(g["n_sub"]), nperseg=int(g["nperseg"]))
    return g, kw


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_the_reference_metrics(name):
    g, kw = load(name)
    # the reference evaluates in float32 / complex64 (net_eval hands float32 arrays over): agreement to ~1e-6 dB
    assert abs(mo.nmse(g["pred"], g["truth"]) - float(g["nmse"])) < 1e-4
    assert abs(mo.evm(g["pred"], g["truth"], **kw) - float(g["evm"])) < 1e-4
    assert np.abs(np.array(mo.aclr(g["pred"], fs=float(g["fs"]), **kw)) - g["aclr"]).max() < 1e-4


def test_welch_restatement_matches_scipy_with_overlapping_segments():
    from scipy.signal import welch
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 4000)) + 1j * rng.standard_normal((3, 4000))
    _, ps = welch(x, fs=800e6, nperseg=512, return_onesided=False, scaling="spectrum", axis=-1)
    mine = mo.welch_spectrum(x, 512)
    assert np.abs(mine - ps).max() <= 1e-12 * ps.max()

"""GPU: opendpd_b200.metrics (csrc/metrics.cu through the C ABI) against the reference's golden vectors and the CPU oracle."""
import numpy as np
import pytest
import torch

from tests.test_metrics_oracle import CASES, load

pytestmark = [pytest.mark.gpu, pytest.mark.filterwarnings("ignore:Mean of empty slice"), pytest.mark.filterwarnings("ignore:invalid value encountered")]
TOL_DB = 1e-4          # dB; the reference itself works in float32 / complex64


@pytest.mark.parametrize("name", CASES)
def test_metrics_match_the_reference_goldens(name):
    from opendpd_b200 import metrics as M
    g, kw = load(name)
    assert abs(M.NMSE(g["pred"], g["truth"]) - float(g["nmse"])) < TOL_DB
    assert abs(M.EVM(g["pred"], g["truth"], **kw) - float(g["evm"])) < TOL_DB
    assert np.abs(np.array(M.ACLR(g["pred"], fs=float(g["fs"]), **kw)) - g["aclr"]).max() < TOL_DB


@pytest.mark.parametrize("S,N,nperseg", [(5, 1000, 256), (2, 300, 300), (3, 2560, 1024), (1, 700, 64)])
def test_metrics_match_the_oracle_on_seeded_segments(S, N, nperseg):
    """Overlapping Welch segments (N > nperseg), truncated EVM transforms, odd transform lengths."""
    from opendpd_b200 import metrics as M
    from oracle import metrics_oracle as mo
    rng = np.random.default_rng(S * 1000 + N)
    truth = (0.3 * rng.standard_normal((S, N, 2))).astype(np.float32)
    pred = (truth * (1 - 0.2 * (truth ** 2).sum(-1, keepdims=True)) + 1e-3 * rng.standard_normal((S, N, 2))).astype(np.float32)
    kw = dict(bw_main_ch=200e6, n_sub_ch=4, nperseg=nperseg)
    assert abs(M.NMSE(pred, truth) - mo.nmse(pred, truth)) < 1e-9
    if N == nperseg or True:
        a, b = M.EVM(torch.from_numpy(pred).cuda(), torch.from_numpy(truth).cuda(), **kw), mo.evm(pred, truth, **kw)
        assert (np.isnan(a) and np.isnan(b)) or abs(a - b) < 1e-8, (a, b)
    if nperseg % 2 == 0:
        assert np.abs(np.array(M.ACLR(pred, fs=800e6, **kw)) - np.array(mo.aclr(pred, fs=800e6, **kw))).max() < 1e-8


def test_calculate_metrics_is_a_drop_in():
    from opendpd_b200 import metrics as M
    g, kw = load("apa")

    class Args:
        bw_main_ch, n_sub_ch, nperseg, input_signal_fs = kw["bw_main_ch"], kw["n_sub_ch"], kw["nperseg"], float(g["fs"])
    stat = M.calculate_metrics(Args, {}, g["pred"], g["truth"])
    assert set(stat) == {"NMSE", "EVM", "ACLR_L", "ACLR_R", "ACLR_AVG"}
    assert abs(stat["ACLR_AVG"] - float(g["aclr"].mean())) < TOL_DB and abs(stat["NMSE"] - float(g["nmse"])) < TOL_DB

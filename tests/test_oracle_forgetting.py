"""The premise of the time-chunked kernels, checked on the CPU oracle alone: a gated cell started from the zero
state Wu steps before a chunk boundary produces, from the boundary on, the same outputs as the full-history run — to well inside the
north-star tolerance — for the freshly initialised reference models.  (At run time the CUDA path does not rely on this: it verifies
every boundary and re-runs failing sequences serially; this test documents why the verify pass normally passes.)"""
import numpy as np
import pytest
import torch

from oracle import oracle
from opendpd_b200 import models


@pytest.mark.parametrize("kind,H,warm", [("gru", 16, 128), ("dgru", 13, 128), ("dgru", 13, 64), ("qgru", 10, 128), ("lstm", 9, 128),
                                         ("pgjanet", 15, 256), ("dvrjanet", 15, 256)])
def test_zero_state_warmup_reproduces_the_full_history_outputs(kind, H, warm):
    torch.manual_seed(0)
    net = models.CoreModel(2, H, 1, kind, num_dvr_units=3)
    params = np.concatenate([p.detach().numpy().ravel() for _, p in net.backbone.named_parameters()])
    g = torch.Generator().manual_seed(3)
    x = (0.2 * torch.randn(4, 1024, 2, generator=g)).clamp(-0.7, 0.7).numpy()
    full = oracle.run(kind, x, params, H=H, K=3, dtype=np.float64, want_grads=False)["out"]
    scale = np.abs(full).max()
    for s in (384, 768):
        part = oracle.run(kind, x[:, s - warm:s + 128], params, H=H, K=3, dtype=np.float64, want_grads=False)["out"][:, warm:]
        err = np.abs(part - full[:, s:s + 128]).max() / scale
        assert err < 2.0 ** -18, (kind, s, err)      # the verify pass's tolerance, here on the outputs in fp64 (no rounding-noise floor)
    # a warm-up that is far too short does NOT reproduce them: the verify pass has something to catch
    short = oracle.run(kind, x[:, 768 - 4:768 + 8], params, H=H, K=3, dtype=np.float64, want_grads=False)["out"][:, 4:]
    assert np.abs(short - full[:, 768:776]).max() / scale > 2.0 ** -18

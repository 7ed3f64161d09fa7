"""The reference initialises APNRRU's Z to zeros (apnrru.py:19), so at init — and therefore in the goldens — the two dense layers of the RRU
cell get a zero gradient.  This pins the C oracle's hand-derived backward through them: random parameters, fp64, against autograd of the
PyTorch-op restatement (oracle/torch_port.py, itself pinned to the reference goldens by tests/test_torch_port.py)."""
import numpy as np
import pytest
import torch
from oracle import oracle, torch_port
from tests.util import rel_err


@pytest.mark.parametrize("H,B,T", [(8, 3, 40), (14, 2, 33), (3, 2, 16)])
def test_apnrru_oracle_matches_autograd_with_random_parameters(H, B, T):
    g = torch.Generator().manual_seed(100 + H)
    P = oracle.n_params("apnrru", H)
    flat = (0.3 * torch.randn(P, generator=g, dtype=torch.float64)).requires_grad_(True)
    x = (0.3 * torch.randn(B, T, 2, generator=g, dtype=torch.float64)).requires_grad_(True)
    y = 0.5 * torch.randn(B, T, 2, generator=g, dtype=torch.float64)
    out = torch_port.forward("apnrru", x, flat, H)
    loss = torch.nn.MSELoss()(out, y)
    loss.backward()
    r = oracle.run("apnrru", x.detach().numpy(), flat.detach().numpy(), target=y.numpy(), H=H, dtype=np.float64)
    assert rel_err(r["out"], out.detach().numpy()) < 1e-11
    assert abs(r["loss"] - loss.item()) <= 1e-12 * abs(loss.item())
    assert rel_err(r["gx"], x.grad.numpy()) < 1e-10
    assert rel_err(r["gparams"], flat.grad.numpy()) < 1e-10

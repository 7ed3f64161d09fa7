"""Pins oracle/torch_port.py (the PyTorch-op restatement used as the reference-like CPU arm of bench.py) to the
golden vectors of the unmodified reference."""
import numpy as np
import pytest
import torch
from tests.util import golden_cases, load_golden, rel_err, tol_for
from oracle import torch_port


@pytest.mark.parametrize("name", [c for c in golden_cases() if "qat" not in c])
def test_torch_port_matches_reference(name):
    g = load_golden(name)
    torch.set_num_threads(4)
    flat = torch.tensor(g["params"], requires_grad=True)
    x = torch.tensor(g["x"], requires_grad=True)
    out = torch_port.forward(g["kind"], x, flat, g["H"], g["K"], g["thx"], g["thh"])
    loss = torch.nn.MSELoss()(out, torch.tensor(g["y"]))
    loss.backward()
    assert rel_err(out.detach().numpy(), g["out"]) < tol_for(g, "out")
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel_err(x.grad.numpy(), g["gx"]) < tol_for(g, "gx")
    assert rel_err(flat.grad.numpy(), g["gparams"]) < tol_for(g, "gparams")


@pytest.mark.parametrize("name", __import__("tests.util", fromlist=["wide_cases"]).wide_cases())
def test_torch_port_layers_match_reference(name):
    """forward_layers (hidden sizes above 32, num_layers > 1) against goldens of the unmodified reference — it is the fp64 arbiter of
    tests/test_gpu_wide.py's seeded cases."""
    g = load_golden(name)
    torch.set_num_threads(4)
    flat = torch.tensor(g["params"], requires_grad=True)
    x = torch.tensor(g["x"], requires_grad=True)
    out = torch_port.forward_layers(g["kind"], x, flat, g["H"], g["L"])
    loss = torch.nn.MSELoss()(out, torch.tensor(g["y"]))
    loss.backward()
    assert rel_err(out.detach().numpy(), g["out"]) < tol_for(g, "out")
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel_err(x.grad.numpy(), g["gx"]) < tol_for(g, "gx")
    assert rel_err(flat.grad.numpy(), g["gparams"]) < tol_for(g, "gparams")

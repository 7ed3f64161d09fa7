"""CPU: oracle/next_cells.py (restatements prepared ahead of their CUDA kernels) against goldens from the unmodified
reference (oracle/make_next_golden.py).  There is no CUDA path for these cells yet; this pins the checker the next round will use."""
import glob
import os
import numpy as np
import pytest

from oracle import next_cells
from tests.util import GOLDEN

CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "next_vdlstm_*.npz")))


@pytest.mark.parametrize("name", CASES)
def test_vdlstm_oracle_matches_the_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    r = next_cells.vdlstm(g["x"], g["params"], int(g["H"]), target=g["y"])
    rel = lambda a, b: float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))
    assert rel(r["out"], g["out"]) < 1e-12
    assert abs(r["loss"] - float(g["loss"])) < 1e-13 * float(g["loss"]) + 1e-16
    assert rel(r["gx"], g["gx"]) < 1e-10
    assert rel(r["gparams"], g["gparams"]) < 1e-10
    assert list(g["names"]) == ["rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0", "fc_lambda_1.weight",
                                "fc_lambda_1.bias", "fc_lambda_2.weight", "fc_lambda_2.bias", "fc_out.weight", "fc_out.bias"]

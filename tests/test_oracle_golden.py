"""Pins the CPU oracle (oracle/odpd_oracle.c) against golden vectors produced by the unmodified reference
(oracle/make_golden.py).  fp64 oracle vs fp64 reference must agree to ~1e-12 (same algorithm), fp32 to 1e-5."""
import numpy as np
import pytest
from tests.util import golden_cases, load_golden, rel_err, tol_for
from oracle import oracle


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_fp64_matches_reference(name):
    g = load_golden(name)
    r = oracle.run(g["kind"], g["x"], g["params"], target=g["y"], H=g["H"], K=g["K"], thx=g["thx"], thh=g["thh"],
                   dtype=np.float64, want_masks=True)
    assert rel_err(r["out"], g["out64"]) < 1e-11
    assert abs(r["loss"] - float(g["loss64"])) <= 1e-12 * abs(float(g["loss64"]))
    assert rel_err(r["gx"], g["gx64"]) < 1e-10
    assert rel_err(r["gparams"], g["gparams64"]) < 1e-10
    if "mask_x64" in g:
        assert np.array_equal(r["mask_x"], g["mask_x64"])
        assert np.array_equal(r["mask_h"], g["mask_h64"])
        assert r["stats"].tolist() == [int(v) for v in g["stats64"]]


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_fp32_matches_reference(name):
    g = load_golden(name)
    r = oracle.run(g["kind"], g["x"], g["params"], target=g["y"], H=g["H"], K=g["K"], thx=g["thx"], thh=g["thh"],
                   dtype=np.float32, want_masks=True)
    assert rel_err(r["out"], g["out"]) < tol_for(g, "out")
    assert abs(r["loss"] - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel_err(r["gx"], g["gx"]) < tol_for(g, "gx")
    assert rel_err(r["gparams"], g["gparams"]) < tol_for(g, "gparams")
    if "mask_x" in g:
        # delta-x masks depend only on separately-rounded feature arithmetic: bit exact (SURVEY §7 hard part 1)
        assert np.array_equal(r["mask_x"], g["mask_x"])
        assert int((r["mask_h"] != g["mask_h"]).sum()) == 0
        assert r["stats"].tolist() == [int(v) for v in g["stats"]]


def test_threads_do_not_change_result():
    g = load_golden("dgru_h13_b8_t256")
    a = oracle.run("dgru", g["x"], g["params"], target=g["y"], H=13, dtype=np.float64, nthreads=1)
    b = oracle.run("dgru", g["x"], g["params"], target=g["y"], H=13, dtype=np.float64, nthreads=4)
    assert np.array_equal(a["out"], b["out"]) and rel_err(a["gparams"], b["gparams"]) < 1e-13


def test_external_gout_equals_mse_path():
    g = load_golden("gru_h23_b3_t17")
    a = oracle.run("gru", g["x"], g["params"], target=g["y"], H=23, dtype=np.float64)
    gout = 2.0 * (a["out"] - g["y"].astype(np.float64)) / a["out"].size
    b = oracle.run("gru", g["x"], g["params"], gout=gout, H=23, dtype=np.float64)
    assert rel_err(b["gparams"], a["gparams"]) < 1e-13 and rel_err(b["gx"], a["gx"]) < 1e-13

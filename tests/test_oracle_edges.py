"""The C oracle against the UNMODIFIED reference at the edges of every backbone's frame-length domain (authoring container only:
needs /root/reference; the GPU-side counterpart is tests/test_gpu_edges.py, which trusts the oracle at exactly these lengths)."""
import json, os, subprocess, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/backbones"), reason="reference checkout not present (GPU box)")
def test_oracle_equals_reference_at_short_and_block_edge_frames():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "check_edges.py")], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    r = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert r["cases"] == 183
    assert max(r["worst_rel_err"].values()) < 1e-11, r["worst_rel_err"]
    # the windowed backbones reject frames shorter than their window minus one (their padding is a slice of the frame itself);
    # the library raises for the same lengths (tests/test_gpu_edges.py::test_frames_the_reference_rejects_raise)
    rej = r["reference_rejects"]
    assert rej["rvtdcnn"] == [1, 2] and rej["mcldnn"] == [1, 2, 3] and rej["bojanet"] == [1, 2, 3, 4, 5] and rej["apnrru"] == [1, 2, 3, 4, 5]
    assert all(v == [] for k, v in rej.items() if k not in ("rvtdcnn", "mcldnn", "bojanet", "apnrru"))

import glob, json, os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_cases(kinds=None):
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        name = os.path.basename(f)[:-4]
        if name.startswith(("metrics_", "next_", "iq_streams", "full_", "wide_")):   # evaluation-metric goldens (tests/test_metrics_oracle.py); full-size and wide / stacked cases have their own tests
            continue
        if kinds is None or any(name.startswith(k) for k in kinds):
            out.append(name)
    return out


def wide_cases():
    """Goldens of the layered path (csrc/wide.cu): hidden sizes 33..64 and / or num_layers > 1, from the unmodified reference."""
    return [os.path.basename(f)[:-4] for f in sorted(glob.glob(os.path.join(GOLDEN, "wide_*.npz")))]


RNN_KINDS = ("gru", "lstm", "dgru", "qgru", "qgru_amp1")


def dims_K(g):
    """OdpdDims.K for a golden case: num_layers for the nn.GRU / nn.LSTM based backbones (0 = one layer), else the stored K
    (DVRJANET units, QAT bit widths)."""
    if g["kind"] in RNN_KINDS:
        return g["L"] if g["L"] > 1 else 0
    return g["K"]


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: d[k] for k in d.files}
    g["kind"] = str(g["kind"]); g["H"] = int(g["H"]); g["K"] = int(g["K"]); g["L"] = int(g["L"]) if "L" in g else 1
    g["thx"] = float(g["thx"]); g["thh"] = float(g["thh"])
    g["param_index"] = json.loads(str(g["param_index"]))
    return g


def rel_err(a, b):
    """max|a-b| / max|b| — the relative-fp32 figure of merit used throughout (north_star: <=1e-5)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


ACHIEVED = os.path.join(ROOT, "gpurun_out", "parity_achieved.jsonl")


def note_achieved(what, **kw):
    """Append the error figures a parity assertion actually saw to gpurun_out/parity_achieved.jsonl (when that directory exists:
    GPU box runs), so tolerances can be judged against what is achieved, not only against what is allowed."""
    try:
        if os.path.isdir(os.path.dirname(ACHIEVED)):
            with open(ACHIEVED, "a") as f:
                f.write(json.dumps(dict(test=os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], what=what, **kw)) + "\n")
    except OSError:
        pass


def tol_for(g, key, base=1e-5):
    """Tolerance for comparing an fp32 implementation with the fp32 reference: the north-star 1e-5, widened only
    where the reference's own fp32-vs-fp64 disagreement shows the case is ill-conditioned (e.g. DVRJANET's
    cos/sin of an unbounded learned phase): 5x that disagreement."""
    cond = rel_err(g[key], g[key + "64"])
    return max(base, 5.0 * cond)


def assert_close(mine, ref, tol, what="", worst_factor=10.0):
    """Parity criterion for fp32 implementations of piecewise-smooth networks: the 99.99th percentile of
    |mine-ref|/max|ref| must be below `tol` and the WORST element below worst_factor*tol (10 by default).  The slack on isolated elements exists
    because ReLU / hardswish / delta-threshold kinks turn a 1-ulp difference of a pre-activation that sits at the kink
    into a finite jump of one gradient element (observed: 1 element in 262144 at 5.6e-5).  The achieved figures are logged."""
    a = np.asarray(mine, dtype=np.float64); b = np.asarray(ref, dtype=np.float64)
    e = np.abs(a - b) / (np.max(np.abs(b)) + 1e-300)
    q = float(np.quantile(e, 0.9999)) if e.size >= 10000 else float(e.max())
    worst = float(e.max())
    note_achieved(what, p9999=q, worst=worst, worst_index=int(e.argmax()), tol=tol, n=int(e.size))
    assert q < tol, f"{what}: p99.99 rel err {q:.3e} >= {tol:.1e} (worst {worst:.3e})"
    assert worst < worst_factor * tol, f"{what}: max rel err {worst:.3e} (element {int(e.argmax())}) >= {worst_factor * tol:.1e}"

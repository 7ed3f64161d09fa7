"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/odpd.h declares (no compute calls),
parameter counts / flat layout / state_dict names agree with the reference (via the golden fixtures), loud failure without CUDA."""
import ctypes
import os
import re
import numpy as np
import pytest
import torch

from tests.util import ROOT, dims_K, golden_cases, load_golden, wide_cases
from opendpd_b200 import _ffi, models


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "odpd.h")).read()
    declared = set(re.findall(r"\b(odpd_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_ffi.SYMBOLS), (declared ^ set(_ffi.SYMBOLS))
    L = _ffi.lib()
    for name in declared:
        assert hasattr(L, name)
    assert L.odpd_version() >= 100


def test_error_reporting_without_gpu_work():
    L = _ffi.lib()
    d = _ffi.OdpdDims(99, 1, 1, 8, 0, 0, 0.0, 0.0)
    assert L.odpd_saved_bytes(ctypes.byref(d)) < 0
    assert b"unknown cell" in L.odpd_last_error()
    d = _ffi.OdpdDims(_ffi.CELLS["gru"], 1, 1, 65, 0, 0, 0.0, 0.0)          # the layered path stops at 64
    assert L.odpd_saved_bytes(ctypes.byref(d)) < 0 and b"hidden_size" in L.odpd_last_error()
    d = _ffi.OdpdDims(_ffi.CELLS["pgjanet"], 1, 1, 64, 0, 0, 0.0, 0.0)      # cells without a layered path stop at the fused tiers
    assert L.odpd_saved_bytes(ctypes.byref(d)) < 0 and b"hidden_size" in L.odpd_last_error()
    d = _ffi.OdpdDims(_ffi.CELLS["lstm"], 1, 1, 16, 9, 0, 0.0, 0.0)         # K = num_layers for the nn.GRU / nn.LSTM backbones
    assert L.odpd_saved_bytes(ctypes.byref(d)) < 0 and b"num_layers" in L.odpd_last_error()


@pytest.mark.parametrize("name", golden_cases() + wide_cases())
def test_param_count_and_flat_layout_match_reference(name):
    g = load_golden(name)
    if g["kind"].endswith("_qat"):
        from opendpd_b200.quant import get_quant_model

        class _Proj:
            quant, n_bits_w, n_bits_a, pretrained_model = True, g["K"] & 255, (g["K"] >> 8) & 255, ""
        net = get_quant_model(_Proj(), models.CoreModel(2, g["H"], 1, g["kind"][:-4]))
    else:
        net = models.CoreModel(2, max(g["H"], 1), g["L"], g["kind"], num_dvr_units=g["K"], thx=g["thx"], thh=g["thh"])
    names = [n for n, _ in net.backbone.named_parameters()]
    assert names == [n for n, _ in g["param_index"]]
    assert [list(p.shape) for _, p in net.backbone.named_parameters()] == [s for _, s in g["param_index"]]
    n = sum(p.numel() for p in net.backbone.parameters())
    assert n == g["params"].size == _ffi.n_params(g["kind"], g["H"], dims_K(g))
    flat, layout = net.backbone._flat_sync()
    off = 0
    for (o, cnt, shape), (_, p) in zip(layout, net.backbone.named_parameters()):
        assert o == off and p.data_ptr() == flat.data_ptr() + 4 * off
        off += cnt
    with torch.no_grad():      # in-place optimiser-style updates go through to the flat buffer
        for p in net.backbone.parameters():
            p.add_(1.0)
    assert torch.equal(flat[:off], torch.cat([p.detach().reshape(-1) for p in net.backbone.parameters()]))


def test_reference_checkpoint_names():
    net = models.CoreModel(2, 15, 1, "deltagru_tcnskip", thx=0.01, thh=0.05)
    assert list(net.state_dict()) == ["backbone.rnn.x2h.weight", "backbone.rnn.h2h.weight", "backbone.fc_out.weight",
                                      "backbone.tcn.0.weight", "backbone.tcn.2.weight"]
    assert sum(p.numel() for p in net.parameters()) == 999          # README's "996" model, SURVEY §4
    assert sum(p.numel() for p in models.CoreModel(2, 13, 1, "dgru").parameters()) == 1041
    assert net.backbone.thx == 0.01 and net.backbone.thh == 0.05 and net.backbone.get_temporal_sparsity() == {}


def test_cpu_tensor_is_rejected_loudly():
    net = models.CoreModel(2, 8, 1, "gru")
    with pytest.raises(_ffi.OdpdError):
        net(torch.zeros(1, 4, 2))


def test_vdlstm_container_matches_reference_golden():
    """Row f-4, first cell: names, count and initial weights of the native VDLSTM equal the reference's (fixture from its own ctor)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "next_vdlstm_h8_b3_t40.npz"))
    torch.manual_seed(0)
    net = models.CoreModel(2, 8, 1, "vdlstm")
    assert [n for n, _ in net.backbone.named_parameters()] == list(g["names"])
    mine = np.concatenate([p.detach().numpy().ravel() for _, p in net.backbone.named_parameters()])
    assert mine.size == 538 == _ffi.n_params("vdlstm", 8) and np.array_equal(mine, g["params"].astype(np.float32))


def test_unsupported_configs_raise():
    assert models.CoreModel(2, 8, 2, "gru").backbone._spec().K == 2     # stacked layers: the layered path (csrc/wide.cu), K = num_layers
    with pytest.raises(NotImplementedError):
        models.CoreModel(2, 8, 9, "gru")            # more layers than the layered path holds
    with pytest.raises(NotImplementedError):
        models.CoreModel(2, 8, 2, "vdlstm")         # cells outside the layered path stay single-layer
    with pytest.raises(NotImplementedError):
        models.CoreModel(2, 8, 2, "deltagru")
    with pytest.raises(ValueError):
        models.CoreModel(2, 8, 1, "mamba")           # a name arguments.py:46 lists but models.py has no branch for


def test_dims_struct_matches_the_header_layout():
    """ctypes mirror of `OdpdDims` (include/odpd.h): ten 4-byte fields then two 8-byte device pointers, and a C compiler agrees."""
    import subprocess, tempfile
    assert ctypes.sizeof(_ffi.OdpdDims) == 56
    assert _ffi.OdpdDims.x_starts.offset == 40 and _ffi.OdpdDims.target_starts.offset == 48
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "odpd.h"\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(OdpdDims), '
           'offsetof(OdpdDims, tchunks), offsetof(OdpdDims, x_starts), offsetof(OdpdDims, target_starts)); return 0;}\n')
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o", exe])   # the header is plain C
        out = subprocess.check_output([exe]).decode().split()
    assert out == ["56", "32", "40", "48"], out


def test_iq_stream_is_the_reference_framing():
    """functional.IqStream(stream, starts, T).frames() == the stride-1 windows IQFrameDataset materialises
    (modules/data_collector.py:233-252: frame k = rows [k, k+T) of the stream)."""
    from opendpd_b200.functional import IqStream
    g = torch.Generator().manual_seed(0)
    stream = torch.randn(500, 2, generator=g)
    T = 37
    ref = torch.stack([stream[k:k + T] for k in range(500 - T + 1)])      # IQFrameDataset.__init__ (stride 1)
    starts = torch.tensor([0, 5, 463, 17, 17], dtype=torch.int32)
    v = IqStream(stream, starts, T)
    assert v.shape == (5, T, 2) and torch.equal(v.frames(), ref[starts.long()])
    with pytest.raises(_ffi.OdpdError):
        IqStream(stream, starts.long(), T)                                # indices must be int32
    with pytest.raises(_ffi.OdpdError):
        IqStream(stream.t(), starts, T)                                   # (N,2) contiguous


def test_cell_spec_chunk_policy_from_arguments_and_environment(monkeypatch):
    from opendpd_b200.functional import CellSpec
    monkeypatch.delenv("ODPD_TCHUNKS", raising=False); monkeypatch.delenv("ODPD_TWARM", raising=False)
    s = CellSpec("dgru", 13)
    assert s.tchunks == (0, 0) and s.twarm == (0, 0)
    d = s.dims(64, 2048, 0, backward=True)
    assert (d.tchunks, d.twarm, d.x_starts, d.target_starts) == (0, 0, None, None)
    s = CellSpec("dgru", 13, tchunks=(8, 4), twarm=(64, 128))
    assert s.dims(64, 2048, 0).tchunks == 8 and s.dims(64, 2048, 0, backward=True).tchunks == 4
    assert s.dims(64, 2048, 0).twarm == 64 and s.dims(64, 2048, 0, backward=True).twarm == 128
    monkeypatch.setenv("ODPD_TCHUNKS", "1"); monkeypatch.setenv("ODPD_TCHUNKS_BWD", "2"); monkeypatch.setenv("ODPD_TWARM", "96")
    s = CellSpec("lstm", 9)
    assert s.tchunks == (1, 2) and s.twarm == (96, 96)


def test_warmup_policy_of_the_chunk_controller():
    """train.warmup_policy: shrink after three clean checks (never below the floor), grow on failures / near-failures, remember
    a length that failed."""
    from opendpd_b200.train import warmup_policy
    st = dict(clean=0, floor=64, base=128)
    assert warmup_policy(st, 128, False, 0.1) is None and warmup_policy(st, 128, False, 0.2) is None
    assert warmup_policy(st, 128, False, 0.1) == 64 and st["clean"] == 0          # third clean check: halve
    for _ in range(5):
        assert warmup_policy(st, 64, False, 0.05) is None                          # 32 < floor 64: stays
    assert warmup_policy(st, 64, False, 0.3) is None and st["clean"] == 0          # between the thresholds: no change, streak reset
    assert warmup_policy(st, 64, False, 0.6) == 128 and st["floor"] == 128         # below the default and > tol/2: grow proactively
    st = dict(clean=2, floor=64, base=128)
    assert warmup_policy(st, 128, True, 0.1) == 256 and st["floor"] == 256 and st["clean"] == 0      # a failed boundary: double, remember
    assert warmup_policy(st, 256, False, 0.9) is None                              # at/above the default only real failures grow it
    st = dict(clean=0, floor=64, base=256)
    for _ in range(2):
        assert warmup_policy(st, 256, False, 0.0) is None
    assert warmup_policy(st, 256, False, 0.0) == 128


def _plan(B, T, tchunks=0, twarm=0, slots=592, default_warm=128):
    out = (ctypes.c_int32 * 3)()
    assert _ffi.lib().odpd_chunk_plan_model(B, T, tchunks, twarm, slots, default_warm, out) == 0
    return tuple(out)


def test_chunk_planner_headline_plans():
    """csrc/chunking.cuh chunk_make_plan through its host-only entry point (no GPU needed)."""
    assert _plan(64, 2048, slots=592) == (8, 256, 128)              # C2a forward: 4 CTAs/SM x 148
    assert _plan(64, 2048, slots=296) == (4, 512, 128)              # C2a backward: 2 CTAs/SM
    assert _plan(64, 2048, twarm=64, slots=296) == (4, 512, 64)
    assert _plan(64, 2048, tchunks=1) == (1, 2048, 0)               # serial on request
    assert _plan(256, 2048, slots=296) == (1, 2048, 0)              # C3's PA: the sequences alone fill the device
    assert _plan(8, 1024, slots=592) == (8, 128, 128)               # C1
    assert _plan(128, 4096, slots=592, default_warm=256) == (4, 1024, 256)
    assert _plan(3, 19662, slots=592)[0] >= 16                      # net_eval's long segments
    assert _plan(64, 40) == (1, 40, 0)                              # too short to cut
    assert _plan(4, 512, tchunks=4, twarm=64) == (4, 128, 64)


def test_chunk_planner_invariants():
    """Every plan: chunks of whole 32-step blocks that tile [0,T) with no empty chunk, warm-up a multiple of 32, and (auto mode) never
    more CTAs than 2048 rows of scratch, never a warm-up longer than the chunk, never slower than serial under the planner's own cost model."""
    rng = np.random.default_rng(0)
    for _ in range(400):
        B = int(rng.integers(1, 600)); T = int(rng.integers(1, 20000)); slots = int(rng.integers(1, 1300))
        req = int(rng.choice([0, 0, 0, 1, 2, 3, 5, 8, 16, 32, 40])); tw = int(rng.choice([0, 0, 32, 50, 64, 128, 256]))
        C, Lc, Wu = _plan(B, T, req, tw, slots)
        assert C >= 1
        if C == 1:
            assert (Lc, Wu) == (T, 0)
            continue
        assert Lc % 32 == 0 and Wu % 32 == 0 and Wu > 0 and (C - 1) * Lc < T <= C * Lc
        if req > 1:
            assert C <= min(req, 32)
        if req == 0:
            assert Lc >= Wu and B * C <= 2048
            waves = -(-B * C // min(slots, 2048))
            assert waves * (Lc + Wu) < -(-B // min(slots, 2048)) * T


def test_sharded_frame_starts_tile_the_reference_batch():
    """dp.shard_batch_indices + dp.frame_starts: the ranks' start vectors concatenate to the single-process batch of the same seeded
    permutation, and IqStream on them reproduces IQFrameDataset's frames (CPU tensors: framing logic only)."""
    from opendpd_b200 import dp
    from opendpd_b200.functional import IqStream
    g = torch.Generator().manual_seed(1)
    stream = torch.randn(700, 2, generator=g)
    T, B, world = 64, 10, 3
    perm = dp.epoch_permutation(700 - T + 1, seed=0)
    for step in (0, 1, 63):
        parts = [dp.frame_starts(dp.shard_batch_indices(perm, step, B, r, world)[0], "cpu") for r in range(world)]
        full = perm[step * B:(step + 1) * B]
        assert torch.equal(torch.cat(parts).long(), full) and all(p.dtype == torch.int32 for p in parts)
        for p in parts:
            if p.numel():
                assert torch.equal(IqStream(stream, p, T).frames(), torch.stack([stream[k:k + T] for k in p.tolist()]))


def _manifest():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "MANIFEST.json")))


@pytest.mark.parametrize("name", golden_cases() + wide_cases())
def test_init_is_bit_identical_to_the_reference(name):
    """`reset_parameters` is kept in Python on the same RNG stream, so for the seed the reference-made fixture
    was built with (oracle/make_golden.py: torch.manual_seed(seed) then the reference constructor) the native model must start
    from bit-identical weights — model ids embed the parameter count and checkpoints interchange."""
    g = load_golden(name)
    seed = _manifest()[name]["seed"]
    torch.manual_seed(seed)
    if g["kind"].endswith("_qat"):
        from opendpd_b200.quant import get_quant_model

        class _Proj:
            quant, n_bits_w, n_bits_a, pretrained_model = True, g["K"] & 255, (g["K"] >> 8) & 255, ""
        net = get_quant_model(_Proj(), models.CoreModel(2, g["H"], 1, g["kind"][:-4]))
    else:
        net = models.CoreModel(2, max(g["H"], 1), g["L"], g["kind"], num_dvr_units=g["K"], thx=g["thx"], thh=g["thh"])
    mine = np.concatenate([p.detach().numpy().ravel() for _, p in net.backbone.named_parameters()])
    assert mine.dtype == np.float32 and np.array_equal(mine, g["params"]), f"{name}: initial weights differ from the reference's"


def _integration_snippet():
    """The python block of INTEGRATION.md §2 (the reference-side ctypes binding), verbatim."""
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", md, flags=re.S)
    code = [b for b in blocks if "class OdpdDims(ctypes.Structure)" in b]
    assert len(code) == 1
    return code[0]


def test_integration_snippet_matches_header():
    """INTEGRATION.md's documented binding must describe the struct the library reads: exec the block verbatim (the library is
    already loaded under its SONAME, so the snippet's CDLL("libodpd.so") resolves to it) and compare with _ffi / include/odpd.h."""
    _ffi.lib()
    ns = {}
    exec(compile(_integration_snippet(), "INTEGRATION.md", "exec"), ns)
    D = ns["OdpdDims"]
    assert ctypes.sizeof(D) == ctypes.sizeof(_ffi.OdpdDims) == 56
    assert [(n, t) for n, t in D._fields_] == [(n, t) for n, t in _ffi.OdpdDims._fields_]
    assert [(n, getattr(D, n).offset) for n, _ in D._fields_] == [(n, getattr(_ffi.OdpdDims, n).offset) for n, _ in _ffi.OdpdDims._fields_]
    # and against the C header itself: field order of `typedef struct OdpdDims`
    hdr = open(os.path.join(ROOT, "include", "odpd.h")).read()
    body = hdr[hdr.index("typedef struct OdpdDims {"):hdr.index("} OdpdDims;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if decl:
            names += [v.strip().lstrip("*") for v in decl.split(None, 1)[1].replace("const int32_t", "").split(",")] if "," in decl else [decl.split()[-1].lstrip("*")]
    assert names == [n for n, _ in D._fields_], names


def test_qat_pretrained_model_loads_weights_only(tmp_path):
    """quant/quant_envs.py:173-182 + quant_layers.py:53-60: a --pretrained_model checkpoint of the GRU-swapped float model provides the
    three weight matrices; the INT_Linear biases stay the wrapper's own fresh draws (same RNG positions as without a checkpoint)."""
    from opendpd_b200.quant import get_quant_model
    H = 10
    sd = {"backbone.rnn.rnn_cell_list.0.x2h.weight": torch.full((3 * H, 4), 0.25), "backbone.rnn.rnn_cell_list.0.x2h.bias": torch.zeros(3 * H),
          "backbone.rnn.rnn_cell_list.0.h2h.weight": torch.full((3 * H, H), -0.5), "backbone.rnn.rnn_cell_list.0.h2h.bias": torch.zeros(3 * H),
          "backbone.fc_out.weight": torch.full((2, H), 0.125), "backbone.fc_out.bias": torch.zeros(2)}
    path = str(tmp_path / "pre.pt")
    torch.save(sd, path)

    class _P:
        quant, n_bits_w, n_bits_a = True, 8, 8
        pretrained_model = ""
    torch.manual_seed(3)
    plain = get_quant_model(_P(), models.CoreModel(2, H, 1, "qgru"))
    _P.pretrained_model = path
    torch.manual_seed(3)
    pre = get_quant_model(_P(), models.CoreModel(2, H, 1, "qgru"))
    c0, c1 = plain.backbone.rnn.rnn_cell_list[0], pre.backbone.rnn.rnn_cell_list[0]
    assert torch.equal(c1.x2h.weight, sd["backbone.rnn.rnn_cell_list.0.x2h.weight"]) and torch.equal(c1.h2h.weight, sd["backbone.rnn.rnn_cell_list.0.h2h.weight"])
    assert torch.equal(pre.backbone.fc_out.weight, sd["backbone.fc_out.weight"])
    assert torch.equal(c1.x2h.bias, c0.x2h.bias) and torch.equal(c1.h2h.bias, c0.h2h.bias) and torch.equal(pre.backbone.fc_out.bias, plain.backbone.fc_out.bias)
    torch.save({"backbone.rnn.weight_ih_l0": torch.zeros(3 * H, 4)}, path)       # a float nn.GRU checkpoint: the reference silently trains in float
    with pytest.raises(ValueError):
        get_quant_model(_P(), models.CoreModel(2, H, 1, "qgru"))

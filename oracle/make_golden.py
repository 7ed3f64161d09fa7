#!/usr/bin/env python
"""Generate golden input/output vectors from the UNMODIFIED reference (test infrastructure only).

Runs in the authoring container only: imports lab-emi/OpenDPD from /root/reference (read-only),
builds every hot-path backbone through the reference's own constructors, runs
forward + nn.MSELoss + backward on CPU in fp32 (and in fp64 as arbiter) on frames cut from the
reference's APA_200MHz dataset, and writes small .npz fixtures under tests/golden/.

The fixtures travel to the GPU box; /root/reference does not.  Nothing in the product imports
this file.  Re-run:  python oracle/make_golden.py
"""
import os, sys, json, hashlib
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
REF = "/root/reference"
sys.path.insert(0, REF)
import numpy as np
import torch
import torch.nn as nn

import quant  # noqa: E402  (reference package)
from quant.modules.ops import Sqrt, Pow  # qgru.py:7 import bug shim
quant.Sqrt, quant.Pow = Sqrt, Pow
import models  # noqa: E402  reference models.py
from backbones.pgjanet import PGJANET  # CoreModel('pgjanet') raises TypeError (models.py:111-114)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
torch.set_num_threads(4)


def load_apa():
    import pandas as pd
    d = os.path.join(REF, "datasets", "APA_200MHz")
    X = pd.read_csv(os.path.join(d, "train_input.csv")).to_numpy(dtype=np.float64)
    Y = pd.read_csv(os.path.join(d, "train_output.csv")).to_numpy(dtype=np.float64)
    return X, Y


def frames(X, Y, B, T, seed):
    """Frames exactly as IQFrameDataset makes them (data_collector.py:240-247): rows [k,k+T), f64->f32."""
    g = torch.Generator().manual_seed(seed)
    n = X.shape[0] - T + 1
    idx = torch.randperm(n, generator=g)[:B].tolist()
    x = np.stack([X[k:k + T] for k in idx]).astype(np.float32)
    y = np.stack([Y[k:k + T] for k in idx]).astype(np.float32)
    return x, y


def build(kind, H, seed, thx=0.0, thh=0.0, K=3, L=1):
    torch.manual_seed(seed)
    if kind.endswith("_qat"):
        # config 5: the reference's own QAT environment (quant/__init__.py:20-37 -> quant_envs.py Base_GRUQuantEnv) around the float QGRU
        from quant import get_quant_model
        float_net = models.CoreModel(input_size=2, hidden_size=H, num_layers=1, backbone_type=kind[:-4], thx=thx, thh=thh)

        class _Proj:
            quant, n_bits_w, n_bits_a, pretrained_model, quant_dir_label = True, K & 255, (K >> 8) & 255, "", ""
        qnet = get_quant_model(_Proj(), float_net)
        assert qnet is not float_net and type(qnet.backbone.fc_out).__name__ == "INT_Linear", "quant env fell back to the float model"
        qnet.train()
        return qnet
    if kind == "pgjanet":
        bb = PGJANET(hidden_size=H, output_size=2, bias=True)
        bb.reset_parameters()
        net = models.CoreModel.__new__(models.CoreModel)
        nn.Module.__init__(net)
        net.num_layers, net.hidden_size, net.backbone = 1, H, bb
        return net
    return models.CoreModel(input_size=2, hidden_size=H, num_layers=L, backbone_type=kind,
                            num_dvr_units=K, thx=thx, thh=thh)


def flat_params(net):
    return np.concatenate([p.detach().cpu().numpy().reshape(-1).astype(np.float64)
                           for _, p in net.backbone.named_parameters()])


def param_index(net):
    return [(n, list(p.shape)) for n, p in net.backbone.named_parameters()]


class MaskTap:
    """Record the delta keep-masks of the reference's DeltaGRULayer.compute_deltas without editing it."""

    def __init__(self, layer_cls):
        self.cls, self.mx, self.mh = layer_cls, [], []
        self.orig = layer_cls.__dict__["compute_deltas"].__func__

    def __enter__(self):
        tap = self

        def wrapped(x, x_p, h, h_p, th_x, th_h):
            r = tap.orig(x, x_p, h, h_p, th_x, th_h)
            tap.mx.append((r[2] >= th_x).numpy().copy())
            tap.mh.append((r[3] >= th_h).numpy().copy())
            return r
        self.cls.compute_deltas = staticmethod(wrapped)
        return self

    def __exit__(self, *a):
        self.cls.compute_deltas = staticmethod(self.orig)

    def packed(self):
        def pack(lst):  # list over t of (B, n) bool -> (B, T) uint64 bitfield, bit k = unit k kept
            a = np.stack(lst, axis=1)
            w = (1 << np.arange(a.shape[-1], dtype=np.uint64))
            return (a.astype(np.uint64) * w).sum(-1).astype(np.uint64)
        return pack(self.mx), pack(self.mh)


def run(net, x, y, dtype, tap_cls=None):
    # CoreModel.forward (models.py:154-155) and GMP (gmp.py:21,26) allocate default-dtype tensors, so the
    # fp64 arbiter run switches the default dtype instead of editing the reference.
    torch.set_default_dtype(dtype)
    try:
        return _run(net, x, y, dtype, tap_cls)
    finally:
        torch.set_default_dtype(torch.float32)


def _run(net, x, y, dtype, tap_cls=None):
    net = net.to(dtype)
    xt = torch.tensor(x, dtype=dtype, requires_grad=True)
    yt = torch.tensor(y, dtype=dtype)
    for p in net.parameters():
        p.grad = None
    res = {}
    if tap_cls is not None:
        net.backbone.set_debug(1)
        with MaskTap(tap_cls) as tap:
            out = net(xt)
        res["mask_x"], res["mask_h"] = tap.packed()
        st = net.backbone.rnn.statistics
        res["stats"] = np.array([float(st["num_dx_zeros"]), float(st["num_dx_numel"]),
                                 float(st["num_dh_zeros"]), float(st["num_dh_numel"])], dtype=np.float64)
    else:
        out = net(xt)
    loss = nn.MSELoss()(out, yt)
    loss.backward()
    res["out"] = out.detach().numpy()
    res["loss"] = np.array(loss.item(), dtype=np.float64)
    res["gx"] = xt.grad.numpy()
    res["gparams"] = np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).numpy().reshape(-1)
                                     for _, p in net.backbone.named_parameters()])   # unused quantiser scales: grad None == 0
    return res


# (name, kind, H, B, T, data seed, thx, thh)
CASES = [
    ("gru_h32_b8_t128",      "gru",  32, 8, 128, 1, 0, 0),
    ("gru_h23_b3_t17",       "gru",  23, 3, 17, 2, 0, 0),
    ("gru_h8_b1_t1",         "gru",   8, 1, 1, 3, 0, 0),
    ("dgru_h13_b8_t256",     "dgru", 13, 8, 256, 4, 0, 0),
    ("dgru_h13_b4_t48",      "dgru", 13, 4, 48, 5, 0, 0),
    ("dgru_h8_b3_t33",       "dgru",  8, 3, 33, 6, 0, 0),
    ("dgru_h23_b2_t65",      "dgru", 23, 2, 65, 7, 0, 0),
    ("lstm_h9_b4_t64",       "lstm",  9, 4, 64, 8, 0, 0),
    ("lstm_h16_b2_t33",      "lstm", 16, 2, 33, 9, 0, 0),
    ("deltagru_h15_b4_t96",  "deltagru", 15, 4, 96, 10, 0.01, 0.05),
    ("deltagru_h15_b2_t40_th0", "deltagru", 15, 2, 40, 11, 0.0, 0.0),
    ("deltagru_h10_b3_t33",  "deltagru", 10, 3, 33, 12, 0.02, 0.02),
    ("tres_h15_b4_t96",      "deltagru_tcnskip", 15, 4, 96, 13, 0.01, 0.05),
    ("tres_h15_b2_t200",     "deltagru_tcnskip", 15, 2, 200, 14, 0.01, 0.05),
    ("tres_h15_b3_t20",      "deltagru_tcnskip", 15, 3, 20, 15, 0.0, 0.0),
    ("pgjanet_h15_b4_t64",   "pgjanet", 15, 4, 64, 16, 0, 0),
    ("pgjanet_h10_b2_t33",   "pgjanet", 10, 2, 33, 17, 0, 0),
    ("dvrjanet_h15_b4_t64",  "dvrjanet", 15, 4, 64, 18, 0, 0),
    ("dvrjanet_h10_b2_t33",  "dvrjanet", 10, 2, 33, 19, 0, 0),
    ("gmp_b4_t64",           "gmp", 0, 4, 64, 20, 0, 0),
    ("gmp_b2_t7",            "gmp", 0, 2, 7, 21, 0, 0),
    ("qgru_h10_b4_t50",      "qgru", 10, 4, 50, 22, 0, 0),
    ("qgru_amp1_h10_b4_t50", "qgru_amp1", 10, 4, 50, 23, 0, 0),
    # QAT (config 5): K packs bits_w | bits_a<<8
    ("qgruqat_w8a8_h10_b4_t50", "qgru_qat", 10, 4, 50, 24, 0, 0, 8 | (8 << 8)),
    ("qgruqat_w16a16_h10_b3_t33", "qgru_qat", 10, 3, 33, 25, 0, 0, 16 | (16 << 8)),
    ("qgruamp1qat_w8a8_h8_b2_t20", "qgru_amp1_qat", 8, 2, 20, 26, 0, 0, 8 | (8 << 8)),
    # row f-4: RVTDCNN (H = fc_hid_size, models.py:80-81)
    ("rvtdcnn_h6_b3_t40",          "rvtdcnn", 6, 3, 40, 50, 0, 0),
    ("rvtdcnn_h20_b2_t70",         "rvtdcnn", 20, 2, 70, 51, 0, 0),
    ("rvtdcnn_h64_b2_t3",          "rvtdcnn", 64, 2, 3, 52, 0, 0),
    ("bojanet_h10_b3_t50",         "bojanet", 10, 3, 50, 53, 0, 0),
    ("bojanet_h18_b2_t70",         "bojanet", 18, 2, 70, 54, 0, 0),
    ("bojanet_h4_b2_t15",          "bojanet", 4, 2, 15, 55, 0, 0),
    ("tcnn_h8_b3_t70",             "tcnn", 8, 3, 70, 56, 0, 0),
    ("tcnn_h20_b2_t150",           "tcnn", 20, 2, 150, 57, 0, 0),
    ("tcnn_h64_b2_t9",             "tcnn", 64, 2, 9, 58, 0, 0),
    ("neuraltx_h8_b3_t70",         "neuraltx", 8, 3, 70, 59, 0, 0),
    ("neuraltx_h24_b2_t131",       "neuraltx", 24, 2, 131, 60, 0, 0),
    ("neuraltx_h3_b2_t4",          "neuraltx", 3, 2, 4, 61, 0, 0),
    ("apnrru_h8_b3_t50",           "apnrru", 8, 3, 50, 62, 0, 0),
    ("apnrru_h14_b2_t70",          "apnrru", 14, 2, 70, 63, 0, 0),
    ("apnrru_h3_b2_t15",           "apnrru", 3, 2, 15, 64, 0, 0),
    ("mcldnn_h8_b3_t50",           "mcldnn", 8, 3, 50, 65, 0, 0),
    ("mcldnn_h12_b2_t70",          "mcldnn", 12, 2, 70, 66, 0, 0),
    ("mcldnn_h2_b2_t5",            "mcldnn", 2, 2, 5, 67, 0, 0),
    ("deltajanet_h10_b3_t50",      "deltajanet", 10, 3, 50, 68, 0.01, 0.05),     # thx / thh are accepted and ignored by the reference (deltajanet.py:22-26)
    ("deltajanet_h16_b2_t70",      "deltajanet", 16, 2, 70, 69, 0, 0),
    ("deltajanet_h3_b2_t2",        "deltajanet", 3, 2, 2, 70, 0, 0),
    # the W16A16 stage of bash_scripts/OpenDPDv2.sh:47-49: the reference's QAT env around TRes-DeltaGRU (K packs bits_w | bits_a<<8)
    ("tresqat_w16a16_h15_b3_t60",  "deltagru_tcnskip_qat", 15, 3, 60, 71, 0.01, 0.05, 16 | (16 << 8)),
    ("tresqat_w8a8_h10_b2_t40",    "deltagru_tcnskip_qat", 10, 2, 40, 72, 0.02, 0.05, 8 | (8 << 8)),
    # hidden sizes above the fused tiers and stacked layers (arguments.py:51,60 -> nn.GRU/nn.LSTM num_layers): 10th field = num_layers
    ("wide_gru_h48_b3_t70",        "gru",  48, 3, 70, 40, 0, 0, 3, 1),
    ("wide_gru_h16_l2_b3_t40",     "gru",  16, 3, 40, 41, 0, 0, 3, 2),
    ("wide_dgru_h40_l2_b2_t70",    "dgru", 40, 2, 70, 42, 0, 0, 3, 2),
    ("wide_dgru_h64_b2_t33",       "dgru", 64, 2, 33, 43, 0, 0, 3, 1),
    ("wide_lstm_h64_b2_t40",       "lstm", 64, 2, 40, 44, 0, 0, 3, 1),
    ("wide_lstm_h12_l3_b3_t37",    "lstm", 12, 3, 37, 45, 0, 0, 3, 3),
    ("wide_qgru_h36_b2_t50",       "qgru", 36, 2, 50, 46, 0, 0, 3, 1),
    ("wide_qgru_amp1_h20_l2_b2_t45", "qgru_amp1", 20, 2, 45, 47, 0, 0, 3, 2),
]


def main():
    X, Y = load_apa()
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    only = set(sys.argv[1:])
    for case in CASES:
        name, kind, H, B, T, seed, thx, thh = case[:8]
        K = case[8] if len(case) > 8 else 3
        L = case[9] if len(case) > 9 else 1
        if only and name not in only:
            continue
        x, y = frames(X, Y, B, T, seed)
        net = build(kind, max(H, 1), seed, thx, thh, K, L)
        tap_cls = None
        if kind == "deltagru":
            from backbones.deltagru import DeltaGRULayer as tap_cls
        elif kind in ("deltagru_tcnskip", "deltagru_tcnskip_qat"):
            from backbones.deltagru_tcnskip import DeltaGRULayer as tap_cls
        params = flat_params(net)
        r32 = run(net, x, y, torch.float32, tap_cls)
        r64 = run(net, x, y, torch.float64, tap_cls)
        net.float()
        rec = dict(x=x, y=y, params=params.astype(np.float32), kind=np.array(kind), H=np.array(H),
                   thx=np.array(thx, dtype=np.float64), thh=np.array(thh, dtype=np.float64),
                   K=np.array(K), L=np.array(L), param_index=np.array(json.dumps(param_index(net))))
        for k, v in r32.items():
            rec[k] = v
        for k, v in r64.items():
            rec[k + "64"] = v
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        manifest[name] = dict(kind=kind, H=H, B=B, T=T, seed=seed, n_params=int(params.size), loss=float(r32["loss"]),
                              bytes=os.path.getsize(path))
        print(name, manifest[name], flush=True)
    mpath = os.path.join(OUT, "MANIFEST.json")
    if only and os.path.exists(mpath):
        old = json.load(open(mpath)); old.update(manifest); manifest = old
    with open(mpath, "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()

"""`get_quant_model` — native counterpart of the reference's quant.get_quant_model (quant/__init__.py:20-37)."""
from .backbones.qgru_quant import QGRUQuant


def _load_pretrained(qbb, path):
    """`--pretrained_model` (reference quant/quant_envs.py:173-182): the checkpoint is the state_dict of the float model AFTER the
    nn.GRU -> Python GRU swap (`backbone.rnn.rnn_cell_list.0.{x2h,h2h}.{weight,bias}`, `backbone.fc_out.{weight,bias}`), loaded before
    the layers are wrapped.  The wrapping (INT_Linear, quant_layers.py:53-60) then adopts only the WEIGHT of the layer it replaces — the
    bias stays INT_Linear's own fresh draw — so exactly the three weight matrices come from the checkpoint.  A checkpoint with other
    keys (e.g. saved from the nn.GRU-based float QGRU, `backbone.rnn.weight_ih_l0`) makes the reference fall back to the FLOAT model
    with only a warning; here it raises."""
    import torch
    sd = torch.load(path, map_location="cpu")
    want = {"backbone.rnn.rnn_cell_list.0.x2h.weight": qbb.rnn.rnn_cell_list[0].x2h.weight,
            "backbone.rnn.rnn_cell_list.0.h2h.weight": qbb.rnn.rnn_cell_list[0].h2h.weight,
            "backbone.fc_out.weight": qbb.fc_out.weight}
    need = set(want) | {"backbone.rnn.rnn_cell_list.0.x2h.bias", "backbone.rnn.rnn_cell_list.0.h2h.bias", "backbone.fc_out.bias"}
    if set(sd) != need:
        raise ValueError(f"--pretrained_model {path}: expected the state_dict of the GRU-swapped float QGRU (keys {sorted(need)}), "
                         f"got {sorted(sd)[:6]}...; the reference would silently continue with the float model here")
    with torch.no_grad():
        for k, p in want.items():
            if tuple(sd[k].shape) != tuple(p.shape):
                raise ValueError(f"--pretrained_model {path}: {k} has shape {tuple(sd[k].shape)}, model wants {tuple(p.shape)}")
            p.copy_(sd[k])


def _load_pretrained_tres(qbb, path):
    """`--pretrained_model` for the TRes-DeltaGRU (OpenDPDv2.sh:84-111: the float DPD trained in the previous stage).  This backbone has no
    nn.GRU to swap, so the checkpoint is simply the float model's state_dict and all five weight tensors come from it
    (quant_envs.py:173-182 loads it before the layers are wrapped; the wrapping keeps the weights)."""
    import torch
    sd = torch.load(path, map_location="cpu")
    want = {"backbone.rnn.x2h.weight": qbb.rnn.x2h.weight, "backbone.rnn.h2h.weight": qbb.rnn.h2h.weight, "backbone.fc_out.weight": qbb.fc_out.weight,
            "backbone.tcn.0.weight": qbb.tcn[0].weight, "backbone.tcn.2.weight": qbb.tcn[2].weight}
    if set(sd) != set(want):
        raise ValueError(f"--pretrained_model {path}: expected the state_dict of the float deltagru_tcnskip model (keys {sorted(want)}), got "
                         f"{sorted(sd)[:6]}...; the reference would silently continue with the float model here")
    with torch.no_grad():
        for k, p in want.items():
            if tuple(sd[k].shape) != tuple(p.shape):
                raise ValueError(f"--pretrained_model {path}: {k} has shape {tuple(sd[k].shape)}, model wants {tuple(p.shape)}")
            p.copy_(sd[k])


def get_quant_model(proj, model):
    """If proj.quant is truthy return a CoreModel whose backbone is the fake-quantised QGRU built from `model`'s float QGRU
    (n_bits_w / n_bits_a from proj, default 8; proj.pretrained_model honoured); otherwise return `model` unchanged.  Unlike the reference this does NOT
    silently fall back to the float model when the setup fails: it raises."""
    if not getattr(proj, "quant", False):
        return model
    bb = model.backbone
    if getattr(bb, "cell", None) not in ("qgru", "qgru_amp1", "deltagru_tcnskip"):
        raise ValueError("native QAT is available for the qgru / qgru_amp1 / deltagru_tcnskip backbones (the ones the reference's QAT scripts use: "
                         "quant_qgru_dpd_regr.sh, quant_mp_dpd.sh, OpenDPDv2.sh)")
    dev = next(bb.parameters()).device
    pretrained = getattr(proj, "pretrained_model", "")
    if bb.cell == "deltagru_tcnskip":
        from .backbones.tres_quant import TResQuant
        qbb = TResQuant.from_float(bb, getattr(proj, "n_bits_w", 8), getattr(proj, "n_bits_a", 8))
        if pretrained:
            _load_pretrained_tres(qbb, pretrained)
    else:
        qbb = QGRUQuant.from_float(bb, getattr(proj, "n_bits_w", 8), getattr(proj, "n_bits_a", 8))
        if pretrained:
            _load_pretrained(qbb, pretrained)
    qbb = qbb.to(dev)
    import copy
    qmodel = copy.copy(model)
    qmodel._modules = dict(model._modules)
    qmodel.backbone = qbb
    return qmodel

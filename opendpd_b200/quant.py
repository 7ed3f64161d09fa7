"""`get_quant_model` — native counterpart of the reference's quant.get_quant_model (quant/__init__.py:20-37)."""
from .backbones.qgru_quant import QGRUQuant


def get_quant_model(proj, model):
    """If proj.quant is truthy return a CoreModel whose backbone is the fake-quantised QGRU built from `model`'s float QGRU
    (n_bits_w / n_bits_a from proj, default 8); otherwise return `model` unchanged.  Unlike the reference this does NOT
    silently fall back to the float model when the setup fails: it raises."""
    if not getattr(proj, "quant", False):
        return model
    bb = model.backbone
    if getattr(bb, "cell", None) not in ("qgru", "qgru_amp1"):
        raise ValueError("native QAT is available for the qgru / qgru_amp1 backbones (the ones the reference's QAT scripts use)")
    if getattr(proj, "pretrained_model", ""):
        raise NotImplementedError("--pretrained_model for QAT: load the state_dict into the returned model instead")
    dev = next(bb.parameters()).device
    qbb = QGRUQuant.from_float(bb, getattr(proj, "n_bits_w", 8), getattr(proj, "n_bits_a", 8)).to(dev)
    import copy
    qmodel = copy.copy(model)
    qmodel._modules = dict(model._modules)
    qmodel.backbone = qbb
    return qmodel

"""Drop-in for the reference's `quant` package as far as the step scripts use it (steps/train_dpd.py:12, steps/run_dpd.py:16-17):
`get_quant_model(proj, model)` returns the native fake-quantised QGRU (csrc/qgru_qat.cu) instead of performing module surgery on an
nn.GRU.  `AttrDict` is kept because callers of the reference's quant package construct it (quant/quant_envs.py)."""
from opendpd_b200.quant import get_quant_model  # noqa: F401


class AttrDict(dict):
    """quant/quant_envs.py AttrDict: dict with attribute access."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


__all__ = ["get_quant_model", "AttrDict"]

"""Fake-quantised (QAT) TRes-DeltaGRU — native drop-in for `quant.get_quant_model(proj, model)` (quant/__init__.py:20-37 ->
Base_GRUQuantEnv, quant/quant_envs.py:132-305) applied to backbones/deltagru_tcnskip.py: the W16A16 stage of bash_scripts/OpenDPDv2.sh:47-49.

Same parameter / buffer names and order as the surgered reference model (`rnn.x2h.weight`, `rnn.x2h.weight_quantizer.scale`, ...,
`rnn.add.quantizer.scale`, `fc_out.out_quantizer.scale`, `tcn.0.weight`), same count (1012 at H=15) and the same initial values: the
weights are the float model's (INT_Linear adopts them, quant_layers.py:53-60; there is no nn.GRU to swap in this backbone, so
create_pygru_model changes nothing), the 13 scales start at 2^(2-bits) (quantizers.py:44-48, 86-88; 2^-14 for the 16-bit output
quantisers).  The arithmetic runs in csrc/tres_qat.cu."""
import torch
from torch import nn
from .deltagru import _DeltaBase
from .qgru_quant import _Quantizer, _QOp
from ..functional import CellSpec


class _QLinearNoBias(nn.Module):
    """Container of INT_Linear (quant_layers.py:53-82) for a bias-free Linear: weight, three quantisers, two bit-width buffers."""

    def __init__(self, weight, bits_w, bits_a):
        super().__init__()
        self.weight = nn.Parameter(weight)
        self.bias = None
        self.weight_quantizer = _Quantizer(bits_w, 2.0 ** (2 - bits_w))
        self.act_quantizer = _Quantizer(bits_a, 2.0 ** (2 - bits_a))
        self.out_quantizer = _Quantizer(16, 2.0 ** (2 - 16))
        self.out_quant = False
        self.register_buffer("n_bits_w", torch.Tensor([bits_w]))
        self.register_buffer("n_bits_a", torch.Tensor([bits_a]))


class _QTResLayer(nn.Module):
    """rnn.* of the surgered model: x2h, h2h (INT_Linear) and the four quantised ops in the layer's registration order
    (deltagru_tcnskip.py:156-162: add, mul, sigmoid, tanh)."""

    def __init__(self, x2h_w, h2h_w, bits_w, bits_a, thx, thh):
        super().__init__()
        self.input_size, self.hidden_size, self.th_x, self.th_h, self.debug = 6, h2h_w.shape[1], thx, thh, 1
        self.x2h = _QLinearNoBias(x2h_w, bits_w, bits_a)
        self.h2h = _QLinearNoBias(h2h_w, bits_w, bits_a)
        self.add, self.mul, self.sigmoid, self.tanh = _QOp(bits_a), _QOp(bits_a), _QOp(bits_a), _QOp(bits_a)


class TResQuant(_DeltaBase):
    cell = "deltagru_tcnskip_qat"

    def __init__(self, hidden_size, x2h_w, h2h_w, fc_w, tcn0_w, tcn2_w, thx=0.0, thh=0.0, n_bits_w=16, n_bits_a=16):
        super().__init__()
        if not 1 <= hidden_size <= 16:
            raise NotImplementedError(f"native fake-quantised TRes-DeltaGRU: hidden_size 1..16 (OpenDPDv2.sh uses 15; got {hidden_size})")
        self.hidden_size, self.input_size, self.output_size, self.num_layers = hidden_size, 6, 2, 1
        self.thx, self.thh, self.bias = thx, thh, True
        self.n_bits_w, self.n_bits_a = int(n_bits_w), int(n_bits_a)
        self.rnn = _QTResLayer(x2h_w, h2h_w, n_bits_w, n_bits_a, thx, thh)
        self.fc_out = _QLinearNoBias(fc_w, n_bits_w, n_bits_a)
        self.fc_out.out_quant = True                                            # set_last_layer_quant, quant_envs.py:276-284
        self.tcn = nn.Sequential(
            nn.Conv1d(in_channels=2, out_channels=3, kernel_size=3, padding=16, stride=1, dilation=16, bias=False),
            nn.Hardswish(),
            nn.Conv1d(in_channels=3, out_channels=2, kernel_size=1, padding=0, stride=1, dilation=1, bias=False),
            nn.Hardswish(),
        )
        with torch.no_grad():
            self.tcn[0].weight.copy_(tcn0_w)
            self.tcn[2].weight.copy_(tcn2_w)
        self._init_stats()

    def _spec(self):
        spec = CellSpec(self.cell, self.hidden_size, self.n_bits_w | (self.n_bits_a << 8) | ((0 if self.training else 1) << 16), self.thx, self.thh)
        if getattr(self, "keep_masks", False):
            spec.keep_saved = self
        return spec

    def reset_parameters(self):
        raise AttributeError("the quantised model is built from a float model (TResQuant.from_float)")

    def sync_quant_buffers(self):
        for m in self.modules():
            if isinstance(m, _Quantizer):
                m.sync_buffers()

    def _fc_numel(self):
        return self.fc_out.weight.numel() + sum(p.numel() for p in self.tcn.parameters())

    @classmethod
    def from_float(cls, bb, n_bits_w=16, n_bits_a=16):
        """Same RNG consumption as Base_GRUQuantEnv.__init__ for this backbone: INT_Linear.__init__ runs nn.Linear.__init__ (a fresh
        weight draw each, then discarded) for x2h, h2h, fc_out in traversal order (quant_envs.py:41-60, quant_layers.py:53-57)."""
        H = bb.hidden_size
        nn.Linear(6, 3 * H, bias=False); nn.Linear(H, 3 * H, bias=False); nn.Linear(H, 2, bias=False)
        cp = lambda t: t.detach().clone().cpu()
        return cls(H, cp(bb.rnn.x2h.weight), cp(bb.rnn.h2h.weight), cp(bb.fc_out.weight), cp(bb.tcn[0].weight), cp(bb.tcn[2].weight),
                   bb.thx, bb.thh, n_bits_w, n_bits_a)

"""Fake-quantised (QAT) QGRU — native drop-in for what the reference builds by module surgery:
`quant.get_quant_model(proj, model)` (quant/__init__.py:20-37) -> `Base_GRUQuantEnv` (quant/quant_envs.py:132-305) applied to
backbones/qgru.py / qgru_amp1.py.  Same parameter/buffer names as the surgered reference model
(`rnn.rnn_cell_list.0.x2h.weight`, `...x2h.weight_quantizer.scale`, `fc_out.out_quantizer.scale`, ...), same parameter count
(515 at H=10) and — through `from_float` — the same initial values for the same RNG state, including the reference quirks:
the swapped GRUCell is re-initialised (weights are NOT taken from the float nn.GRU, quant_envs.py:215-248) and INT_Linear keeps
the freshly drawn nn.Linear bias instead of the float layer's (quant_layers.py:53-60).  The arithmetic runs in csrc/qgru_qat.cu."""
import math
import torch
from torch import nn
from ._base import NativeBackbone
from ..functional import CellSpec


class _Quantizer(nn.Module):
    """Parameter/buffer container of quant/qmodules/quantizers.py INT_Quantizer (:15-54)."""

    def __init__(self, bits, init_scale):
        super().__init__()
        self.bits = bits
        self.scale = nn.Parameter(torch.Tensor([init_scale]))
        self.register_buffer("pow2_scale", torch.Tensor([0.0]))
        self.register_buffer("decimal_num", torch.Tensor([1.0]))
        self.register_buffer("integer_num", bits - 1 - self.decimal_num)

    def sync_buffers(self):
        """What INT_Quantizer.forward maintains on every call (quantizers.py:67-71)."""
        with torch.no_grad():
            l2 = self.scale.abs().log2().round()
            self.pow2_scale.copy_(2 ** l2)
            self.decimal_num.copy_(l2.abs())
            self.integer_num.copy_(self.bits - 1 - self.decimal_num)


class _QLinear(nn.Module):
    """Container of quant/qmodules/quant_layers.py INT_Linear (:53-82): weight, bias, three quantisers, two bit-width buffers."""

    def __init__(self, weight, bias, bits_w, bits_a):
        super().__init__()
        self.weight = nn.Parameter(weight)
        self.bias = nn.Parameter(bias)
        self.weight_quantizer = _Quantizer(bits_w, 2.0 ** (2 - bits_w))     # init_act_params, quantizers.py:44-48
        self.act_quantizer = _Quantizer(bits_a, 2.0 ** (2 - bits_a))
        self.out_quantizer = _Quantizer(16, 2.0 ** (2 - 16))
        self.out_quant = False
        self.register_buffer("n_bits_w", torch.Tensor([bits_w]))
        self.register_buffer("n_bits_a", torch.Tensor([bits_a]))


class _QOp(nn.Module):
    """Container of Quant_sigmoid / Quant_tanh / Quant_add / Quant_mult (quant_ops.py:14-66): one OP_INT_Quantizer each."""

    def __init__(self, bits):
        super().__init__()
        self.quantizer = _Quantizer(bits, 2.0 ** (2 - bits))                 # OP_INT_Quantizer.init_params


class _QCell(nn.Module):
    def __init__(self, x2h, h2h, bits_a):
        super().__init__()
        self.x2h, self.h2h = x2h, h2h
        self.sigmoid, self.tanh, self.add, self.mul = _QOp(bits_a), _QOp(bits_a), _QOp(bits_a), _QOp(bits_a)


class _QRnn(nn.Module):
    def __init__(self, cell):
        super().__init__()
        self.rnn_cell_list = nn.ModuleList([cell])


class QGRUQuant(NativeBackbone):
    def __init__(self, hidden_size, x2h_w, x2h_b, h2h_w, h2h_b, fc_w, fc_b, n_bits_w=8, n_bits_a=8, amp1=False):
        super().__init__()
        self.hidden_size, self.input_size, self.output_size, self.num_layers = hidden_size, 4, 2, 1
        self.n_bits_w, self.n_bits_a, self.amp1 = int(n_bits_w), int(n_bits_a), bool(amp1)
        self.cell = "qgru_amp1_qat" if amp1 else "qgru_qat"
        cell = _QCell(_QLinear(x2h_w, x2h_b, n_bits_w, n_bits_a), _QLinear(h2h_w, h2h_b, n_bits_w, n_bits_a), n_bits_a)
        self.rnn = _QRnn(cell)
        self.fc_out = _QLinear(fc_w, fc_b, n_bits_w, n_bits_a)
        self.fc_out.out_quant = True                                            # set_last_layer_quant, quant_envs.py:268-277

    def _spec(self):
        # the 16-bit output quantiser is active only in eval (quant_layers.py:77-80)
        return CellSpec(self.cell, self.hidden_size, self.n_bits_w | (self.n_bits_a << 8) | ((0 if self.training else 1) << 16))

    def reset_parameters(self):
        raise AttributeError("the quantised model is built from a float model (QGRUQuant.from_float)")

    def sync_quant_buffers(self):
        for m in self.modules():
            if isinstance(m, _Quantizer):
                m.sync_buffers()

    @classmethod
    def from_float(cls, float_backbone, n_bits_w=8, n_bits_a=8):
        """Same construction order — hence the same RNG consumption — as Base_GRUQuantEnv.__init__ (quant_envs.py:139-171)."""
        H = float_backbone.hidden_size
        if getattr(float_backbone, "num_layers", 1) != 1 or H > 32:
            raise NotImplementedError("native QAT QGRU: num_layers=1, hidden_size <= 32 (quant_qgru_dpd_regr.sh uses 1 layer, H <= 30)")
        amp1 = float_backbone.cell == "qgru_amp1"
        # (1) create_pygru_model: GRUCell.__init__ (quant/modules/gru.py:9-30) draws two nn.Linear inits, then uniform(-1/sqrt(H), 1/sqrt(H))
        x2h, h2h = nn.Linear(4, 3 * H, bias=True), nn.Linear(H, 3 * H, bias=True)
        std = 1.0 / math.sqrt(H)
        for w in (x2h.weight, x2h.bias, h2h.weight, h2h.bias):
            nn.init.uniform_(w, -std, std)
        #     _reset_parameters (quant_envs.py:222-232): zero biases, per-gate orthogonal weights, per-gate xavier for x2h.weight
        for name, param in (("x2h.weight", x2h.weight), ("x2h.bias", x2h.bias), ("h2h.weight", h2h.weight), ("h2h.bias", h2h.bias)):
            num_gates = int(param.shape[0] / H)
            if "bias" in name:
                nn.init.constant_(param, 0)
            if "weight" in name:
                for i in range(num_gates):
                    nn.init.orthogonal_(param[i * H:(i + 1) * H, :])
            if "x2h.weight" in name:
                for i in range(num_gates):
                    nn.init.xavier_uniform_(param[i * H:(i + 1) * H, :])
        # (2) create_quantized_model: INT_Linear.__init__ runs nn.Linear.__init__ again (fresh RNG draws) and then adopts only the
        #     WEIGHT of the layer it replaces — the bias stays the fresh draw.  Traversal order: x2h, h2h, fc_out.
        bx = nn.Linear(4, 3 * H, bias=True).bias.detach().clone()
        bh = nn.Linear(H, 3 * H, bias=True).bias.detach().clone()
        bo = nn.Linear(H, 2, bias=True).bias.detach().clone()
        return cls(H, x2h.weight.detach().clone(), bx, h2h.weight.detach().clone(), bh,
                   float_backbone.fc_out.weight.detach().clone().cpu(), bo, n_bits_w, n_bits_a, amp1)

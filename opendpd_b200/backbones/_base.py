"""Shared host-side scaffolding of the native backbones (drop-in for reference backbones/*.py)."""
import math
import torch
from torch import nn
from ..flat import FlatParams
from ..functional import BackboneFn, CellSpec


class RNNParams(nn.Module):
    """Parameter container with torch.nn.GRU/LSTM's names, shapes and default init (so state_dicts interchange with
    the reference's nn.GRU-based backbones and the RNG stream is consumed identically) — but it is NOT an nn.GRU: it
    has no forward; the arithmetic runs in libodpd.so.  Ref: nn.RNNBase.__init__/reset_parameters
    (uniform(-1/sqrt(H), 1/sqrt(H)) over weight_ih, weight_hh, bias_ih, bias_hh in that order)."""

    MAX_LAYERS = 8      # csrc/wide.cu WLMAX

    def __init__(self, input_size, hidden_size, gates, num_layers=1, bias=True):
        super().__init__()
        if not 1 <= num_layers <= self.MAX_LAYERS:
            raise NotImplementedError(f"native RNN backbones implement num_layers 1..{self.MAX_LAYERS} (got {num_layers})")
        if not bias:
            raise NotImplementedError("native RNN containers assume bias=True (models.py:23)")
        self.input_size, self.hidden_size, self.num_layers, self.bias = input_size, hidden_size, num_layers, bias
        # nn.RNNBase registers, per layer k: weight_ih_l{k} (layer 0 reads the features, layer k > 0 the hidden sequence of layer
        # k-1), weight_hh_l{k}, bias_ih_l{k}, bias_hh_l{k} — in that order, which is also the flat layout the kernels read
        for k in range(num_layers):
            fin = input_size if k == 0 else hidden_size
            setattr(self, f"weight_ih_l{k}", nn.Parameter(torch.empty(gates * hidden_size, fin)))
            setattr(self, f"weight_hh_l{k}", nn.Parameter(torch.empty(gates * hidden_size, hidden_size)))
            setattr(self, f"bias_ih_l{k}", nn.Parameter(torch.empty(gates * hidden_size)))
            setattr(self, f"bias_hh_l{k}", nn.Parameter(torch.empty(gates * hidden_size)))
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.hidden_size) if self.hidden_size > 0 else 0
        for w in self.parameters():
            nn.init.uniform_(w, -stdv, stdv)


def gatewise_rnn_init(rnn, hidden_size, ih_key="weight_ih_l0"):
    """The reset_parameters idiom shared by gru.py:27-36, dgru.py:35-44, lstm.py, deltagru.py:42-51, qgru.py:36-45."""
    for name, param in rnn.named_parameters():
        num_gates = int(param.shape[0] / hidden_size)
        if "bias" in name:
            nn.init.constant_(param, 0)
        if "weight" in name:
            for i in range(num_gates):
                nn.init.orthogonal_(param[i * hidden_size:(i + 1) * hidden_size, :])
        if ih_key in name:
            for i in range(num_gates):
                nn.init.xavier_uniform_(param[i * hidden_size:(i + 1) * hidden_size, :])


def linear_xavier_zero(lin):
    for name, param in lin.named_parameters():
        if "weight" in name:
            nn.init.xavier_uniform_(param)
        if "bias" in name:
            nn.init.constant_(param, 0)


class NativeBackbone(nn.Module, FlatParams):
    """forward(x:(B,T,2), h_0) -> (B,T,2) through libodpd.so.  h_0 is accepted for signature compatibility and must be
    the all-zero state every reference call site passes (models.py:154-155)."""

    cell = None

    def _spec(self):
        # OdpdDims.K: DVRJANET's num_dvr_units; for the nn.GRU / nn.LSTM based backbones the number of stacked layers
        K = getattr(self, "num_dvr_units", 0)
        if self.cell in ("gru", "lstm", "dgru", "qgru", "qgru_amp1") and getattr(self, "num_layers", 1) > 1:
            K = self.num_layers
        return CellSpec(self.cell, getattr(self, "hidden_size", 0), K,
                        getattr(self, "thx", 0.0), getattr(self, "thh", 0.0), getattr(self, "time_chunks", None),
                        getattr(self, "time_warmup", None))

    def _stats_tensor(self, device):
        return None

    def _run(self, x, target, loss_count):
        flat, layout = self._flat_sync()
        params = [p for _, p in self.named_parameters()]
        return BackboneFn.apply(self._spec(), layout, flat, self._stats_tensor(x.device), target, loss_count, x, *params)

    def forward(self, x, h_0=None):
        out, _ = self._run(x, None, None)
        return out

    def forward_mse(self, x, target, loss_count=None):
        """Fused forward + nn.MSELoss()(out, target); loss_count = number of scalars the mean runs over
        (defaults to out.numel(); data-parallel callers pass the GLOBAL count). Returns (out, loss)."""
        return self._run(x, target, loss_count)

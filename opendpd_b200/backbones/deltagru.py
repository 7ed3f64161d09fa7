"""DeltaGRU and TRes-DeltaGRU backbones — drop-ins for reference backbones/deltagru.py (DeltaGRU :9-100, DeltaGRULayer
:103-276) and backbones/deltagru_tcnskip.py (DeltaGRU :11-126, DeltaGRULayer :129-304).

Kept from the reference: parameter names/shapes/init (rnn.weight_ih_l0 ... / rnn.x2h.weight ...), thx/thh attributes,
set_debug / get_temporal_sparsity (modules/paths.py:49-59 calls them) with the same four counters — counted exactly in
int64 on the device (the reference accumulates them in fp32 0-dim tensors, which loses integers above 2^24)."""
import numpy as np
import torch
from torch import nn
from ._base import NativeBackbone, gatewise_rnn_init, linear_xavier_zero


class _DeltaLayerParams(nn.Module):
    """rnn.* container of deltagru.py: DeltaGRULayer subclasses nn.GRU, whose __init__ calls the overridden
    reset_parameters (full-matrix orthogonal weights, zero biases, deltagru.py:142-147)."""

    def __init__(self, input_size, hidden_size, thx, thh):
        super().__init__()
        self.input_size, self.hidden_size, self.th_x, self.th_h = input_size, hidden_size, thx, thh
        self.debug = 1
        self.weight_ih_l0 = nn.Parameter(torch.empty(3 * hidden_size, input_size))
        self.weight_hh_l0 = nn.Parameter(torch.empty(3 * hidden_size, hidden_size))
        self.bias_ih_l0 = nn.Parameter(torch.empty(3 * hidden_size))
        self.bias_hh_l0 = nn.Parameter(torch.empty(3 * hidden_size))
        for name, param in self.named_parameters():
            if "weight" in name:
                nn.init.orthogonal_(param)
            elif "bias" in name:
                nn.init.constant_(param, 0)


class _TResLayerParams(nn.Module):
    """rnn.* container of deltagru_tcnskip.py:156-157: bias-free x2h / h2h Linears."""

    def __init__(self, input_size, hidden_size, thx, thh):
        super().__init__()
        self.input_size, self.hidden_size, self.th_x, self.th_h = input_size, hidden_size, thx, thh
        self.debug = 1
        self.x2h = nn.Linear(input_size, 3 * hidden_size, bias=False)
        self.h2h = nn.Linear(hidden_size, 3 * hidden_size, bias=False)


class _DeltaBase(NativeBackbone):
    def _init_stats(self):
        self._stats = None
        self._masks = None
        self.set_debug(1)

    def set_debug(self, value):
        """reference deltagru.py:33-40: also resets the sparsity counters."""
        setattr(self, "debug", value)
        if getattr(self, "_stats", None) is not None:
            self._stats.zero_()
        self.rnn.statistics = {"num_dx_zeros": 0, "num_dx_numel": 0, "num_dh_zeros": 0, "num_dh_numel": 0}

    def _stats_tensor(self, device):
        if self._stats is None or self._stats.device != device:
            self._stats = torch.zeros(4, dtype=torch.int64, device=device)
        return self._stats

    def _spec(self):
        spec = super()._spec()
        if getattr(self, "keep_masks", False):
            spec.keep_saved = self
        return spec

    def last_masks(self):
        """(mask_x, mask_h) keep-bitfields (B,T) uint64 of the last forward that saved activations; needs
        `self.keep_masks = True` before that forward.  Test/debug helper (reference analogue: none)."""
        saved, B, T = self._last_saved
        rows = saved.view(torch.int32).view(B, T, -1).cpu().numpy()
        return rows[..., -2].astype(np.uint32).astype(np.uint64), rows[..., -1].astype(np.uint32).astype(np.uint64)

    def raw_statistics(self):
        if self._stats is None:
            return [0, 0, 0, 0]
        return [int(v) for v in self._stats.cpu().tolist()]

    def get_temporal_sparsity(self):
        """reference deltagru.py:79-100 / deltagru_tcnskip.py:105-126."""
        st = self.raw_statistics()
        self.rnn.statistics = dict(num_dx_zeros=st[0], num_dx_numel=st[1], num_dh_zeros=st[2], num_dh_numel=st[3])
        out = {}
        if self.rnn.debug and st[1] > 0 and st[3] > 0:
            rnn_numel = sum(p.numel() for n, p in self.rnn.named_parameters() if "weight" in n)
            rnn_bias_numel = sum(p.numel() for n, p in self.rnn.named_parameters() if "bias" in n)
            fc_numel = self._fc_numel()
            tot_n, tot_z = st[1] + st[3], st[0] + st[2]
            out["SP_T_DX"] = float(st[0] / st[1])
            out["SP_T_DH"] = float(st[2] / st[3])
            out["SP_T_DV"] = float(tot_z / tot_n)
            out["HW_PARAM"] = float(fc_numel + rnn_numel * (1 - float(tot_z / tot_n)) + rnn_bias_numel)
        return out


class DeltaGRU(_DeltaBase):
    cell = "deltagru"

    def __init__(self, input_size, hidden_size, output_size, num_layers, thx=0, thh=0, bias=True):
        super().__init__()
        if num_layers != 1 or output_size != 2:
            raise NotImplementedError("native DeltaGRU: num_layers=1, 2 outputs")
        self.hidden_size, self.input_size, self.output_size, self.num_layers = hidden_size, 6, output_size, num_layers
        self.thh, self.thx, self.bias = thh, thx, bias
        self.rnn = _DeltaLayerParams(self.input_size, hidden_size, thx, thh)
        self.fc_out = nn.Linear(in_features=hidden_size, out_features=output_size, bias=True)
        self._init_stats()

    def reset_parameters(self):
        gatewise_rnn_init(self.rnn, self.hidden_size)
        linear_xavier_zero(self.fc_out)

    def _fc_numel(self):
        return sum(p.numel() for p in self.fc_out.parameters())


class TResDeltaGRU(_DeltaBase):
    """reference class name is also `DeltaGRU` (backbones/deltagru_tcnskip.py:11); exported under both names."""
    cell = "deltagru_tcnskip"

    def __init__(self, input_size, hidden_size, output_size, num_layers, thx=0, thh=0, bias=True):
        super().__init__()
        if num_layers != 1 or output_size != 2:
            raise NotImplementedError("native TRes-DeltaGRU: num_layers=1 (SURVEY App. A.10), 2 outputs")
        self.hidden_size, self.input_size, self.output_size, self.num_layers = hidden_size, 6, output_size, num_layers
        self.thh, self.thx, self.bias = thh, thx, bias
        self.rnn = _TResLayerParams(self.input_size, hidden_size, thx, thh)
        self.fc_out = nn.Linear(in_features=hidden_size, out_features=output_size, bias=False)
        self.tcn = nn.Sequential(
            nn.Conv1d(in_channels=2, out_channels=3, kernel_size=3, padding=16, stride=1, dilation=16, bias=False),
            nn.Hardswish(),
            nn.Conv1d(in_channels=3, out_channels=2, kernel_size=1, padding=0, stride=1, dilation=1, bias=False),
            nn.Hardswish(),
        )
        self._init_stats()

    def reset_parameters(self):
        for name, param in self.tcn.named_parameters():
            if "weight" in name:
                nn.init.xavier_uniform_(param)
        for name, param in self.rnn.named_parameters():
            num_gates = int(param.shape[0] / self.hidden_size)
            if "weight" in name:
                for i in range(num_gates):
                    nn.init.orthogonal_(param[i * self.hidden_size:(i + 1) * self.hidden_size, :])
            if "x2h.weight" in name:
                for i in range(num_gates):
                    nn.init.xavier_uniform_(param[i * self.hidden_size:(i + 1) * self.hidden_size, :])
        linear_xavier_zero(self.fc_out)

    def _fc_numel(self):
        return sum(p.numel() for p in self.fc_out.parameters()) + sum(p.numel() for p in self.tcn.parameters())

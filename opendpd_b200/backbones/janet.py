"""PGJANET / DVRJANET backbones — drop-ins for reference backbones/pgjanet.py (:5-84) and backbones/dvrjanet.py (:5-111)."""
import torch
from torch import nn
from ._base import NativeBackbone


class PGJANET(NativeBackbone):
    cell = "pgjanet"

    def __init__(self, hidden_size, output_size, bias=True, window_size=None):
        # window_size: models.py:111-114 passes it, the reference ctor rejects it; accepted and ignored here.
        super().__init__()
        if output_size != 2 or not bias:
            raise NotImplementedError("native PGJANET: bias=True, 2 outputs")
        self.hidden_size, self.output_size, self.bias = hidden_size, output_size, bias
        self.W_a = nn.Linear(hidden_size + 1, hidden_size, bias=bias)
        self.W_p1 = nn.Linear(hidden_size + 1, hidden_size, bias=bias)
        self.W_p2 = nn.Linear(hidden_size + 1, hidden_size, bias=bias)
        self.W_f = nn.Linear(hidden_size + hidden_size, hidden_size, bias=bias)
        self.W_g = nn.Linear(hidden_size + hidden_size, hidden_size, bias=bias)
        self.W_o = nn.Linear(hidden_size, output_size, bias=bias)

    def reset_parameters(self):
        for module in [self.W_a, self.W_p1, self.W_p2, self.W_f, self.W_g, self.W_o]:
            nn.init.xavier_uniform_(module.weight)
            if module.bias is not None:
                nn.init.constant_(module.bias, 0)


class DVRJANET(NativeBackbone):
    cell = "dvrjanet"

    def __init__(self, hidden_size, output_size, num_dvr_units=4, bias=True):
        super().__init__()
        if output_size != 2 or not bias:
            raise NotImplementedError("native DVRJANET: bias=True, 2 outputs")
        self.hidden_size, self.output_size, self.num_dvr_units, self.bias = hidden_size, output_size, num_dvr_units, bias
        self.W_ph = nn.Linear(hidden_size, hidden_size, bias=False)
        setattr(self, "W_pθ", nn.Linear(1, hidden_size, bias=False))
        self.W_ah = nn.Linear(hidden_size, hidden_size, bias=False)
        self.W_ax = nn.Linear(1, hidden_size, bias=False)
        self.cs = nn.Parameter(torch.randn(num_dvr_units))   # not touched by reset_parameters (dvrjanet.py:21,104-111)
        self.W_f = nn.Linear(hidden_size, hidden_size, bias=bias)
        self.W_ccos = nn.Linear(hidden_size + hidden_size, hidden_size, bias=bias)
        self.W_csin = nn.Linear(hidden_size + hidden_size, hidden_size, bias=bias)
        self.W_o1 = nn.Linear(hidden_size, 1, bias=bias)
        self.W_o2 = nn.Linear(hidden_size, 1, bias=bias)

    def reset_parameters(self):
        for module in [self.W_ph, getattr(self, "W_pθ"), self.W_ah, self.W_ax, self.W_f, self.W_ccos, self.W_csin, self.W_o1, self.W_o2]:
            nn.init.xavier_uniform_(module.weight)
            if module.bias is not None:
                nn.init.constant_(module.bias, 0)

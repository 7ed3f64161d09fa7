"""BOJANET backbone — drop-in for reference backbones/bojanet.py (ctor :5-28, forward :54-106, reset_parameters :108-134).

The constructor draws the eight nn.Linear default inits in the reference's order and then — like the reference's — calls
reset_parameters() itself; CoreModel calls it once more (models.py:143-148).  Same RNG stream, same initial weights."""
from torch import nn
from ._base import NativeBackbone


class BOJANET(NativeBackbone):
    cell = "bojanet"

    def __init__(self, hidden_size, output_size, bias=True):
        super().__init__()
        if output_size != 2 or not bias:
            raise NotImplementedError("native BOJANET: 2 outputs, bias=True (models.py:86-90)")
        self.window_size, self.num_vd_units = 16, 6
        if not 1 <= hidden_size <= 3 * self.num_vd_units:
            raise NotImplementedError(f"native BOJANET: hidden_size 1..18 — the reference's phase-rotation block (bojanet.py:41-52) covers "
                                      f"at most 3 x 6 units and fails beyond (got {hidden_size})")
        self.hidden_size, self.output_size, self.bias = hidden_size, output_size, bias
        P, M = self.num_vd_units, self.window_size
        self.fir_I = nn.Linear(M, P, bias=False)
        self.fir_Q = nn.Linear(M, P, bias=False)
        self.W_fi = nn.Linear(P * 2, hidden_size, bias=bias)
        self.W_fh = nn.Linear(hidden_size, hidden_size, bias=False)
        self.W_gi = nn.Linear(P * 2, hidden_size, bias=bias)
        self.W_gh = nn.Linear(hidden_size, hidden_size, bias=False)
        self.W_out_I = nn.Linear(hidden_size, 1, bias=bias)
        self.W_out_Q = nn.Linear(hidden_size, 1, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        for lin in (self.fir_I, self.fir_Q):
            nn.init.xavier_uniform_(lin.weight, gain=0.1)
        for lin in (self.W_fi, self.W_gi):
            nn.init.xavier_uniform_(lin.weight, gain=1.0)
            nn.init.constant_(lin.bias, 0)
        for lin in (self.W_fh, self.W_gh):
            nn.init.orthogonal_(lin.weight, gain=1.0)
        for lin in (self.W_out_I, self.W_out_Q):
            nn.init.xavier_uniform_(lin.weight, gain=1.0)
            nn.init.constant_(lin.bias, 0)

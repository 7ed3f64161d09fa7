"""APNRRU backbone — drop-in for reference backbones/apnrru.py (RRU :5-31, APNRRU ctor :34-51, forward :52-135, reset_parameters :137-152).

Same construction order, hence the same RNG stream: fir_I, fir_Q, the RRU cell (W_u, W_h, C ~ U(0,1), Z = 0), the two output layers.
The reference's reset_parameters re-draws the FIR and RRU weights and then raises AttributeError on a non-existent `self.output_layer`
(:149); CoreModel swallows it (models.py:143-148), so the model starts from exactly that partially re-initialised state — reproduced here."""
import torch
from torch import nn
from ._base import NativeBackbone


class RRU(nn.Module):
    """Parameter container of the cell (apnrru.py:5-20); the arithmetic runs in csrc/apnrru.cu."""

    def __init__(self, hidden_size, window_size, bias=True):
        super().__init__()
        self.hidden_size, self.hidden_size_A, self.num_fir_filters, self.hidden_node = hidden_size, 3, 3, 16
        S = hidden_size * 2 + self.hidden_size_A
        self.W_u = nn.Linear(S + self.num_fir_filters * 2 + 2, self.hidden_node, bias=bias)
        self.W_h = nn.Linear(self.hidden_node, S, bias=bias)
        self.C = nn.Parameter(torch.rand(1))
        self.Z = nn.Parameter(torch.zeros(1, S))


class APNRRU(NativeBackbone):
    cell = "apnrru"

    def __init__(self, hidden_size, bias=True):
        super().__init__()
        if not bias:
            raise NotImplementedError("native APNRRU: bias=True (models.py:82-85)")
        if not 1 <= hidden_size <= 14:
            raise NotImplementedError(f"native APNRRU: hidden_size 1..14 (one warp lane per state value: 2H+3 <= 31; got {hidden_size})")
        self.hidden_size, self.hidden_size_A, self.window_size, self.num_fir_filters, self.hidden_node = hidden_size, 3, 16, 3, 16
        self.fir_I = nn.Linear(self.window_size, self.num_fir_filters, bias=False)
        self.fir_Q = nn.Linear(self.window_size, self.num_fir_filters, bias=False)
        self.rru = RRU(hidden_size, self.window_size, bias)
        self.output_layer_I = nn.Linear(hidden_size, 1, bias=False)
        self.output_layer_Q = nn.Linear(hidden_size, 1, bias=False)

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.fir_I.weight)
        nn.init.xavier_uniform_(self.fir_Q.weight)
        for lin in (self.rru.W_u, self.rru.W_h):
            nn.init.xavier_uniform_(lin.weight)
            nn.init.constant_(lin.bias, 0)
        # apnrru.py:149 touches `self.output_layer`, which does not exist: the reference stops here with an AttributeError that
        # CoreModel catches — the output layers keep their default init
        raise AttributeError("'APNRRU' object has no attribute 'output_layer'")

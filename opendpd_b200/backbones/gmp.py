"""GMP backbone — drop-in for reference backbones/gmp.py (:6-51): generalized memory polynomial, memory 11, degree 5."""
import torch
from torch import nn
from ._base import NativeBackbone


class GMP(NativeBackbone):
    cell = "gmp"

    def __init__(self, memory_length=11, degree=5):
        super().__init__()
        if memory_length != 11 or degree != 5:
            raise NotImplementedError("native GMP: memory_length=11, degree=5 (the only configuration models.py:26-28 builds)")
        self.memory_length, self.degree = memory_length, degree
        self.W = 1 + (degree - 1) * memory_length
        self.Weight = nn.Parameter(torch.Tensor(1, memory_length * self.W))

    def reset_parameters(self):
        for name, param in self.named_parameters():
            if "W" in name:
                nn.init.xavier_uniform_(param)

"""Native drop-ins for the reference's backbones/ package (same class names, ctor signatures, parameter names)."""
from .gru import GRU
from .dgru import DGRU
from .qgru import QGRU, QGRUAmp1

__all__ = ["GRU", "DGRU", "QGRU", "QGRUAmp1"]

"""reference module path backbones/gru.py, class `GRU` -> the native backbone (opendpd_b200.backbones.GRU)."""
from opendpd_b200.backbones import GRU as GRU  # noqa: F401

"""reference module path backbones/deltagru.py, class `DeltaGRU` -> the native backbone (opendpd_b200.backbones.DeltaGRU)."""
from opendpd_b200.backbones import DeltaGRU as DeltaGRU  # noqa: F401

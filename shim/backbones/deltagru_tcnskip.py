"""reference module path backbones/deltagru_tcnskip.py, class `DeltaGRU` -> the native backbone (opendpd_b200.backbones.TResDeltaGRU)."""
from opendpd_b200.backbones import TResDeltaGRU as DeltaGRU  # noqa: F401

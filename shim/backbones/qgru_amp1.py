"""reference module path backbones/qgru_amp1.py, class `QGRU` -> the native backbone (opendpd_b200.backbones.QGRUAmp1)."""
from opendpd_b200.backbones import QGRUAmp1 as QGRU  # noqa: F401

"""reference module path backbones/qgru.py, class `QGRU` -> the native backbone (opendpd_b200.backbones.QGRU)."""
from opendpd_b200.backbones import QGRU as QGRU  # noqa: F401

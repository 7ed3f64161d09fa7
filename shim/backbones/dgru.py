"""reference module path backbones/dgru.py, class `DGRU` -> the native backbone (opendpd_b200.backbones.DGRU)."""
from opendpd_b200.backbones import DGRU as DGRU  # noqa: F401

"""reference module path backbones/bojanet.py, class `BOJANET` -> the native backbone (opendpd_b200.backbones.BOJANET)."""
from opendpd_b200.backbones import BOJANET  # noqa: F401

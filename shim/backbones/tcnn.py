"""reference module path backbones/tcnn.py, class `TCNN` -> the native backbone (opendpd_b200.backbones.TCNN)."""
from opendpd_b200.backbones import TCNN  # noqa: F401

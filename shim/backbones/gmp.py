"""reference module path backbones/gmp.py, class `GMP` -> the native backbone (opendpd_b200.backbones.GMP)."""
from opendpd_b200.backbones import GMP as GMP  # noqa: F401

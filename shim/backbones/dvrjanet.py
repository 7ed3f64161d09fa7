"""reference module path backbones/dvrjanet.py, class `DVRJANET` -> the native backbone (opendpd_b200.backbones.DVRJANET)."""
from opendpd_b200.backbones import DVRJANET as DVRJANET  # noqa: F401

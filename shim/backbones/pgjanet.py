"""reference module path backbones/pgjanet.py, class `PGJANET` -> the native backbone (opendpd_b200.backbones.PGJANET)."""
from opendpd_b200.backbones import PGJANET as PGJANET  # noqa: F401

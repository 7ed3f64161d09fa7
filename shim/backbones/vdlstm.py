"""reference module path backbones/vdlstm.py, class `VDLSTM` -> the native backbone (opendpd_b200.backbones.VDLSTM)."""
from opendpd_b200.backbones import VDLSTM  # noqa: F401

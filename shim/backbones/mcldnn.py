"""reference module path backbones/mcldnn.py, class `MCLDNN` -> the native backbone (opendpd_b200.backbones.MCLDNN)."""
from opendpd_b200.backbones import MCLDNN  # noqa: F401

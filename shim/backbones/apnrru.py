"""reference module path backbones/apnrru.py, class `APNRRU` -> the native backbone (opendpd_b200.backbones.APNRRU)."""
from opendpd_b200.backbones import APNRRU  # noqa: F401

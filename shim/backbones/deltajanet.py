"""reference module path backbones/deltajanet.py, class `DeltaJANET` -> the native backbone (opendpd_b200.backbones.DeltaJANET)."""
from opendpd_b200.backbones import DeltaJANET  # noqa: F401

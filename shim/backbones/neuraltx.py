"""reference module path backbones/neuraltx.py, class `NeuralTX` -> the native backbone (opendpd_b200.backbones.NeuralTX)."""
from opendpd_b200.backbones import NeuralTX  # noqa: F401

"""reference module path backbones/lstm.py, class `LSTM` -> the native backbone (opendpd_b200.backbones.LSTM)."""
from opendpd_b200.backbones import LSTM as LSTM  # noqa: F401

"""reference module path backbones/rvtdcnn.py, class `RVTDCNN` -> the native backbone (opendpd_b200.backbones.RVTDCNN)."""
from opendpd_b200.backbones import RVTDCNN  # noqa: F401

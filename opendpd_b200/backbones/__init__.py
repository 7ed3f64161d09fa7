"""Native drop-ins for the reference's backbones/ package (same class names, ctor signatures, parameter names)."""
from .gru import GRU
from .dgru import DGRU
from .qgru import QGRU, QGRUAmp1
from .lstm import LSTM
from .vdlstm import VDLSTM
from .deltagru import DeltaGRU, TResDeltaGRU
from .janet import PGJANET, DVRJANET
from .gmp import GMP
from .rvtdcnn import RVTDCNN
from .bojanet import BOJANET
from .tcnn import TCNN, NeuralTX
from .apnrru import APNRRU
from .mcldnn import MCLDNN
from .deltajanet import DeltaJANET

__all__ = ["GRU", "DGRU", "QGRU", "QGRUAmp1", "LSTM", "VDLSTM", "DeltaGRU", "TResDeltaGRU", "PGJANET", "DVRJANET", "GMP", "RVTDCNN", "BOJANET", "TCNN", "NeuralTX", "APNRRU", "MCLDNN", "DeltaJANET"]

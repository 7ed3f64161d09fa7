"""QGRU float host — drop-in for reference backbones/qgru.py (features I,Q,|x|^2,|x|^4 :61-66) and
backbones/qgru_amp1.py (features I,Q,|x|,|x|^3 :61-70).  reset_parameters reproduces the reference quirk: it
initialises rnn and fc_out and then raises AttributeError on the non-existent fc_hid (qgru.py:52-56), which
models.CoreModel swallows (models.py:144-148)."""
from torch import nn
from ._base import NativeBackbone, RNNParams, gatewise_rnn_init, linear_xavier_zero


class QGRU(NativeBackbone):
    cell = "qgru"

    def __init__(self, hidden_size, output_size, num_layers, bidirectional=False, batch_first=True, bias=True):
        super().__init__()
        if bidirectional or not batch_first or output_size != 2 or not bias:
            raise NotImplementedError("native QGRU: unidirectional, batch_first, bias, 2 outputs")
        self.hidden_size, self.input_size, self.output_size = hidden_size, 4, output_size
        self.num_layers, self.bidirectional, self.batch_first, self.bias = num_layers, bidirectional, batch_first, bias
        self.rnn = RNNParams(self.input_size, hidden_size, gates=3, num_layers=num_layers, bias=bias)
        self.fc_out = nn.Linear(in_features=hidden_size, out_features=output_size, bias=bias)

    def reset_parameters(self):
        gatewise_rnn_init(self.rnn, self.hidden_size)
        linear_xavier_zero(self.fc_out)
        raise AttributeError("'QGRU' object has no attribute 'fc_hid'")  # reference behaviour, qgru.py:52


class QGRUAmp1(QGRU):
    cell = "qgru_amp1"

"""LSTM backbone — drop-in for reference backbones/lstm.py (ctor :5-25, reset_parameters :27-43, forward :45-48: (h0,c0)=(h_0,h_0)=0)."""
from torch import nn
from ._base import NativeBackbone, RNNParams, gatewise_rnn_init, linear_xavier_zero


class LSTM(NativeBackbone):
    cell = "lstm"

    def __init__(self, input_size, hidden_size, output_size, num_layers, bidirectional=False, batch_first=True, bias=True):
        super().__init__()
        if bidirectional or not batch_first or input_size != 2 or output_size != 2:
            raise NotImplementedError("native LSTM: unidirectional, batch_first, I/Q in and out (models.py:21-23)")
        self.hidden_size, self.input_size, self.output_size = hidden_size, input_size, output_size
        self.num_layers, self.bidirectional, self.batch_first, self.bias = num_layers, bidirectional, batch_first, bias
        self.rnn = RNNParams(input_size, hidden_size, gates=4, num_layers=num_layers, bias=bias)
        self.fc_out = nn.Linear(in_features=hidden_size, out_features=output_size, bias=True)

    def reset_parameters(self):
        gatewise_rnn_init(self.rnn, self.hidden_size)
        linear_xavier_zero(self.fc_out)

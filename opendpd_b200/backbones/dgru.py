"""DGRU backbone — drop-in for reference backbones/dgru.py (ctor :9-33, reset_parameters :35-57, forward :59-74)."""
from torch import nn
from ._base import NativeBackbone, RNNParams, gatewise_rnn_init, linear_xavier_zero


class DGRU(NativeBackbone):
    cell = "dgru"

    def __init__(self, hidden_size, output_size, num_layers, bidirectional=False, batch_first=True, bias=True):
        super().__init__()
        if bidirectional or not batch_first or output_size != 2 or not bias:
            raise NotImplementedError("native DGRU: unidirectional, batch_first, bias, 2 outputs (models.py:21-23)")
        self.hidden_size, self.input_size, self.output_size = hidden_size, 6, output_size
        self.num_layers, self.bidirectional, self.batch_first, self.bias = num_layers, bidirectional, batch_first, bias
        self.rnn = RNNParams(self.input_size, hidden_size, gates=3, num_layers=num_layers, bias=bias)
        self.fc_out = nn.Linear(in_features=hidden_size + self.input_size, out_features=output_size, bias=bias)
        self.fc_hid = nn.Linear(in_features=hidden_size, out_features=hidden_size, bias=bias)

    def reset_parameters(self):
        gatewise_rnn_init(self.rnn, self.hidden_size)
        linear_xavier_zero(self.fc_out)
        for name, param in self.fc_hid.named_parameters():
            if "weight" in name:
                nn.init.kaiming_uniform_(param)
            if "bias" in name:
                nn.init.constant_(param, 0)

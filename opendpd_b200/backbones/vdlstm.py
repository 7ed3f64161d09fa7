"""VDLSTM backbone — drop-in for reference backbones/vdlstm.py (ctor :5-41, forward :58-82, reset_parameters :84-111): an LSTM over a
4-tap window of the amplitude |x| (wrap-around padding with the last 3 samples of the frame) whose state scales the window's cos / sin
through two linear maps (fc_lambda_1/2: H -> 4) before fc_out (8 -> 2).  Arithmetic: csrc/lstm.cu (VD variant)."""
from torch import nn
from ._base import NativeBackbone, RNNParams, gatewise_rnn_init, linear_xavier_zero


class VDLSTM(NativeBackbone):
    cell = "vdlstm"

    def __init__(self, input_size, hidden_size, output_size, num_layers, window_length=4, stride=1, bidirectional=False, batch_first=True,
                 bias=True):
        super().__init__()
        if bidirectional or not batch_first or output_size != 2 or window_length != 4 or stride != 1:
            raise NotImplementedError("native VDLSTM: unidirectional, batch_first, window_length=4, stride=1, 2 outputs (vdlstm.py defaults; "
                                      "models.py:71-79 never passes others)")
        if num_layers != 1 or hidden_size > 32:
            raise NotImplementedError("native VDLSTM: num_layers=1, hidden_size <= 32 (the layered kernels cover GRU/LSTM/DGRU/QGRU only)")
        self.hidden_size, self.input_size, self.output_size = hidden_size, window_length, output_size      # vdlstm.py:19: input_size = window
        self.num_layers, self.bidirectional, self.batch_first, self.bias = num_layers, bidirectional, batch_first, bias
        self.window_length, self.stride, self.pad_size = window_length, stride, window_length - 1
        self.rnn = RNNParams(window_length, hidden_size, gates=4, num_layers=num_layers, bias=True)
        self.fc_lambda_1 = nn.Linear(in_features=hidden_size, out_features=window_length, bias=True)
        self.fc_lambda_2 = nn.Linear(in_features=hidden_size, out_features=window_length, bias=True)
        self.fc_out = nn.Linear(in_features=2 * window_length, out_features=2, bias=True)

    def reset_parameters(self):
        gatewise_rnn_init(self.rnn, self.hidden_size)
        linear_xavier_zero(self.fc_lambda_1)
        linear_xavier_zero(self.fc_lambda_2)
        linear_xavier_zero(self.fc_out)

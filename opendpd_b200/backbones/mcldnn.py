"""MCLDNN backbone — drop-in for reference backbones/mcldnn.py (ctor :9-29, reset_parameters :31-37, forward :83-113).

Parameter containers with the reference's names, shapes and construction order (conv2d_1, conv1d, conv2d_2, an nn.LSTM-shaped block of
width 8, fc_out, fc_out_2); the constructor calls reset_parameters() itself like the reference's, and CoreModel calls it again."""
from torch import nn
from ._base import NativeBackbone, RNNParams


class MCLDNN(NativeBackbone):
    cell = "mcldnn"

    def __init__(self, hidden_size=8):
        super().__init__()
        if not 1 <= hidden_size <= 12:
            raise NotImplementedError(f"native MCLDNN: hidden_size (conv channels) 1..12 (got {hidden_size})")
        self.memory_length, self.order = 5, 3
        self.input_height, self.input_width = 2 + self.order, self.memory_length
        self.channels = self.hidden_size = hidden_size            # hidden_size: what the C ABI calls H (the LSTM inside is always 8 wide)
        self.kernel_size = 3
        self.conv2d_1 = nn.Conv2d(1, self.channels, kernel_size=self.kernel_size, padding=1)
        self.conv1d = nn.Conv1d(self.input_height, self.input_height * self.channels, kernel_size=self.kernel_size, padding=1,
                                groups=self.input_height)
        self.conv2d_2 = nn.Conv2d(2 * self.input_height, 1, kernel_size=self.kernel_size, padding=1)
        self.lstm = RNNParams(self.channels * self.memory_length, 8, gates=4, num_layers=1, bias=True)     # nn.LSTM's names, shapes, init
        self.fc_out = nn.Linear(in_features=8, out_features=16)
        self.fc_out_2 = nn.Linear(in_features=16, out_features=2)
        self.reset_parameters()

    def reset_parameters(self):
        for name, param in self.named_parameters():
            if "weight" in name:
                nn.init.xavier_uniform_(param)
            elif "bias" in name:
                nn.init.constant_(param, 0)

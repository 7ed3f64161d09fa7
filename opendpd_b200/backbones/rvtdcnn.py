"""RVTDCNN backbone — drop-in for reference backbones/rvtdcnn.py (ctor :10-33, forward :35-62; no reset_parameters: the reference
keeps torch's default Conv2d / Linear init, and CoreModel swallows the missing method, models.py:143-148)."""
from torch import nn
from ._base import NativeBackbone


class RVTDCNN(NativeBackbone):
    cell = "rvtdcnn"

    def __init__(self, window_size=4, out_channels=3, kernel_size=3, stride=1, padding=(1, 0), dilation=1, fc_hid_size=6):
        super().__init__()
        if (window_size, out_channels, kernel_size, stride, tuple(padding), dilation) != (4, 3, 3, 1, (1, 0), 1):
            raise NotImplementedError("native RVTDCNN: the reference's defaults (window 4, 3 channels, 3x3, stride 1, padding (1,0)); "
                                      "models.py:80-81 only ever passes fc_hid_size")
        if not 1 <= fc_hid_size <= 64:
            raise NotImplementedError(f"native RVTDCNN: fc_hid_size 1..64 (got {fc_hid_size})")
        self.out_channels, self.window_size, self.stride = out_channels, window_size, stride
        self.feature_size_new = 3
        self.fc_in_features = self.out_channels * self.feature_size_new * self.window_size
        self.fc_hid_size = self.hidden_size = fc_hid_size           # hidden_size: what the C ABI calls H
        # parameter containers with the reference's names, shapes and (default) init, created in the reference's order
        self.Conv2d = nn.Conv2d(in_channels=1, out_channels=out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                                dilation=dilation, bias=True, padding_mode="zeros")
        self.fc_hid = nn.Linear(in_features=self.fc_in_features, out_features=fc_hid_size, bias=True)
        self.fc_out = nn.Linear(in_features=self.fc_hid_size, out_features=2, bias=True)

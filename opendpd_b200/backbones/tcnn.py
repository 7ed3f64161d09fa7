"""TCNN and NeuralTX backbones — drop-ins for reference backbones/tcnn.py (ctor :5-32, forward :83-97; no reset_parameters: torch's default
Conv1d init stays) and backbones/neuraltx.py (ctor :5-39, reset_parameters :40-55, forward :107-124).

The `network` Sequential, `conv_I` / `conv_Q` and `IQ_match` exist as PARAMETER CONTAINERS with the reference's names, shapes and
construction order (same RNG stream, same state_dict keys); their forward is never called — the arithmetic runs in csrc/tcnn.cu."""
from torch import nn
from ._base import NativeBackbone


def _tcn_stack(in_channels, hidden_channels, out_channels=2, kernel_size=5, dilation=1, stride=1):
    C, k, d = hidden_channels, kernel_size, dilation
    return nn.Sequential(
        nn.Conv1d(in_channels=in_channels, out_channels=C, kernel_size=1),
        nn.Hardswish(),
        nn.Conv1d(C, C, k, stride=stride, padding=(k - 3) * d, dilation=d, groups=C, bias=False),
        nn.Hardswish(),
        nn.Conv1d(C, C, k, stride=stride, padding=(k - 3) * d * 2, dilation=d * 2, groups=C, bias=False),
        nn.Hardswish(),
        nn.Conv1d(C, C, k, stride=stride, padding=(k - 3) * d * 4, dilation=d * 4, groups=C, bias=False),
        nn.Hardswish(),
        nn.Conv1d(C, C, k, stride=stride, padding=(k - 3) * d * 8, dilation=d * 8, groups=C, bias=False),
        nn.Hardswish(),
        nn.Conv1d(C, out_channels, kernel_size=1, bias=False),
    )


def _check_channels(name, hidden_channels):
    if not 1 <= hidden_channels <= 64:
        raise NotImplementedError(f"native {name}: hidden_channels 1..64 (got {hidden_channels})")


class TCNN(NativeBackbone):
    cell = "tcnn"

    def __init__(self, hidden_channels):
        super().__init__()
        _check_channels("TCNN", hidden_channels)
        self.in_channels, self.hidden_channels, self.out_channels = 6, hidden_channels, 2
        self.hidden_size = hidden_channels                      # what the C ABI calls H
        self.kernel_size, self.dilation, self.stride = 5, 1, 1
        self.network = _tcn_stack(self.in_channels, hidden_channels)


class NeuralTX(NativeBackbone):
    cell = "neuraltx"

    def __init__(self, hidden_channels):
        super().__init__()
        _check_channels("NeuralTX", hidden_channels)
        self.in_channels, self.hidden_channels, self.out_channels = 4, hidden_channels, 2
        self.hidden_size = hidden_channels
        self.kernel_size, self.dilation, self.stride, self.window_size, self.bias = 5, 1, 1, 5, False
        self.conv_I = nn.Conv1d(in_channels=1, out_channels=1, kernel_size=self.window_size, bias=False, padding=2)
        self.conv_Q = nn.Conv1d(in_channels=1, out_channels=1, kernel_size=self.window_size, bias=False, padding=2)
        self.network = _tcn_stack(self.in_channels, hidden_channels)
        self.IQ_match = nn.Linear(in_features=2, out_features=self.out_channels, bias=False)
        self.reset_parameters()

    def reset_parameters(self):
        # neuraltx.py:40-55: the loop over [self.network] tests the Sequential itself for a `weight` attribute and therefore touches
        # nothing — the stack keeps torch's default init; only the FIR kernels and IQ_match are re-drawn
        for conv in (self.conv_I, self.conv_Q):
            nn.init.xavier_uniform_(conv.weight, gain=0.1)
        nn.init.xavier_uniform_(self.IQ_match.weight, gain=1.0)

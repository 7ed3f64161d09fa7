"""DeltaJANET backbone — drop-in for reference backbones/deltajanet.py (DeltaJANET :10-60, DeltaJANETLayer :63-274).

RNG order of the reference, reproduced: the layer allocates its four tensors uninitialised and runs its own reset_parameters (orthogonal
weights — the full (2H,6) and (2H,H) matrices —, zero biases, :124-129); then fc_out draws nn.Linear's default init; CoreModel finally calls
the outer reset_parameters (per-gate orthogonal, per-gate xavier for weight_ih_l0, xavier + zero bias for fc_out, :31-47).

Like the reference, the layer is built with thx = thh = 0 whatever the constructor receives (:22-26): `thx` / `thh` are stored and ignored.
The reference's outer class has no get_temporal_sparsity()/set_debug(), so its own logging (modules/paths.py:56-59) fails for this backbone;
they are provided here (the counters of a threshold-free delta cell: only exact-zero deltas count) so that the training flow runs."""
import torch
from torch import nn
from ._base import NativeBackbone


class DeltaJANETLayer(nn.Module):
    """Parameter container of the layer (deltajanet.py:63-106)."""

    def __init__(self, input_size=6, hidden_size=256, num_layers=1, thx=0.1, thh=0):
        super().__init__()
        if num_layers != 1:
            raise NotImplementedError("native DeltaJANET: num_layers=1")
        self.input_size, self.hidden_size, self.num_layers, self.th_x, self.th_h = input_size, hidden_size, num_layers, thx, thh
        self.weight_ih_l0 = nn.Parameter(torch.empty(2 * hidden_size, input_size))
        self.weight_hh_l0 = nn.Parameter(torch.empty(2 * hidden_size, hidden_size))
        self.bias_ih_l0 = nn.Parameter(torch.empty(2 * hidden_size))
        self.bias_hh_l0 = nn.Parameter(torch.empty(2 * hidden_size))
        self.reset_parameters()

    def reset_parameters(self):
        for name, param in self.named_parameters():
            if "weight" in name:
                nn.init.orthogonal_(param)
            elif "bias" in name:
                nn.init.constant_(param, 0)


class DeltaJANET(NativeBackbone):
    cell = "deltajanet"

    def __init__(self, input_size, hidden_size, output_size, num_layers, thx=0, thh=0, bias=True):
        super().__init__()
        if num_layers != 1 or output_size != 2 or input_size != 6:
            raise NotImplementedError("native DeltaJANET: 6 features in, 2 outputs, num_layers=1 (models.py:100-108)")
        if not 1 <= hidden_size <= 16:
            raise NotImplementedError(f"native DeltaJANET: hidden_size 1..16 (got {hidden_size})")
        self.hidden_size, self.input_size, self.output_size, self.num_layers = hidden_size, input_size, output_size, num_layers
        self.thh, self.thx, self.bias = thh, thx, bias                  # stored, never used: deltajanet.py:22-26
        self.rnn = DeltaJANETLayer(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers, thx=0, thh=0)
        self.fc_out = nn.Linear(in_features=hidden_size, out_features=output_size, bias=True)
        self.debug = 1

    def _spec(self):
        spec = super()._spec()
        spec.thx = spec.thh = 0.0
        return spec

    def reset_parameters(self):
        H = self.hidden_size
        for name, param in self.rnn.named_parameters():
            num_gates = int(param.shape[0] / H)
            if "bias" in name:
                nn.init.constant_(param, 0)
            if "weight" in name:
                for i in range(num_gates):
                    nn.init.orthogonal_(param[i * H:(i + 1) * H, :])
            if "weight_ih_l0" in name:
                for i in range(num_gates):
                    nn.init.xavier_uniform_(param[i * H:(i + 1) * H, :])
        for name, param in self.fc_out.named_parameters():
            if "weight" in name:
                nn.init.xavier_uniform_(param)
            if "bias" in name:
                nn.init.constant_(param, 0)

    def set_debug(self, value):
        self.debug = value

    def get_temporal_sparsity(self):
        """With both thresholds fixed at zero nothing is ever skipped; reported as zero temporal sparsity (the reference cannot report
        anything here: its outer class lacks this method)."""
        return {"SP_T_DX": 0.0, "SP_T_DH": 0.0, "SP_T_DV": 0.0}

"""CoreModel / CascadedModel — mirror of reference models.py:10-176 that instantiates the native backbones.

Same constructor signature, attributes and parameter names (`backbone.*`), so `project.py`/`steps/*.py` use it
unchanged.  Every backbone the reference's CoreModel can build (models.py:26-141) has a native class here."""
import torch
from torch import nn

NATIVE_BACKBONES = ("gmp", "gru", "dgru", "qgru", "qgru_amp1", "lstm", "vdlstm", "deltagru", "deltagru_tcnskip", "pgjanet", "dvrjanet", "rvtdcnn", "bojanet", "tcnn", "neuraltx", "apnrru", "mcldnn", "deltajanet")


class CoreModel(nn.Module):
    def __init__(self, input_size, hidden_size, num_layers, backbone_type, window_size=None, num_dvr_units=None, thx=0, thh=0):
        super().__init__()
        self.output_size = 2
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        self.backbone_type, self.thx, self.thh = backbone_type, thx, thh
        self.window_size, self.num_dvr_units = window_size, num_dvr_units
        self.batch_first, self.bidirectional, self.bias = True, False, True
        from . import backbones as bb
        kw = dict(hidden_size=hidden_size, output_size=2, num_layers=num_layers, bidirectional=False, batch_first=True, bias=True)
        if backbone_type == "gru":
            self.backbone = bb.GRU(input_size=input_size, **kw)
        elif backbone_type == "dgru":
            self.backbone = bb.DGRU(**kw)
        elif backbone_type == "qgru":
            self.backbone = bb.QGRU(**kw)
        elif backbone_type == "qgru_amp1":
            self.backbone = bb.QGRUAmp1(**kw)
        elif backbone_type == "lstm":
            self.backbone = bb.LSTM(input_size=input_size, **kw)
        elif backbone_type == "vdlstm":
            self.backbone = bb.VDLSTM(input_size=input_size, **kw)
        elif backbone_type == "deltagru":
            self.backbone = bb.DeltaGRU(input_size=6, hidden_size=hidden_size, output_size=2, num_layers=num_layers,
                                        thx=thx, thh=thh, bias=True)
        elif backbone_type == "deltagru_tcnskip":
            self.backbone = bb.TResDeltaGRU(input_size=6, hidden_size=hidden_size, output_size=2, num_layers=num_layers,
                                            thx=thx, thh=thh, bias=True)
        elif backbone_type == "pgjanet":
            # reference models.py:111-114 passes window_size= which pgjanet.py:6 rejects (TypeError); we accept and ignore it
            self.backbone = bb.PGJANET(hidden_size=hidden_size, output_size=2, bias=True)
        elif backbone_type == "dvrjanet":
            self.backbone = bb.DVRJANET(hidden_size=hidden_size, output_size=2, num_dvr_units=num_dvr_units, bias=True)
        elif backbone_type == "gmp":
            self.backbone = bb.GMP()
        elif backbone_type == "deltajanet":
            self.backbone = bb.DeltaJANET(input_size=6, hidden_size=hidden_size, output_size=2, num_layers=num_layers,
                                          thx=thx, thh=thh, bias=True)                          # models.py:100-108
        elif backbone_type == "mcldnn":
            self.backbone = bb.MCLDNN(hidden_size=hidden_size)                                  # models.py:136-138
        elif backbone_type == "apnrru":
            self.backbone = bb.APNRRU(hidden_size=hidden_size, bias=True)                       # models.py:82-85
        elif backbone_type == "bojanet":
            self.backbone = bb.BOJANET(hidden_size=hidden_size, output_size=2, bias=True)      # models.py:86-90
        elif backbone_type == "tcnn":
            self.backbone = bb.TCNN(hidden_channels=hidden_size)          # models.py:130-132
        elif backbone_type == "neuraltx":
            self.backbone = bb.NeuralTX(hidden_channels=hidden_size)      # models.py:133-135
        elif backbone_type == "rvtdcnn":
            self.backbone = bb.RVTDCNN(fc_hid_size=hidden_size)       # models.py:80-81
        else:
            raise ValueError(f"The backbone type '{backbone_type}' is not provided natively (native: {NATIVE_BACKBONES}); "
                             "use the reference's PyTorch backbone for it.")
        try:
            self.backbone.reset_parameters()
            print("Backbone Initialized...")
        except AttributeError:
            pass

    def forward(self, x, h_0=None):
        # reference models.py:154-155 materialises h_0 = zeros(L,B,H); the native kernels start from the zero state.
        return self.backbone(x, h_0)

    def forward_mse(self, x, target, loss_count=None):
        return self.backbone.forward_mse(x, target, loss_count)


class CascadedModel(nn.Module):
    """reference models.py:163-176"""

    def __init__(self, dpd_model, pa_model):
        super().__init__()
        self.dpd_model, self.pa_model = dpd_model, pa_model

    def freeze_pa_model(self):
        for param in self.pa_model.parameters():
            param.requires_grad = False

    def forward(self, x):
        return self.pa_model(self.dpd_model(x))

    def forward_mse(self, x, target, loss_count=None):
        return self.pa_model.forward_mse(self.dpd_model(x), target, loss_count)

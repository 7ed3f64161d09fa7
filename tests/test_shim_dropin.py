"""Zero-edit drop-in: the UNMODIFIED reference steps (steps/train_pa.py, steps/train_dpd.py) run on the native backbones
through the shim/ directory.  Needs the reference checkout (authoring container); skipped where /root/reference is absent (GPU box).
Checked up to the point where training would start (no GPU here): model construction through the reference's own call sites, model
ids identical to the reference's (they embed the parameter count, project.py:57-92), the PA checkpoint written by the REFERENCE's
CoreModel loading into the native one, frozen-PA cascade assembled, optimizer built."""
import importlib
import importlib.util
import os
import sys
import pytest
import torch

from tests.util import ROOT

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "main.py")), reason="reference checkout not present")

_SHADOWED = ("models", "quant", "backbones", "steps", "project", "modules", "utils", "arguments")


def _clean_modules():
    for k in list(sys.modules):
        if k in _SHADOWED or k.split(".")[0] in _SHADOWED:
            del sys.modules[k]


@pytest.fixture()
def shimmed(tmp_path, monkeypatch):
    """sys.path = [shim, repo, reference root, ...] exactly as `python -m opendpd_b200.run_reference` builds it; cwd = a scratch dir."""
    _clean_modules()
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("PYTHONDONTWRITEBYTECODE", "1")
    monkeypatch.setattr(sys, "dont_write_bytecode", True)
    for p in (REF, ROOT, os.path.join(ROOT, "shim")):
        monkeypatch.syspath_prepend(p)
    yield tmp_path
    _clean_modules()


def _reference_models():
    """The reference's own models.py under a private name (the shim shadows `models`)."""
    spec = importlib.util.spec_from_file_location("_ref_models", os.path.join(REF, "models.py"))
    # reference backbones import `quant.modules.ops`: load them with the reference's quant package visible
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "quant" or k.startswith("quant.") or k == "backbones" or k.startswith("backbones.")}
    path = list(sys.path)
    try:
        sys.path[:] = [REF] + [p for p in path if not p.endswith("shim")]
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        net = mod.CoreModel(input_size=2, hidden_size=13, num_layers=1, backbone_type="dgru")
    finally:
        sys.path[:] = path
        for k in list(sys.modules):
            if k == "quant" or k.startswith("quant.") or k == "backbones" or k.startswith("backbones."):
                del sys.modules[k]
        sys.modules.update(saved)
    return net


def _run_step(step, argv, monkeypatch):
    import project
    captured = {}

    def fake_train(self, net, criterion, optimizer, lr_scheduler, train_loader, val_loader, test_loader, best_model_metric):
        captured.update(net=net, optimizer=optimizer, criterion=criterion, metric=best_model_metric, proj=self)
    monkeypatch.setattr(project.Project, "train", fake_train)
    monkeypatch.setattr(sys, "argv", ["main.py"] + argv)
    steps = importlib.import_module("steps." + step)
    steps.main(project.Project())
    return captured


def test_train_pa_runs_unmodified_on_native_backbones(shimmed, monkeypatch):
    import opendpd_b200.models as native
    got = _run_step("train_pa", ["--dataset_name", "APA_200MHz", "--step", "train_pa", "--accelerator", "cpu", "--PA_backbone", "dgru",
                                 "--PA_hidden_size", "13", "--frame_length", "200", "--batch_size", "64", "--n_epochs", "1"], monkeypatch)
    net, proj = got["net"], got["proj"]
    assert type(net) is native.CoreModel and type(net.backbone).__module__.startswith("opendpd_b200.backbones")
    n = sum(p.numel() for p in net.parameters())
    assert n == 1041
    assert proj.gen_pa_model_id(n).endswith("_P_1041") and "DGRU" in proj.gen_pa_model_id(n)
    assert isinstance(got["criterion"], torch.nn.MSELoss) and got["metric"] == "NMSE"
    assert sum(p.numel() for g in got["optimizer"].param_groups for p in g["params"]) == 1041


def test_train_dpd_runs_unmodified_and_loads_a_reference_checkpoint(shimmed, monkeypatch):
    import opendpd_b200.models as native
    # a PA checkpoint written by the REFERENCE's CoreModel (its nn.GRU-based DGRU), at the path train_dpd.py:39-40 looks for
    ref_pa = _reference_models()
    import project
    monkeypatch.setattr(sys, "argv", ["main.py", "--dataset_name", "APA_200MHz", "--step", "train_pa", "--accelerator", "cpu",
                                      "--PA_backbone", "dgru", "--PA_hidden_size", "13"])
    pa_id = project.Project().gen_pa_model_id(1041)
    os.makedirs(os.path.join("save", "APA_200MHz", "train_pa"), exist_ok=True)
    torch.save(ref_pa.state_dict(), os.path.join("save", "APA_200MHz", "train_pa", pa_id + ".pt"))
    got = _run_step("train_dpd", ["--dataset_name", "APA_200MHz", "--step", "train_dpd", "--accelerator", "cpu", "--PA_backbone", "dgru",
                                  "--PA_hidden_size", "13", "--DPD_backbone", "deltagru_tcnskip", "--DPD_hidden_size", "15", "--thx", "0.01",
                                  "--thh", "0.05", "--frame_length", "200", "--batch_size", "64", "--n_epochs", "1"], monkeypatch)
    net = got["net"]
    assert type(net) is native.CascadedModel
    assert sum(p.numel() for p in net.dpd_model.parameters()) == 999 and net.dpd_model.backbone.thx == 0.01 and net.dpd_model.backbone.thh == 0.05
    assert all(not p.requires_grad for p in net.pa_model.parameters())                      # freeze_pa_model (train_dpd.py:63)
    for (ka, va), (kb, vb) in zip(ref_pa.state_dict().items(), net.pa_model.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)                                              # the reference checkpoint went in bit for bit
    assert got["metric"] == "ACLR_AVG"

"""Golden vectors for the NEXT backbones (oracle/next_cells.py), from the unmodified reference in this container:
    python oracle/make_next_golden.py   ->  tests/golden/next_vdlstm_*.npz   (fp64 forward + nn.MSELoss + autograd backward)"""
import os, sys
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import numpy as np
import torch
import models                      # reference models.py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import load_apa, frames     # same framing of the reference's APA_200MHz data as the hot-path goldens


def main():
    X, Y = load_apa()
    for H, B, T, seed in ((8, 3, 40, 0), (12, 2, 129, 1)):
        torch.manual_seed(seed)
        net = models.CoreModel(input_size=2, hidden_size=H, num_layers=1, backbone_type="vdlstm").double()
        x, y = frames(X, Y, B, T, seed)
        xt = torch.from_numpy(x).double().requires_grad_(True)
        out = net(xt, torch.zeros(1, B, H, dtype=torch.float64)) if False else net.backbone(xt, None)
        loss = torch.nn.MSELoss()(out, torch.from_numpy(y).double())
        loss.backward()
        params = np.concatenate([p.detach().numpy().ravel() for _, p in net.backbone.named_parameters()])
        grads = np.concatenate([p.grad.numpy().ravel() for _, p in net.backbone.named_parameters()])
        names = [n for n, _ in net.backbone.named_parameters()]
        f = os.path.join(ROOT, "tests", "golden", f"next_vdlstm_h{H}_b{B}_t{T}.npz")
        np.savez_compressed(f, x=x, y=y, params=params, out=out.detach().numpy(), loss=float(loss.item()), gx=xt.grad.numpy(), gparams=grads,
                            H=H, names=np.array(names))
        print(f, params.size, float(loss.item()))


if __name__ == "__main__":
    main()

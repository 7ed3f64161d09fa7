"""Comparator of BASELINE.md §2: the reference's op sequence (oracle/torch_port.py: torch._VF.gru -> cuDNN, ATen linear/relu/cat, nn.MSELoss,
clip_grad_norm_, torch AdamW) executed by stock PyTorch ON THE GPU for the headline workload — the "existing GPU path" of the reference.
    python scripts/torch_cuda_leg.py            -> one JSON line"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import torch_port, oracle

kind, H, B, T = "dgru", 13, 64, 2048
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
flat = torch.nn.Parameter((0.3 * torch.randn(oracle.n_params(kind, H), generator=g)).to(dev))
opt = torch.optim.AdamW([flat], lr=5e-4)
crit = torch.nn.MSELoss()
x = (0.2 * torch.randn(B, T, 2, generator=g)).to(dev)
y = (0.9 * x).contiguous()


def step():
    opt.zero_grad()
    loss = crit(torch_port.forward(kind, x, flat, H), y)
    loss.backward()
    torch.nn.utils.clip_grad_norm_([flat], 200.0)
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"what": "stock PyTorch on the B200 (cuDNN GRU + ATen ops), DGRU H13 B64xT2048 train step", "ms_per_step": ms,
                  "iq_samples_per_s": B * T / (ms * 1e-3), "loss": float(loss.item()), "torch": torch.__version__}))

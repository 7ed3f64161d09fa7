"""CPU restatement (numpy, float64) of backbones that are NEXT on the list — TEST INFRASTRUCTURE prepared ahead of
their CUDA kernels; nothing in opendpd_b200 imports this file and no CUDA path exists for these cells yet.

VDLSTM — /root/reference/backbones/vdlstm.py:58-82 (forward), parameters :28-41:
  window W = 4 over the frame with WRAP-AROUND padding (the last W-1 samples of the frame are prepended, :65-73), per window position k:
  amp_k = sqrt(I^2+Q^2), cos_k = I/amp, sin_k = Q/amp;  h_t = LSTM(amp window)  (gate order i,f,g,o; (h,c) start at 0);
  lambda1 = W1 h + b1, lambda2 = W2 h + b2 (H -> W each);  out_t = Wo [lambda1 * cos ; lambda2 * sin] + bo   (2W -> 2).
Flat parameter order = named_parameters(): rnn.weight_ih_l0 (4H,W) weight_hh_l0 (4H,H) bias_ih_l0 (4H) bias_hh_l0 (4H)
fc_lambda_1.weight (W,H) .bias (W) fc_lambda_2.weight (W,H) .bias (W) fc_out.weight (2,2W) .bias (2).
The backward is hand-derived (reverse-time LSTM adjoint + scatter of the window gradients onto the samples) and pinned, with the
forward, to golden vectors made by the unmodified reference's autograd (oracle/make_next_golden.py -> tests/golden/next_vdlstm_*.npz)."""
import numpy as np

W = 4


def vdlstm_split(flat, H):
    shapes = [("w_ih", (4 * H, W)), ("w_hh", (4 * H, H)), ("b_ih", (4 * H,)), ("b_hh", (4 * H,)), ("w1", (W, H)), ("b1", (W,)),
              ("w2", (W, H)), ("b2", (W,)), ("wo", (2, 2 * W)), ("bo", (2,))]
    out, off = {}, 0
    for name, shp in shapes:
        n = int(np.prod(shp))
        out[name] = flat[off:off + n].reshape(shp)
        off += n
    assert off == flat.size, (off, flat.size)
    return out


def _sig(v):
    return 1.0 / (1.0 + np.exp(-v))


def vdlstm(x, flat, H, target=None):
    """x (B,T,2) -> dict(out, loss, gx, gparams); loss = mean squared error over all B*T*2 outputs (nn.MSELoss)."""
    x = np.asarray(x, np.float64)
    flat = np.asarray(flat, np.float64)
    p = vdlstm_split(flat, H)
    B, T, _ = x.shape
    idx = (np.arange(T)[:, None] - (W - 1) + np.arange(W)[None, :]) % T            # window t = samples (t-3+k) mod T
    I, Q = x[..., 0][:, idx], x[..., 1][:, idx]                                    # (B,T,W)
    amp = np.sqrt(I ** 2 + Q ** 2)
    cos, sin = I / amp, Q / amp
    h = np.zeros((B, H)); c = np.zeros((B, H))
    gates, hs, cs, cprev = [], [], [], []
    for t in range(T):
        z = amp[:, t] @ p["w_ih"].T + p["b_ih"] + h @ p["w_hh"].T + p["b_hh"]
        i, f, g, o = _sig(z[:, :H]), _sig(z[:, H:2 * H]), np.tanh(z[:, 2 * H:3 * H]), _sig(z[:, 3 * H:])
        cprev.append(c)
        c = f * c + i * g
        h = o * np.tanh(c)
        gates.append((i, f, g, o)); hs.append(h); cs.append(c)
    hseq = np.stack(hs, 1)                                                          # (B,T,H)
    l1 = hseq @ p["w1"].T + p["b1"]
    l2 = hseq @ p["w2"].T + p["b2"]
    feat = np.concatenate((l1 * cos, l2 * sin), -1)                                 # (B,T,2W)
    out = feat @ p["wo"].T + p["bo"]
    res = dict(out=out, loss=None, gx=None, gparams=None)
    if target is None:
        return res
    target = np.asarray(target, np.float64)
    n = out.size
    res["loss"] = float(((out - target) ** 2).sum() / n)
    go = 2.0 * (out - target) / n                                                   # (B,T,2)
    g = {k: np.zeros_like(v) for k, v in p.items()}
    g["wo"] = np.einsum("bto,btf->of", go, feat); g["bo"] = go.sum((0, 1))
    gfeat = go @ p["wo"]                                                            # (B,T,2W)
    gl1, gl2 = gfeat[..., :W] * cos, gfeat[..., W:] * sin
    gcos, gsin = gfeat[..., :W] * l1, gfeat[..., W:] * l2
    g["w1"] = np.einsum("btw,bth->wh", gl1, hseq); g["b1"] = gl1.sum((0, 1))
    g["w2"] = np.einsum("btw,bth->wh", gl2, hseq); g["b2"] = gl2.sum((0, 1))
    gh_head = gl1 @ p["w1"] + gl2 @ p["w2"]                                         # (B,T,H)
    gamp = np.zeros((B, T, W))
    gh = np.zeros((B, H)); gc = np.zeros((B, H))
    for t in range(T - 1, -1, -1):
        i, f, gg, o = gates[t]
        gh = gh + gh_head[:, t]
        tc = np.tanh(cs[t])
        gct = gc + gh * o * (1 - tc ** 2)
        dz = np.concatenate((gct * gg * i * (1 - i), gct * cprev[t] * f * (1 - f), gct * i * (1 - gg ** 2), gh * tc * o * (1 - o)), -1)
        hp = hs[t - 1] if t > 0 else np.zeros((B, H))
        g["w_ih"] += dz.T @ amp[:, t]; g["w_hh"] += dz.T @ hp
        g["b_ih"] += dz.sum(0); g["b_hh"] += dz.sum(0)
        gamp[:, t] = dz @ p["w_ih"]
        gh = dz @ p["w_hh"]
        gc = gct * f
    # window quantities -> samples: amp = |x|, cos = I/amp, sin = Q/amp at window position (t,k) <-> sample idx[t,k]
    gI_w = gamp * I / amp + gcos * (1 / amp - I * I / amp ** 3) + gsin * (-Q * I / amp ** 3)
    gQ_w = gamp * Q / amp + gcos * (-I * Q / amp ** 3) + gsin * (1 / amp - Q * Q / amp ** 3)
    gx = np.zeros_like(x)
    for k in range(W):
        np.add.at(gx[..., 0], (slice(None), idx[:, k]), gI_w[..., k])
        np.add.at(gx[..., 1], (slice(None), idx[:, k]), gQ_w[..., k])
    res["gx"] = gx
    res["gparams"] = np.concatenate([g[k].ravel() for k in ("w_ih", "w_hh", "b_ih", "b_hh", "w1", "b1", "w2", "b2", "wo", "bo")])
    return res

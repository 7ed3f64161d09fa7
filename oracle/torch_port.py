"""PyTorch-ATen restatement of the reference backbones (TEST/BENCH INFRASTRUCTURE — never imported by opendpd_b200).

The reference's CPU path *is* a sequence of PyTorch ATen ops: torch.nn.GRU/LSTM (ATen RNN.cpp) for
gru/lstm/dgru/qgru and one ATen call per Python op for the hand-written cells.  This file restates those op sequences
functionally over the flat parameter vector of include/odpd.h, so that (a) bench.py can time what the reference's own
CPU path costs on the GPU box's host cores, where /root/reference does not exist, and (b) tests have a second,
autograd-derived check of the hand-written backward in odpd_oracle.c.  Pinned against tests/golden/*.npz by
tests/test_torch_port.py.  Citations: gru.py:45-48, lstm.py:45-48, dgru.py:59-74, qgru.py:59-71, qgru_amp1.py:59-76,
deltagru.py:60-77/149-266, deltagru_tcnskip.py:89-103/232-304, pgjanet.py:24-77, dvrjanet.py:32-102, gmp.py:18-51."""
import math
import torch
import torch.nn.functional as Fn


def split_params(kind, flat, H, K=3):
    """flat (P,) -> dict of named views, named_parameters() order of the reference module."""
    shapes = {
        "gru": [("w_ih", (3 * H, 2)), ("w_hh", (3 * H, H)), ("b_ih", (3 * H,)), ("b_hh", (3 * H,)), ("wo", (2, H)), ("bo", (2,))],
        "qgru": [("w_ih", (3 * H, 4)), ("w_hh", (3 * H, H)), ("b_ih", (3 * H,)), ("b_hh", (3 * H,)), ("wo", (2, H)), ("bo", (2,))],
        "lstm": [("w_ih", (4 * H, 2)), ("w_hh", (4 * H, H)), ("b_ih", (4 * H,)), ("b_hh", (4 * H,)), ("wo", (2, H)), ("bo", (2,))],
        "dgru": [("w_ih", (3 * H, 6)), ("w_hh", (3 * H, H)), ("b_ih", (3 * H,)), ("b_hh", (3 * H,)), ("wo", (2, H + 6)), ("bo", (2,)),
                 ("wh", (H, H)), ("bh", (H,))],
        "deltagru": [("w_ih", (3 * H, 6)), ("w_hh", (3 * H, H)), ("b_ih", (3 * H,)), ("b_hh", (3 * H,)), ("wo", (2, H)), ("bo", (2,))],
        "deltagru_tcnskip": [("w_ih", (3 * H, 6)), ("w_hh", (3 * H, H)), ("wo", (2, H)), ("c0", (3, 2, 3)), ("c2", (2, 3, 1))],
        "pgjanet": [("wa", (H, H + 1)), ("ba", (H,)), ("wp1", (H, H + 1)), ("bp1", (H,)), ("wp2", (H, H + 1)), ("bp2", (H,)),
                    ("wf", (H, 2 * H)), ("bf", (H,)), ("wg", (H, 2 * H)), ("bg", (H,)), ("wo", (2, H)), ("bo", (2,))],
        "dvrjanet": [("cs", (K,)), ("wph", (H, H)), ("wpt", (H, 1)), ("wah", (H, H)), ("wax", (H, 1)), ("wf", (H, H)), ("bf", (H,)),
                     ("wc", (H, 2 * H)), ("bc", (H,)), ("ws", (H, 2 * H)), ("bs", (H,)), ("wo1", (1, H)), ("bo1", (1,)),
                     ("wo2", (1, H)), ("bo2", (1,))],
        "gmp": [("w", (1, 495))],
        "bojanet": [("fi", (6, 16)), ("fq", (6, 16)), ("wfi", (H, 12)), ("bfi", (H,)), ("wfh", (H, H)), ("wgi", (H, 12)), ("bgi", (H,)),
                    ("wgh", (H, H)), ("woi", (1, H)), ("boi", (1,)), ("woq", (1, H)), ("boq", (1,))],
        "tcnn": [("w0", (H, 6, 1)), ("b0", (H,)), ("d1", (H, 1, 5)), ("d2", (H, 1, 5)), ("d4", (H, 1, 5)), ("d8", (H, 1, 5)), ("w10", (2, H, 1))],
        "neuraltx": [("ci", (1, 1, 5)), ("cq", (1, 1, 5)), ("w0", (H, 4, 1)), ("b0", (H,)), ("d1", (H, 1, 5)), ("d2", (H, 1, 5)), ("d4", (H, 1, 5)),
                     ("d8", (H, 1, 5)), ("w10", (2, H, 1)), ("iq", (2, 2))],
        "apnrru": [("fi", (3, 16)), ("fq", (3, 16)), ("C", (1,)), ("Z", (1, 2 * H + 3)), ("wu", (16, 2 * H + 11)), ("bu", (16,)),
                   ("wh", (2 * H + 3, 16)), ("bh", (2 * H + 3,)), ("oi", (1, H)), ("oq", (1, H))],
        "mcldnn": [("c1w", (H, 1, 3, 3)), ("c1b", (H,)), ("cdw", (5 * H, 1, 3)), ("cdb", (5 * H,)), ("c2w", (1, 10, 3, 3)), ("c2b", (1,)),
                   ("wih", (32, 5 * H)), ("whh", (32, 8)), ("bih", (32,)), ("bhh", (32,)), ("f1w", (16, 8)), ("f1b", (16,)), ("f2w", (2, 16)), ("f2b", (2,))],
        "deltajanet": [("w_ih", (2 * H, 6)), ("w_hh", (2 * H, H)), ("b_ih", (2 * H,)), ("b_hh", (2 * H,)), ("wo", (2, H)), ("bo", (2,))],
        "rvtdcnn": [("wc", (3, 1, 3, 3)), ("bc", (3,)), ("wh", (H, 36)), ("bh", (H,)), ("wo", (2, H)), ("bo", (2,))],
    }
    shapes["qgru_amp1"] = shapes["qgru"]
    shapes["tres"] = shapes["deltagru_tcnskip"]
    out, off = {}, 0
    for name, shp in shapes[kind]:
        n = math.prod(shp)
        out[name] = flat[off:off + n].view(shp)
        off += n
    assert off == flat.numel(), (kind, off, flat.numel())
    return out


def _feat(kind, x):
    i, q = x[..., 0:1], x[..., 1:2]
    if kind in ("gru", "lstm"):
        return x
    a2 = torch.pow(i, 2) + torch.pow(q, 2)
    if kind == "qgru":
        return torch.cat((i, q, a2, torch.pow(a2, 2)), -1)
    a = torch.sqrt(a2)
    a3 = torch.pow(a, 3)
    if kind == "qgru_amp1":
        return torch.cat((i, q, a, a3), -1)
    if kind in ("deltagru_tcnskip", "tres"):
        nxt = torch.roll(x, shifts=-1, dims=1)
        return torch.cat((i, q, a, a3, nxt[..., 0:1], nxt[..., 1:2]), -1)
    return torch.cat((i, q, a, a3, q / a, i / a), -1)


def _delta_layer(f, p, H, thx, thh, bias):
    B, T, _ = f.shape
    thx_t, thh_t = torch.tensor(thx, dtype=f.dtype), torch.tensor(thh, dtype=f.dtype)
    xp = f.new_zeros(B, 6); h = f.new_zeros(B, H); hp = f.new_zeros(B, H)
    M = f.new_zeros(B, 3 * H); Mnh = f.new_zeros(B, H)
    if bias:
        M = M + torch.cat((p["b_ih"][:2 * H] + p["b_hh"][:2 * H], p["b_ih"][2 * H:]))
        Mnh = Mnh + p["b_hh"][2 * H:]
    hs = []
    for t in range(T):
        xt = f[:, t]
        dx, dh = xt - xp, h - hp
        ax, ah = dx.abs(), dh.abs()
        dx = dx.masked_fill(ax < thx_t, 0); dh = dh.masked_fill(ah < thh_t, 0)
        xp = torch.where(ax >= thx_t, xt, xp); hp = torch.where(ah >= thh_t, h, hp)
        mx = torch.mm(dx, p["w_ih"].t()) + M
        mh = torch.mm(dh, p["w_hh"].t())
        Mr, Mz, Mn = mx[:, :H] + mh[:, :H], mx[:, H:2 * H] + mh[:, H:2 * H], mx[:, 2 * H:]
        Mnh = mh[:, 2 * H:] + Mnh
        M = torch.cat((Mr, Mz, Mn), 1)
        r, z = torch.sigmoid(Mr), torch.sigmoid(Mz)
        n = torch.tanh(Mn + r * Mnh)
        h = (1 - z) * n + z * h
        hs.append(h)
    return torch.stack(hs, 1)


def _rnn_stack_split(kind, flat, H, L):
    """Flat vector of an L-layer nn.GRU / nn.LSTM backbone -> (per-layer [w_ih, w_hh, b_ih, b_hh] list in nn.RNNBase's flat order,
    dict of the head tensors).  named_parameters() order: rnn.weight_ih_l0 .. bias_hh_l{L-1}, fc_out.*, (dgru) fc_hid.*."""
    G = 4 if kind == "lstm" else 3
    F = {"gru": 2, "lstm": 2, "dgru": 6, "qgru": 4, "qgru_amp1": 4}[kind]
    ws, off = [], 0
    for l in range(L):
        fin = F if l == 0 else H
        for shp in ((G * H, fin), (G * H, H), (G * H,), (G * H,)):
            n = math.prod(shp)
            ws.append(flat[off:off + n].view(shp)); off += n
    O = H + 6 if kind == "dgru" else H
    head = {}
    for name, shp in (("wo", (2, O)), ("bo", (2,))) + ((("wh", (H, H)), ("bh", (H,))) if kind == "dgru" else ()):
        n = math.prod(shp)
        head[name] = flat[off:off + n].view(shp); off += n
    assert off == flat.numel(), (kind, H, L, off, flat.numel())
    return ws, head


def forward_layers(kind, x, flat, H, L):
    """gru / lstm / dgru / qgru / qgru_amp1 with num_layers = L (the reference passes --*_num_layers straight to nn.GRU / nn.LSTM:
    gru.py:17-24, lstm.py:17-24, dgru.py:22-28, qgru.py:22-28), any hidden size."""
    ws, p = _rnn_stack_split(kind, flat, H, L)
    B = x.shape[0]
    f = _feat(kind, x)
    h0 = x.new_zeros(L, B, H)
    if kind == "lstm":
        hseq, _, _ = torch._VF.lstm(f, (h0, h0), ws, True, L, 0.0, x.is_cuda, False, True)
    else:
        hseq, _ = torch._VF.gru(f, h0, ws, True, L, 0.0, x.is_cuda, False, True)
    if kind == "dgru":
        g = torch.relu(Fn.linear(hseq, p["wh"], p["bh"]))
        return Fn.linear(torch.cat((g, f), -1), p["wo"], p["bo"])
    return Fn.linear(hseq, p["wo"], p["bo"])


def forward(kind, x, flat, H, K=3, thx=0.0, thh=0.0, L=1):
    """x (B,T,2) -> out (B,T,2), differentiable w.r.t. x and flat."""
    if L > 1:
        return forward_layers(kind, x, flat, H, L)
    p = split_params(kind, flat, H, K)
    B, T, _ = x.shape
    if kind in ("gru", "qgru", "qgru_amp1", "dgru"):
        f = _feat(kind, x)
        h0 = x.new_zeros(1, B, H)
        hseq, _ = torch._VF.gru(f, h0, [p["w_ih"], p["w_hh"], p["b_ih"], p["b_hh"]], True, 1, 0.0, x.is_cuda, False, True)   # train flag: cuDNN's backward needs it (dropout is 0: same arithmetic)
        if kind == "dgru":
            g = torch.relu(Fn.linear(hseq, p["wh"], p["bh"]))
            return Fn.linear(torch.cat((g, f), -1), p["wo"], p["bo"])
        return Fn.linear(hseq, p["wo"], p["bo"])
    if kind == "lstm":
        h0 = x.new_zeros(1, B, H)
        hseq, _, _ = torch._VF.lstm(x, (h0, h0), [p["w_ih"], p["w_hh"], p["b_ih"], p["b_hh"]], True, 1, 0.0, x.is_cuda, False, True)
        return Fn.linear(hseq, p["wo"], p["bo"])
    if kind == "deltagru":
        hseq = _delta_layer(_feat("dgru", x), p, H, thx, thh, True)
        return Fn.linear(hseq, p["wo"], p["bo"])
    if kind in ("deltagru_tcnskip", "tres"):
        xt = x.transpose(1, 2)
        s = Fn.hardswish(Fn.conv1d(xt, p["c0"], None, 1, 16, 16))
        s = Fn.hardswish(Fn.conv1d(s, p["c2"])).transpose(1, 2)
        hseq = _delta_layer(_feat("tres", x), p, H, thx, thh, False)
        return Fn.linear(hseq, p["wo"]) + s
    if kind == "pgjanet":
        h = x.new_zeros(B, H); ys = []
        for t in range(T):
            i, q = x[:, t, 0:1], x[:, t, 1:2]
            a = torch.sqrt(i ** 2 + q ** 2); th = torch.atan2(q, i)
            an = torch.tanh(Fn.linear(torch.cat((h, a), -1), p["wa"], p["ba"]))
            p1 = torch.tanh(Fn.linear(torch.cat((h, torch.cos(th)), -1), p["wp1"], p["bp1"]))
            p2 = torch.tanh(Fn.linear(torch.cat((h, torch.sin(th)), -1), p["wp2"], p["bp2"]))
            u = an * p1 * p2 * (1 - an) * (1 - p1) * (1 - p2)
            hu = torch.cat((h, u), -1)
            f = torch.sigmoid(Fn.linear(hu, p["wf"], p["bf"])); g = torch.tanh(Fn.linear(hu, p["wg"], p["bg"]))
            h = f * h + (1 - f) * g
            ys.append(Fn.linear(h, p["wo"], p["bo"]))
        return torch.stack(ys, 1)
    if kind == "dvrjanet":
        hI = x.new_zeros(B, H); hQ = x.new_zeros(B, H); ys = []
        for t in range(T):
            i, q = x[:, t, 0:1], x[:, t, 1:2]
            a = torch.sqrt(i ** 2 + q ** 2); th = torch.atan2(q, i)
            s = hI + hQ
            tht = Fn.linear(th, p["wpt"]) + Fn.linear(s, p["wph"])
            pa = Fn.linear(a, p["wax"]) + Fn.linear(s, p["wah"])
            at = 0
            for k in range(1, K + 1):
                at = at + torch.abs(pa - k / K) * p["cs"][k - 1]
            f = torch.sigmoid(Fn.linear(s, p["wf"], p["bf"]))
            gc = torch.tanh(Fn.linear(torch.cat((hI, at * torch.cos(tht)), -1), p["wc"], p["bc"]))
            gs = torch.tanh(Fn.linear(torch.cat((hQ, at * torch.sin(tht)), -1), p["ws"], p["bs"]))
            hI = f * hI + (1 - f) * gc; hQ = f * hQ + (1 - f) * gs
            ys.append(torch.cat((Fn.linear(hI, p["wo1"], p["bo1"]), Fn.linear(hQ, p["wo2"], p["bo2"])), -1))
        return torch.stack(ys, 1)
    if kind == "deltajanet":  # deltajanet.py:49-60, 203-262 — the layer is built with thx = thh = 0 (:22-26): every delta passes
        f = _feat("dgru", x)
        xp = f.new_zeros(B, 6); h = f.new_zeros(B, H); hp = f.new_zeros(B, H)
        M = f.new_zeros(B, 2 * H) + (p["b_ih"] + p["b_hh"])
        hs = []
        for t in range(T):
            dx, dh = f[:, t] - xp, h - hp
            xp, hp = f[:, t], h
            mx = torch.mm(dx, p["w_ih"].t()) + M
            mh = torch.mm(dh, p["w_hh"].t())
            M = torch.cat((mx[:, :H] + mh[:, :H], mx[:, H:] + mh[:, H:]), 1)
            gf, gg = torch.sigmoid(M[:, :H]), torch.sigmoid(M[:, H:])
            h = (1 - gf) * gg + gf * h
            hs.append(h)
        return Fn.linear(torch.stack(hs, 1), p["wo"], p["bo"])
    if kind == "apnrru":      # apnrru.py:52-135
        xx = torch.cat((torch.zeros_like(x[:, -15:, :]), x), 1)
        win = xx.unfold(1, 16, 1).transpose(2, 3)                   # (B,T,16,2): tap m of window t = sample t+m-15
        wi, wq = win[..., 0], win[..., 1]
        i_fir = Fn.linear(wi, p["fi"]) - Fn.linear(wq, p["fq"])     # (B,T,3)
        q_fir = Fn.linear(wi, p["fq"]) + Fn.linear(wq, p["fi"])
        li, lq = x[..., 0], x[..., 1]
        mag = torch.sqrt(li ** 2 + lq ** 2)
        rr, ri = li / mag, -lq / mag                                 # r = conj(x)/|x|
        a = torch.cat((i_fir, li.unsqueeze(-1)), -1)                 # (B,T,4) real parts
        bq = torch.cat((q_fir, lq.unsqueeze(-1)), -1)
        nre = rr.unsqueeze(-1) * a - ri.unsqueeze(-1) * bq
        nim = ri.unsqueeze(-1) * a + rr.unsqueeze(-1) * bq
        xin = torch.stack((nre, nim), -1).reshape(B, T, 8)
        hI = x.new_zeros(B, H); hQ = x.new_zeros(B, H); hA = x.new_zeros(B, 3); outs = []
        for t in range(T):
            r0, r1 = rr[:, t:t + 1], ri[:, t:t + 1]
            hI, hQ = hI * r0 - hQ * r1, hI * r1 + hQ * r0
            hnew = torch.cat((hI, hQ, hA), -1)
            v = torch.tanh(Fn.linear(torch.cat((xin[:, t], hnew), -1), p["wu"], p["bu"]))
            v = torch.tanh(Fn.linear(v, p["wh"], p["bh"]))
            v = torch.sigmoid(p["C"] * hnew) + p["Z"] * v
            aI, aQ, hA = v[:, :H], v[:, H:2 * H], v[:, 2 * H:]
            hI, hQ = r0 * aI + r1 * aQ, r0 * aQ - r1 * aI           # times conj(r)
            oi, oq = Fn.linear(hI, p["oi"]), Fn.linear(hQ, p["oq"])
            outs.append(torch.cat((oi - oq, oq + oi), -1))
        return torch.stack(outs, 1)
    if kind == "bojanet":     # bojanet.py:54-106
        xx = torch.cat((torch.zeros_like(x[:, -15:, :]), x), 1)
        win = xx.unfold(1, 16, 1).transpose(2, 3)                   # (B,T,16,2): tap m of window t = sample t+m-15
        wi, wq = win[..., 0], win[..., 1]
        i_fir = Fn.linear(wi, p["fi"]) - Fn.linear(wq, p["fq"])
        q_fir = Fn.linear(wi, p["fq"]) + Fn.linear(wq, p["fi"])
        mag = torch.sqrt(torch.pow(i_fir, 2) + torch.pow(q_fir, 2)) + 1e-8
        sin, cos = q_fir / mag, i_fir / mag
        Lf = torch.cat((mag, mag ** 2), -1)
        h = x.new_zeros(B, H); hs = []
        for t in range(T):
            f = torch.sigmoid(Fn.linear(Lf[:, t], p["wfi"], p["bfi"]) + Fn.linear(h, p["wfh"]))
            g = torch.tanh(Fn.linear(Lf[:, t], p["wgi"], p["bgi"]) + Fn.linear(h, p["wgh"]))
            h = f * h + (1 - f) * g
            hs.append(h)
        hseq = torch.stack(hs, 1)
        idx = torch.arange(H) % 6                                  # pr_block :41-52 for hidden_size <= 18
        a = Fn.linear(hseq * cos[..., idx], p["woi"], p["boi"])
        qq = Fn.linear(hseq * sin[..., idx], p["woq"], p["boq"])
        return torch.cat((a - qq, qq + a), -1)
    if kind in ("tcnn", "neuraltx"):     # tcnn.py:83-97, neuraltx.py:107-124
        def stack(u):
            u = Fn.hardswish(Fn.conv1d(u, p["w0"], p["b0"]))
            for name, d in (("d1", 1), ("d2", 2), ("d4", 4), ("d8", 8)):
                u = Fn.hardswish(Fn.conv1d(u, p[name], None, 1, 2 * d, d, H))
            return Fn.conv1d(u, p["w10"])
        if kind == "tcnn":
            return stack(_feat("dgru", x).transpose(1, 2)).transpose(1, 2) + x
        it, qt = x[..., 0:1].transpose(1, 2), x[..., 1:2].transpose(1, 2)
        i_f = (Fn.conv1d(it, p["ci"], None, 1, 2) - Fn.conv1d(qt, p["cq"], None, 1, 2)).transpose(1, 2)
        q_f = (Fn.conv1d(it, p["cq"], None, 1, 2) + Fn.conv1d(qt, p["ci"], None, 1, 2)).transpose(1, 2)
        amp = torch.sqrt(torch.pow(i_f, 2) + torch.pow(q_f, 2))
        iq = torch.cat((i_f, q_f), -1)
        u = torch.cat((i_f, q_f, amp, torch.pow(amp, 3)), -1).transpose(1, 2)
        return stack(u).transpose(1, 2) + Fn.linear(iq, p["iq"]) + iq
    if kind == "mcldnn":      # mcldnn.py:83-113  (H = conv channels; the LSTM is always 8 wide)
        i, q = x[..., 0:1], x[..., 1:2]
        amp2 = torch.pow(i, 2) + torch.pow(q, 2)
        amp = torch.sqrt(amp2)
        f = torch.cat((i, q, amp, amp2, torch.pow(amp, 3)), -1)
        xx = torch.cat((f[:, -4:, :], f), 1)
        win = xx.unfold(1, 5, 1).contiguous().view(-1, 1, 5, 5)      # (N, 1, feature, memory)
        o2 = Fn.conv2d(win, p["c1w"], p["c1b"], 1, 1)                # (N, C, 5, 5)
        o1 = Fn.conv1d(win.squeeze(1), p["cdw"], p["cdb"], 1, 1, 1, 5).view(-1, H, 5, 5)
        o = torch.cat((o2, o1), 2)                                   # (N, C, 10, 5)
        o = Fn.conv2d(o.transpose(1, 2), p["c2w"], p["c2b"], 1, 1).view(B, T, -1)
        h0 = x.new_zeros(1, B, 8)
        hseq, _, _ = torch._VF.lstm(o, (h0, h0), [p["wih"], p["whh"], p["bih"], p["bhh"]], True, 1, 0.0, x.is_cuda, False, True)
        return Fn.linear(Fn.linear(hseq, p["f1w"], p["f1b"]), p["f2w"], p["f2b"])
    if kind == "rvtdcnn":     # rvtdcnn.py:36-62
        i, q = x[..., 0:1], x[..., 1:2]
        amp2 = torch.pow(i, 2) + torch.pow(q, 2)
        amp = torch.sqrt(amp2)
        f = torch.cat((i, q, amp, amp2, torch.pow(amp, 3)), -1)
        xx = torch.cat((f[:, -3:, :], f), 1)
        win = xx.unfold(1, 4, 1).transpose(2, 3).contiguous().view(-1, 1, 4, 5)
        o = torch.tanh(Fn.conv2d(win, p["wc"], p["bc"], 1, (1, 0))).view(-1, 36)
        o = torch.tanh(Fn.linear(o, p["wh"], p["bh"]))
        return Fn.linear(o, p["wo"], p["bo"]).view(B, T, 2)
    if kind == "gmp":
        xc = torch.complex(x[..., 0], x[..., 1])
        xp = torch.cat((xc.new_zeros(B, 20), xc), 1)              # xp[n] = x[n-20]
        amp = xp.abs()
        cols = []
        for m in range(11):
            cols.append(xp[:, 10 + m:10 + m + T])                   # x[j+m-10]
        for pw in range(1, 5):
            ap = amp ** pw
            for k in range(11):
                for m in range(11):
                    cols.append(xp[:, 10 + m:10 + m + T] * ap[:, k + m:k + m + T])   # x[j+m-10] |x[j+k+m-20]|^pw
        basis = torch.stack(cols, -1)                               # (B,T,495)
        y = (basis * p["w"].view(-1)).sum(-1)
        return torch.stack((y.real, y.imag), -1)
    raise ValueError(kind)


def make_train_step(kind, H, seed=0, K=3, thx=0.0, thh=0.0, lr=5e-4, clip=200.0, pa=None):
    """The net_train body (train_funcs.py:33-48) on the PyTorch-CPU restatement: zero_grad, forward, MSELoss, backward,
    clip_grad_norm_, AdamW step, loss.item().  pa=(kind, H): the train_dpd cascade (models.py CascadedModel — the DPD feeds a
    frozen PA model, steps/train_dpd.py:60-66)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle
    g = torch.Generator().manual_seed(seed)
    P = oracle.n_params(kind, H, K)
    flat = torch.nn.Parameter(0.3 * torch.randn(P, generator=g))
    opt = torch.optim.AdamW([flat], lr=lr)
    crit = torch.nn.MSELoss()
    pa_flat = None
    if pa is not None:
        pa_flat = 0.3 * torch.randn(oracle.n_params(pa[0], pa[1], 3), generator=g)     # frozen: no grad, not in the optimiser

    def step(x, y):
        opt.zero_grad()
        out = forward(kind, x, flat, H, K, thx, thh)
        if pa_flat is not None:
            out = forward(pa[0], out, pa_flat, pa[1])
        loss = crit(out, y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([flat], clip)
        opt.step()
        return loss.item()
    return step

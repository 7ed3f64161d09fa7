"""Flat parameter storage: every nn.Parameter of a native backbone is a view into one contiguous fp32 buffer laid
out in named_parameters() order (== include/odpd.h 'Flat parameter layout').  One pointer feeds the kernels, one
buffer feeds the NCCL all-reduce and the fused clip+AdamW step."""
import torch


class FlatParams:
    """Mixin for nn.Module backbones. Call self._flat_sync() before handing self._flat to a kernel."""

    _flat = None
    _flat_layout = None

    def flat_layout(self):
        """[(offset, numel, shape)] in named_parameters() order."""
        off, lay = 0, []
        for _, p in self.named_parameters():
            lay.append((off, p.numel(), tuple(p.shape)))
            off += p.numel()
        return lay, off

    def _flat_sync(self):
        params = [p for _, p in self.named_parameters()]
        lay, total = self.flat_layout()
        flat = self._flat
        ok = (flat is not None and flat.device == params[0].device and flat.numel() >= total and
              all(p.data_ptr() == flat.data_ptr() + 4 * off and p.is_contiguous() for p, (off, _, _) in zip(params, lay)))
        if not ok:
            pad = (total + 3) // 4 * 4
            flat = torch.zeros(pad, dtype=torch.float32, device=params[0].device)
            with torch.no_grad():
                for p, (off, n, shape) in zip(params, lay):
                    flat[off:off + n].copy_(p.detach().reshape(-1).to(torch.float32))
                    p.data = flat[off:off + n].view(shape)
            self._flat, self._flat_layout = flat, lay
        return self._flat, self._flat_layout

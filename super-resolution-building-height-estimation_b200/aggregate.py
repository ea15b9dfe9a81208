"""Block aggregation of a height label to the coarse grid (reference: aggregate_utils.py).

`aggregate_torch` (:29-41) is what the loader calls per sample (BH_loader.py:386); on a CUDA
tensor it runs the `bhsr_aggregate` kernel.  The reference calls it on CPU tensors inside
DataLoader worker processes (no CUDA context there), so host tensors are handled with the same
closed form in plain torch — this is loader-side label preparation, not the GPU hot path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _block_aggregate(data: torch.Tensor, step: int, thr: float, strict: bool) -> torch.Tensor:
    if data.dim() < 2:
        raise ValueError("aggregate expects [..., H, W]")
    h, w = data.shape[-2:]
    oh, ow = h // step, w // step
    lead = data.shape[:-2]
    if data.is_cuda:
        x = data.float().contiguous()
        out = torch.empty(lead + (oh, ow), dtype=torch.float32, device=data.device)
        nimg = int(np.prod(lead)) if lead else 1
        with _lib.on_device(data):
            _lib.check(_lib.load().bhsr_aggregate(x.data_ptr(), nimg, h, w, step, float(thr), int(strict),
                                                  out.data_ptr(), _lib.stream_ptr(data.device)), "bhsr_aggregate")
        return out
    x = data.float()[..., : oh * step, : ow * step].reshape(lead + (oh, step, ow, step))
    s1 = x.sum(dim=(-3, -1))
    mask = (x > thr) if strict else (x >= thr)
    s2 = mask.float().sum(dim=(-3, -1))
    return s1 / (s2 + 1e-10)


def aggregate_torch(data, scale):
    """sum over step x step blocks / (count(data >= 0) + 1e-10), squeezed (aggregate_utils.py:29-41).
    `data` is [1,1,H,W] (or any [...,H,W]); step = int(1/scale)."""
    step = int(1 / scale)
    return _block_aggregate(data, step, 0.0, strict=False).squeeze()


def aggregate_torch_gpu(data, scale, device='cuda'):
    """aggregate_utils.py:44-59: mask is data > 1.0, no squeeze."""
    step = int(1 / scale)
    return _block_aggregate(data.to(device), step, 1.0, strict=True)


def aggregate(data, scale):
    """aggregate_utils.py:11-26: numpy loop version (mean over pixels > 0; output is square with
    int(rows*scale) cells per side, like the reference)."""
    r, c = data.shape
    nr, nc = int(r * scale), int(r * scale)
    step = int(1 / scale)
    res = np.zeros((nr, nc))
    data = data.astype('float')
    for i in range(0, r, step):
        for j in range(0, c, step):
            patch = data[i:i + step, j:j + step]
            res[int(i / step), int(j / step)] = patch.sum() / ((patch > 0).sum() + 1e-6)
    return res

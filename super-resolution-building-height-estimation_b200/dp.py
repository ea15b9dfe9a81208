"""Data-parallel harness around the hot path: one process per GPU,
`torch.distributed` for plumbing, ONE collective per training step.

* `MSE_adapt_weight`, `CE_DICE_adapt_weight` — the uncertainty-weighted losses applied right
  after the hot path (losses_pytorch/selfloss.py:81-90, 145-168; the reference hard-codes
  device="cuda" for log_var, here it follows the `device` argument).
* `FlatGradAllReduce` — all trainable gradients live in one flat fp32 bucket (each `p.grad` is a
  view into it), so a step needs a single NCCL all-reduce over NVLink (sum, then 1/world).
* `train_step` — train.py:243-257 for one batch: frozen RRDBNet features under no_grad, head
  forward, three weighted losses, backward, all-reduce, optimiser step.  No per-step host sync.
* `predict_shard` — predict_realesanet_feature_globe.py:167-177 for the tiles of one rank:
  features, head, clamp/round to uint16 (height*10) and softmax*255 (uint16), sharded rank::world.
* `adjust_learning_rate`, `build_training_state`, `save_checkpoint`, `load_checkpoint` — the
  epoch-level bookkeeping of train.py (66-80, 150-179, 198-212): step LR schedule (with the
  reference's "every group is rescaled" behaviour), Adam + log-variance parameter group, the
  `checkpoint.tar` schema and resume.
* `grid_positions`, `CityMosaic` — the host side of the city sweep (BH_loader.py:908-929,
  predict_realesanet_feature_globe.py:157-204): grid-cell windows from the geotransform, uint16
  overlap-add mosaics with visit counts, per-rank partial rasters merged by summation.
* `hierweight`, `make_labels` — the loader's label side on the device (BH_loader.py:30-61, 327-392):
  level weights from the height histogram, height-level LUT, per-pixel weights, 4x4 aggregates.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn


FUSED_CE_DICE = os.environ.get("BHSR_FUSED_CE_DICE", "1") != "0"   # 0: keep CE_DICE_adapt_weight on stock torch ops


class Dice(nn.Module):
    """losses_pytorch/selfloss.py:6-17."""

    def forward(self, pred, target):
        smooth = 1.
        num = pred.size(0)
        m1 = pred.reshape(num, -1)
        m2 = target.reshape(num, -1)
        intersection = (m1 * m2).sum()
        return 1 - (2. * intersection + smooth) / (m1.sum() + m2.sum() + smooth)


class _WeightedMSEFn(torch.autograd.Function):
    """`bhsr_weighted_mse` (post.cu): loss and both gradients from ONE pass over prediction / target / weight."""

    @staticmethod
    def forward(ctx, pred, target, weight, log_var):
        import ctypes as C  # noqa: F401
        from . import _lib
        p = pred.contiguous().float()
        t = target.expand_as(pred).contiguous().float()
        w = weight.expand_as(pred).contiguous().float()
        dev = p.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[3]
        gp = torch.empty_like(p) if ctx.needs_input_grad[0] else None
        gs = torch.empty((), dtype=torch.float32, device=dev) if need else None
        scratch = torch.empty(1, dtype=torch.float64, device=dev)
        with _lib.on_device(p):
            _lib.check(_lib.load().bhsr_weighted_mse(p.data_ptr(), t.data_ptr(), w.data_ptr(), p.numel(),
                                                     log_var.data_ptr(), loss.data_ptr(), _lib.ptr(gp), _lib.ptr(gs),
                                                     scratch.data_ptr(), _lib.stream_ptr(dev)), "bhsr_weighted_mse")
        ctx.save_for_backward(gp, gs)
        ctx.shape = pred.shape
        return loss

    @staticmethod
    def backward(ctx, go):
        gp, gs = ctx.saved_tensors
        return (gp.view(ctx.shape) * go if gp is not None else None, None, None, gs * go if gs is not None else None)


class MSE_adapt_weight(nn.Module):
    """mean(weight * (x - y)^2) * exp(-log_var) + log_var   (selfloss.py:81-90).  CUDA tensors run the fused
    forward+backward kernel (`bhsr_weighted_mse`); host tensors (loader-side checks, gloo tests) the same formula in
    torch."""

    def __init__(self, log_var=0.0, device="cuda"):
        super().__init__()
        self.log_var = nn.Parameter(torch.tensor(float(log_var), device=device))

    def forward(self, inputs, targets, weight):
        if inputs.is_cuda and inputs.dtype == torch.float32 and self.log_var.is_cuda:
            return _WeightedMSEFn.apply(inputs, targets, weight, self.log_var)
        loss = F.mse_loss(inputs, targets, reduction='none')
        loss = (loss * weight).mean()
        return loss * torch.exp(-self.log_var) + self.log_var


class _CEDiceFn(torch.autograd.Function):
    """`bhsr_ce_dice` (post.cu): weighted cross-entropy + Dice, loss and both gradients in two passes over the logits
    (the Dice gradient needs the global sums of the first pass)."""

    @staticmethod
    def forward(ctx, logits, labels, weight, log_var):
        from . import _lib
        z = logits.contiguous().float()
        t = labels.contiguous()
        w = weight.contiguous().float()
        nb, c, h, wd = z.shape
        dev = z.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[3]
        gz = torch.empty_like(z) if ctx.needs_input_grad[0] else None
        gs = torch.empty((), dtype=torch.float32, device=dev) if need else None
        scratch = torch.empty(4, dtype=torch.float64, device=dev)
        with _lib.on_device(z):
            _lib.check(_lib.load().bhsr_ce_dice(z.data_ptr(), t.data_ptr(), w.data_ptr(), nb, c, h, wd,
                                                log_var.data_ptr(), loss.data_ptr(), _lib.ptr(gz), _lib.ptr(gs),
                                                scratch.data_ptr(), _lib.stream_ptr(dev)), "bhsr_ce_dice")
        ctx.save_for_backward(gz, gs)
        return loss

    @staticmethod
    def backward(ctx, go):
        gz, gs = ctx.saved_tensors
        return (gz * go if gz is not None else None, None, None, gs * go if gs is not None else None)


class CE_DICE_adapt_weight(nn.Module):
    """weighted CE + Dice on P(class>0), uncertainty weighted (selfloss.py:145-168).  CUDA fp32 logits [N,C,H,W] with
    int64 labels and weights [N,H,W] (what train.py:253 passes) run the fused forward+backward kernel (`bhsr_ce_dice`,
    C <= 16); anything else (host tensors, other shapes) the same formula in torch."""

    def __init__(self, log_var=0.0, device="cuda"):
        super().__init__()
        self.ce = nn.CrossEntropyLoss(reduction='none')
        self.dice = Dice()
        self.log_var = nn.Parameter(torch.tensor(float(log_var), device=device))

    def forward(self, pmask, rmask, weight):
        if (FUSED_CE_DICE and pmask.is_cuda and pmask.dtype == torch.float32 and pmask.dim() == 4 and
                2 <= pmask.shape[1] <= 16 and pmask.numel() > 0 and rmask.dtype == torch.int64 and
                rmask.shape == (pmask.shape[0],) + pmask.shape[2:] and weight.shape == rmask.shape and
                weight.is_cuda and rmask.is_cuda and self.log_var.is_cuda and not weight.requires_grad):
            return _CEDiceFn.apply(pmask, rmask, weight, self.log_var)
        loss_ce = (self.ce(pmask, rmask) * weight).mean()
        p = pmask.softmax(dim=1)[:, 1:].sum(dim=1)
        loss = loss_ce + self.dice(p, (rmask > 0))
        return loss * torch.exp(-self.log_var) + self.log_var


def shard_indices(n: int, rank: int, world: int) -> range:
    """Tiles of rank `rank`: n items dealt round-robin (rank::world), SURVEY §8e."""
    return range(rank, n, world)


def broadcast_module(module: nn.Module, src: int = 0, group=None) -> None:
    """DDP-style start: parameters and buffers of rank `src` to every rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


class FlatGradAllReduce:
    """One flat fp32 gradient bucket for `params`; `p.grad` are views into it."""

    def __init__(self, params: Iterable[nn.Parameter], group=None):
        self.params: List[nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.group = group
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatGradAllReduce expects fp32 parameters on one device")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    @property
    def numel(self) -> int:
        return self.flat.numel()

    def zero_(self) -> None:
        """Use instead of optimizer.zero_grad(): keeps the views, zeroes the bucket."""
        self.flat.zero_()
        for p in self.params:  # an optimiser may have set grads to None
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or \
                    p.grad.data_ptr() >= self.flat.data_ptr() + self.flat.numel() * 4:
                raise RuntimeError("a parameter's .grad left the flat bucket; call optimizer.zero_grad(set_to_none=False) "
                                   "or FlatGradAllReduce.zero_() only")

    def all_reduce(self) -> None:
        """Sum over ranks then average — the step's single collective."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.group)
            if world > 1:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.mul_(1.0 / world)


def train_step(net_g, net, criterion: Sequence[nn.Module], optimizer, bucket: Optional[FlatGradAllReduce],
               lr: torch.Tensor, height: torch.Tensor, height_aggre: torch.Tensor, build: torch.Tensor,
               weight: torch.Tensor, weight_aggre: torch.Tensor, rgbseq=(0, 1, 2)) -> torch.Tensor:
    """One iteration of train_epoch_aggre_weight (train.py:243-257); returns the (device) loss."""
    with torch.no_grad():
        hr_fea = net_g.forward_feature(lr[:, list(rgbseq)])
    height_pred, build_pred, height_pred_aggre = net(lr, hr_fea)
    height_pred = height_pred.squeeze(1)
    height_pred_aggre = height_pred_aggre.squeeze(1)
    loss = criterion[0](height_pred, height, weight) + \
        criterion[1](height_pred_aggre, height_aggre, weight_aggre) + \
        criterion[2](build_pred, build, weight)
    if bucket is not None:
        bucket.zero_()
    else:
        optimizer.zero_grad()
    loss.backward()
    if bucket is not None:
        bucket.all_reduce()
    optimizer.step()
    return loss.detach()


class GraphedTrainStep:
    """`train_step` replayed from CUDA graphs.  One iteration is ~4000 kernel launches, most of them the tiny
    kernels of the smp encoder / decoders; captured once, a step costs two graph launches:
        graph A: frozen RRDBNet features, head forward, the three losses, backward into the flat gradient bucket
        (eager)  the step's single NCCL all-reduce of the bucket (world > 1)
        graph B: optimiser step
    The optimiser must be built with `capturable=True` (its step counter lives on the device).  Inputs are copied
    into static buffers before each replay.  Tensors produced inside a graph (the loss) are overwritten by the next
    replay.

    `overlap_smp` (default on, BHSR_TRAIN_OVERLAP=0 disables): inside graph A the smp encoder / decoders (hundreds of
    small stock-PyTorch kernels that leave most SMs idle) run on a forked stream next to the frozen RRDBNet forward
    (148-CTA tensor-core kernels); both feed `forward_head`.  Same arithmetic, fewer exposed microseconds.

    `prefetch` ("RRDBNet is frozen, so its forward for step t+1 is independent of step t's weight update"):
    call the step with `lr_next=<the next batch's tiles>` and the frozen features of the NEXT batch are computed on a
    forked stream during THIS step's head forward / backward (a second captured graph; the features wait in a static
    buffer).  A call whose `lr` was not announced by the previous call computes its features first.  Every step still
    runs exactly one RRDBNet forward.  OPT-IN (prefetch=True / BHSR_TRAIN_PREFETCH=1): measured SLOWER on a B200
    (62.9 -> 66.9 ms per step, same losses): the trunk's persistent CTAs take a whole SM each (352 threads x 168
    registers, 227 KB of shared memory), so while one of its kernels is resident the backward's kernels queue behind
    it instead of filling idle SMs, and the critical path lengthens by more than the 15 ms it sheds."""

    def __init__(self, net_g, net, criterion, optimizer, bucket: "FlatGradAllReduce", example, warmup: int = 3,
                 overlap_smp: Optional[bool] = None, prefetch: Optional[bool] = None):
        import os
        if overlap_smp is None:
            overlap_smp = os.environ.get("BHSR_TRAIN_OVERLAP", "1") != "0"
        if prefetch is None:
            prefetch = os.environ.get("BHSR_TRAIN_PREFETCH", "0") == "1"
        self.overlap_smp = bool(overlap_smp) and hasattr(net, "forward_smp")
        self.prefetch = bool(prefetch) and self.overlap_smp
        self._fork = torch.cuda.Stream() if self.overlap_smp else None
        self._fork2 = torch.cuda.Stream() if self.prefetch else None
        self.bucket, self.optimizer = bucket, optimizer
        self.static = [t.clone() for t in example]          # lr, height, height_aggre, build, weight, weight_aggre
        self._args = (net_g, net, criterion)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                          # allocator / cache warm-up outside the capture
                self._fwd_bwd()
                bucket.all_reduce()
                optimizer.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_a):
            self.loss = self._fwd_bwd()
        self.graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_b):
            optimizer.step()
        self.graph_p = None            # pipelined variant of graph A, captured on first use
        self._announced = None         # (data_ptr, version) of the batch whose features sit in self._hr_next

    def _head_loss_backward(self, height_fea, build_fea, hr_fea):
        net_g, net, criterion = self._args
        lr, height, height_aggre, build, weight, weight_aggre = self.static
        height_pred, build_pred, height_pred_aggre = net.forward_head(height_fea, build_fea, hr_fea)
        loss = criterion[0](height_pred.squeeze(1), height, weight) + \
            criterion[1](height_pred_aggre.squeeze(1), height_aggre, weight_aggre) + \
            criterion[2](build_pred, build, weight)
        self.bucket.zero_()
        loss.backward()
        return loss.detach()

    def _fwd_bwd(self):
        net_g, net, criterion = self._args
        lr, height, height_aggre, build, weight, weight_aggre = self.static
        if self.overlap_smp:
            cur = torch.cuda.current_stream()
            self._fork.wait_stream(cur)
            with torch.cuda.stream(self._fork):
                height_fea, build_fea = net.forward_smp(lr)
            with torch.no_grad():
                hr_fea = net_g.forward_feature(lr[:, :3])
            cur.wait_stream(self._fork)
            height_fea.record_stream(cur)
            build_fea.record_stream(cur)
            return self._head_loss_backward(height_fea, build_fea, hr_fea)
        with torch.no_grad():
            # train.py:244 indexes lr[:, [0, 1, 2]]; a Python index list becomes a host tensor copied to the device,
            # which a graph capture refuses — the equivalent strided view needs no copy at all
            hr_fea = net_g.forward_feature(lr[:, :3])
        height_pred, build_pred, height_pred_aggre = net(lr, hr_fea)
        loss = criterion[0](height_pred.squeeze(1), height, weight) + \
            criterion[1](height_pred_aggre.squeeze(1), height_aggre, weight_aggre) + \
            criterion[2](build_pred, build, weight)
        self.bucket.zero_()
        loss.backward()
        return loss.detach()

    def _fwd_bwd_pipelined(self):
        """Graph A with the RRDBNet forward taken out of the critical path: this step's features were computed during
        the previous replay (self._hr_next); the features of `self._lr_next` are computed on a forked stream now."""
        net_g, net, _ = self._args
        lr = self.static[0]
        cur = torch.cuda.current_stream()
        self._hr_cur.copy_(self._hr_next)                    # before the forked stream overwrites it
        self._fork2.wait_stream(cur)
        with torch.cuda.stream(self._fork2):
            with torch.no_grad():
                self._hr_next.copy_(net_g.forward_feature(self._lr_next[:, :3]))
        self._fork.wait_stream(cur)
        with torch.cuda.stream(self._fork):
            height_fea, build_fea = net.forward_smp(lr)
        cur.wait_stream(self._fork)
        height_fea.record_stream(cur)
        build_fea.record_stream(cur)
        loss = self._head_loss_backward(height_fea, build_fea, self._hr_cur)
        cur.wait_stream(self._fork2)
        return loss

    def _capture_pipelined(self):
        net_g = self._args[0]
        lr = self.static[0]
        with torch.no_grad():
            hr = net_g.forward_feature(lr[:, :3])
        self._hr_cur = torch.empty_like(hr)
        self._hr_next = hr.clone()
        self._lr_next = lr.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        state = [p.detach().clone() for p in self._args[1].parameters()]   # the warm-up pass below must not train
        bufs = [b.detach().clone() for b in self._args[1].buffers()]
        with torch.cuda.stream(side):
            self._fwd_bwd_pipelined()                        # warm-up of the forked streams outside the capture
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            for p, q in zip(self._args[1].parameters(), state):
                p.copy_(q)
            for b, q in zip(self._args[1].buffers(), bufs):
                b.copy_(q)
        self.graph_p = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_p):
            self.loss_p = self._fwd_bwd_pipelined()

    def __call__(self, lr, height, height_aggre, build, weight, weight_aggre, lr_next=None) -> torch.Tensor:
        for dst, src in zip(self.static, (lr, height, height_aggre, build, weight, weight_aggre)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        if self.prefetch and lr_next is not None:
            if self.graph_p is None:
                self._capture_pipelined()
            tag = (lr.data_ptr(), lr._version)
            if self._announced != tag:                       # not prefetched: compute this batch's features now
                with torch.no_grad():
                    self._hr_next.copy_(self._args[0].forward_feature(self.static[0][:, :3]))
            self._lr_next.copy_(lr_next, non_blocking=True)
            self._announced = (lr_next.data_ptr(), lr_next._version)
            self.graph_p.replay()
            loss = self.loss_p
        else:
            self._announced = None
            self.graph_a.replay()
            loss = self.loss
        self.bucket.all_reduce()
        self.graph_b.replay()
        return loss


@torch.no_grad()
def predict_shard(net_g, net, tiles: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """predict_realesanet_feature_globe.py:167-177 on a batch of tiles already on the GPU: features, head, then the
    fused post-processing kernel (`bhsr_predict_postproc`): height -> max(h, 0), round(h * 10) -> uint16; height
    levels -> softmax over the K channels, round(p * 255) -> uint16.  Returns (uint16 [B,1,256,256], uint16
    [B,K,256,256]) — the reference's numbers, one pass over the head's outputs."""
    if hasattr(net, "forward_smp") and tiles.is_cuda and not torch.cuda.is_current_stream_capturing():
        # the stock-PyTorch encoder / decoders (small kernels) on a forked stream next to the frozen RRDBNet forward
        cur = torch.cuda.current_stream()
        side = _side_stream(tiles.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            height_fea, build_fea = net.forward_smp(tiles)
        hr_fea = net_g.forward_feature(tiles[:, :3])
        cur.wait_stream(side)
        height_fea.record_stream(cur)
        build_fea.record_stream(cur)
        ypred, build_pred = net.forward_head(height_fea, build_fea, hr_fea)[:2]
    else:
        hr_fea = net_g.forward_feature(tiles[:, :3])
        ypred, build_pred = net(tiles, hr_fea)[:2]
    return predict_postprocess(ypred, build_pred)


_SIDE_STREAMS = {}


def _side_stream(device) -> "torch.cuda.Stream":
    key = str(device)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


@torch.no_grad()
def predict_postprocess(ypred: torch.Tensor, build_pred: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    from . import _lib
    _lib.require_cuda(ypred, "ypred")
    y = ypred.contiguous().float()
    b = build_pred.contiguous().float()
    nb, k, h, w = b.shape
    assert y.shape == (nb, 1, h, w)
    out_h = torch.empty((nb, 1, h, w), dtype=torch.uint16, device=y.device)
    out_b = torch.empty((nb, k, h, w), dtype=torch.uint16, device=y.device)
    with _lib.on_device(y):
        _lib.check(_lib.load().bhsr_predict_postproc(y.data_ptr(), b.data_ptr(), nb, k, h, w, out_h.data_ptr(),
                                                     out_b.data_ptr(), _lib.stream_ptr(y.device)),
                   "bhsr_predict_postproc")
    return out_h, out_b


# ------------------------------------------------------------------ epoch loops (row N1: train.py:225-343)
class _DeviceMeter:
    """AverageMeter (train.py:55-64) whose sums stay on the device: the reference calls `.item()` three times per
    iteration (train.py:260-265), each a host synchronisation; here the epoch ends with one."""

    def __init__(self, device):
        self.sum = torch.zeros((), dtype=torch.float64, device=device)
        self.count = 0

    def update(self, val: torch.Tensor, n: int) -> None:
        self.sum += val.detach().double() * n
        self.count += n

    @property
    def avg(self) -> float:
        return float(self.sum.item()) / max(self.count, 1)


def train_epoch_aggre_weight(net, net_g, criterion, dataloader, optimizer, device, rgbseq=(0, 1, 2),
                             bucket: Optional["FlatGradAllReduce"] = None):
    """train.py:225-271: one epoch of the isaggre=True recipe.  `dataloader` yields (lr, (height, height_aggre),
    build, (weight, weight_aggre)) like BH_loader.myImageFloder_S12_globe.  Returns (mean loss, mean RMSE,
    [log_var of each criterion]) — per-sample weighted by batch size like the reference's AverageMeter."""
    net.train()
    losses, acc = _DeviceMeter(device), _DeviceMeter(device)
    for lr, heightall, build, weightall in dataloader:
        lr = lr.to(device, non_blocking=True)
        height = heightall[0].to(device, non_blocking=True)
        height_aggre = heightall[1].to(device, non_blocking=True)
        build = build.to(device, non_blocking=True)
        weight = weightall[0].to(device, non_blocking=True)
        weight_aggre = weightall[1].to(device, non_blocking=True)
        with torch.no_grad():
            hr_fea = net_g.forward_feature(lr[:, list(rgbseq)])
        height_pred, build_pred, height_pred_aggre = net(lr, hr_fea)
        height_pred = height_pred.squeeze(1)
        height_pred_aggre = height_pred_aggre.squeeze(1)
        loss = criterion[0](height_pred, height, weight) + \
            criterion[1](height_pred_aggre, height_aggre, weight_aggre) + \
            criterion[2](build_pred, build, weight)
        if bucket is not None:
            bucket.zero_()
        else:
            optimizer.zero_grad()
        loss.backward()
        if bucket is not None:
            bucket.all_reduce()
        optimizer.step()
        bsize = lr.size(0)
        losses.update(loss, bsize)
        with torch.no_grad():
            acc.update(torch.sqrt(((height_pred - height) ** 2).mean()), bsize)
    lossweight = [float(c.log_var.item()) for c in criterion]
    return losses.avg, acc.avg, lossweight


@torch.no_grad()
def vtest_epoch(net, net_g, dataloader, device, rgbseq=(0, 1, 2)):
    """train.py:318-343: validation MSE / RMSE of the height output (eval mode: the head runs on the tensor-core
    eval path).  `dataloader` yields (x, y_true, _, _)."""
    net.eval()
    losses, acc = _DeviceMeter(device), _DeviceMeter(device)
    for x, y_true, _, _ in dataloader:
        x = x.to(device, non_blocking=True)
        y_true = y_true.to(device, non_blocking=True)
        hr_fea = net_g.forward_feature(x[:, list(rgbseq)])
        ypred = net(x, hr_fea)[0].squeeze(1)
        mse = torch.mean((ypred - y_true) ** 2)
        losses.update(mse, x.size(0))
        acc.update(torch.sqrt(mse), x.size(0))
    return losses.avg, acc.avg


def fit(net, net_g, train_loader, val_loader, logdir: str, epochs: int, init_lr: float = 1e-3, device="cuda",
        rgbseq=(0, 1, 2), bucket_params: bool = True, log=None):
    """The epoch loop of train.py:150-222 for the isaggre recipe: resume from `logdir/checkpoint.tar`, per-epoch
    step LR schedule, train epoch, validation, checkpoint (+ best / every-5th copies).  Gradients live in one flat
    bucket (`FlatGradAllReduce`), so under torch.distributed every step is ONE all-reduce.  Returns the per-epoch
    records [{'epoch', 'lr', 'train_loss', 'train_rmse', 'val_loss', 'val_rmse', 'log_vars'}]."""
    start_epoch, best_acc, log_vars = load_checkpoint(logdir, net, map_location=device)
    if best_acc is None:
        best_acc = 0            # train.py:108 (with this start `is_best` never fires; kept as in the reference)
    optimizer, criterion = build_training_state(net, init_lr, True, log_vars, device)
    bucket = None
    if bucket_params:
        params = [p for g in optimizer.param_groups for p in g['params']]
        bucket = FlatGradAllReduce(params)
    history = []
    for epoch in range(epochs):
        if epoch < start_epoch:
            continue
        epoch = epoch + 1
        lr = adjust_learning_rate(init_lr, epoch, optimizer)
        train_loss, train_rmse, lossweight = train_epoch_aggre_weight(net, net_g, criterion, train_loader, optimizer,
                                                                      device, rgbseq, bucket)
        val_loss, val_rmse = vtest_epoch(net, net_g, val_loader, device, rgbseq)
        best_acc, _ = save_checkpoint(logdir, epoch, net, lossweight, best_acc, val_rmse, True)
        rec = {'epoch': epoch, 'lr': lr, 'train_loss': train_loss, 'train_rmse': train_rmse, 'val_loss': val_loss,
               'val_rmse': val_rmse, 'log_vars': lossweight}
        history.append(rec)
        if log is not None:
            log(rec)
    return history


def hierweight(stats, hir: Sequence[int], mode: str = "sqrt") -> torch.Tensor:
    """Class weights over the height levels `hir` from a 256-bin height histogram, BH_loader.py:30-61:
    inverse square-root frequency (`hierweight`, the one train.py uses), inverse frequency
    (`hierweight_simple`) or all ones (`hierweight_equal`), normalised so the weights sum to the
    number of levels.  Known answer for the shipped globe statistics: BH_loader.py:1116-1124."""
    st = torch.as_tensor(stats, dtype=torch.float64)
    st = st / st.sum()
    n = len(hir) - 1
    if mode == "equal":
        return torch.ones(n, dtype=torch.float64)
    freq = torch.stack([st[hir[i]:hir[i + 1]].sum() for i in range(n)])
    w = 1.0 / torch.sqrt(freq) if mode == "sqrt" else 1.0 / freq
    w = w / w.sum()
    return w * (n / w.sum())


def make_labels(height: torch.Tensor, level_weights: torch.Tensor,
                hir: Sequence[int] = (0, 3, 12, 21, 30, 60, 90, 256), scale: float = 0.25):
    """The loader's label pipeline on the tensor's own device (BH_loader.py:327-329, 373-392):
    uint8-valued heights [B,256,256] -> height-level classes through the `hir` LUT, per-pixel loss
    weights, and the 4x4-aggregated height / weight maps [B,64,64] (`aggregate_torch`)."""
    from .aggregate import aggregate_torch
    dev = height.device
    lut = torch.zeros(256, dtype=torch.long)
    for i in range(len(hir) - 1):
        lut[hir[i]:hir[i + 1]] = i
    build = lut.to(dev)[height.long().clamp_(0, 255)]
    weight = level_weights.to(dev, torch.float32)[build]
    nb = height.shape[0]
    side = int(round(height.shape[-1] * scale))
    h_aggre = aggregate_torch(height.float().unsqueeze(1), scale).reshape(nb, side, side)
    w_aggre = aggregate_torch(weight.unsqueeze(1), scale).reshape(nb, side, side)
    return build, weight, h_aggre, w_aggre


def synthetic_labels(nb: int, device, stats: Optional[torch.Tensor] = None, seed: int = 0,
                     hir=(0, 3, 12, 21, 30, 60, 90, 256)):
    """Synthetic targets shaped like the loader's (BH_loader.py:327-392): uint8-valued heights
    (mostly zero), height-level classes via the `hir` LUT, inverse-sqrt-frequency weights, and the
    4x4 aggregated height / weight."""
    from .aggregate import aggregate_torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    h = torch.rand((nb, 256, 256), generator=g) * 60
    h = torch.where(torch.rand((nb, 256, 256), generator=g) < 0.8245, torch.zeros(()), h).floor()
    lut = torch.zeros(256, dtype=torch.long)
    for i in range(len(hir) - 1):
        lut[hir[i]:hir[i + 1]] = i
    build = lut[h.long()]
    if stats is None:
        w_levels = torch.tensor([0.08743518, 0.26821995, 0.32067124, 0.73515255, 0.98135007, 1.60267172, 3.0044993])
    else:
        w_levels = stats
    weight = w_levels[build]
    h, build, weight = h.to(device), build.to(device), weight.to(device)
    h_aggre = aggregate_torch(h.unsqueeze(1), 0.25).reshape(nb, 64, 64)
    w_aggre = aggregate_torch(weight.unsqueeze(1), 0.25).reshape(nb, 64, 64)
    return h, h_aggre, build, weight, w_aggre


# ------------------------------------------------------------------ training-loop bookkeeping (row N1)
def adjust_learning_rate(init_lr: float, epoch: int, optimizer) -> float:
    """train.py:66-80 — step schedule on the 1-based epoch: x1 up to epoch 10, x0.1 up to 20, x0.01
    after.  The reference means to skip the loss-weight group but tests `'lossweight' in
    param_group`, which looks for a KEY of that name and is always False, so every group (the
    log-var group too) is rescaled; that behaviour is kept."""
    lr = init_lr if epoch <= 10 else (0.1 * init_lr if epoch <= 20 else 0.01 * init_lr)
    for group in optimizer.param_groups:
        if ('lossweight' in group) and (group.get('name') == 'lossweight'):
            continue
        group['lr'] = lr
    return lr


def build_training_state(net: nn.Module, init_lr: float = 1e-3, isaggre: bool = True,
                         log_vars: Sequence[float] = (0.0, 0.0, 0.0), device="cuda"):
    """Adam(weight_decay=1e-4) over the head + the uncertainty-weighted criteria whose log-variances
    form a second parameter group `{'lr': 1e-3, 'name': 'lossweight'}` (train.py:170-179; the group
    inherits weight_decay=1e-4 exactly as in the reference)."""
    optimizer = torch.optim.Adam(net.parameters(), lr=init_lr, weight_decay=1e-4)
    criterion: List[nn.Module] = [MSE_adapt_weight(log_vars[0], device=device)]
    if isaggre:
        criterion.append(MSE_adapt_weight(log_vars[1], device=device))
    criterion.append(CE_DICE_adapt_weight(log_vars[2], device=device))
    optimizer.add_param_group({'params': [c.log_var for c in criterion], 'lr': 0.001, 'name': 'lossweight'})
    return optimizer, criterion


def save_checkpoint(logdir: str, epoch: int, net: nn.Module, log_vars, best_acc: float, val_rmse: float,
                    isaggre: bool = True) -> Tuple[float, bool]:
    """train.py:198-212 — `checkpoint.tar` every epoch with the reference's schema
    {'epoch','state_dict','log_vars','best_acc'} (optimizer state is NOT saved), a copy to
    `model_best.tar` when `val_rmse < best_acc`, and to `checkpoint{epoch}.tar` every 5 epochs.
    Returns (new best_acc, is_best).  With the reference's initial best_acc = 0 (train.py:108)
    `is_best` never becomes true; callers wanting a best model start from +inf."""
    import os
    import shutil
    os.makedirs(logdir, exist_ok=True)
    path = os.path.join(logdir, 'checkpoint.tar')
    is_best = val_rmse < best_acc
    best_acc = min(val_rmse, best_acc)
    module = net.module if hasattr(net, "module") else net
    torch.save({'epoch': epoch, 'state_dict': module.state_dict(),
                'log_vars': [float(v) for v in log_vars] if isaggre else 1.0, 'best_acc': best_acc}, path)
    if is_best:
        shutil.copy(path, os.path.join(logdir, 'model_best.tar'))
    if epoch % 5 == 0:
        shutil.copy(path, os.path.join(logdir, f'checkpoint{epoch}.tar'))
    return best_acc, is_best


def load_checkpoint(logdir: str, net: nn.Module, map_location=None):
    """train.py:150-168 — resume from `logdir/checkpoint.tar` if present: returns
    (start_epoch, best_acc, log_vars) and loads the head's state_dict; (0, None, [0,0,0]) otherwise."""
    import os
    path = os.path.join(logdir, 'checkpoint.tar')
    if not os.path.isfile(path):
        return 0, None, [0.0, 0.0, 0.0]
    ckpt = torch.load(path, map_location=map_location)
    net.load_state_dict(ckpt['state_dict'])
    log_vars = ckpt['log_vars']
    if not isinstance(log_vars, (list, tuple)):
        log_vars = [0.0, 0.0, 0.0]
    return int(ckpt['epoch']), ckpt['best_acc'], list(log_vars)


# ------------------------------------------------------------------ city mosaics (row N2)
def grid_positions(bounds: Iterable[Sequence[float]], transform: Sequence[float]) -> List[Tuple[int, int, int, int]]:
    """BH_loader.py:908-929 (`generateindex`) without the shapefile I/O: each grid cell given as
    (minX, minY, maxX, maxY) in map units becomes (xoff, yoff, xcount, ycount) in LR pixels of the
    raster with GDAL geotransform `transform` (Python's round-half-to-even, like the reference)."""
    x0, y0 = transform[0], transform[3]
    pw, ph = transform[1], -transform[5]
    pos = []
    for minx, miny, maxx, maxy in bounds:
        pos.append((round((minx - x0) / pw), round((y0 - maxy) / ph),
                    round((maxx - minx) / pw), round((maxy - miny) / ph)))
    return pos


class CityMosaic:
    """Host-side accumulation of per-tile predictions into city rasters,
    predict_realesanet_feature_globe.py:157-159, 179-185, 195-204: uint16 height*10 sums, uint16
    per-class softmax*255 sums, uint8 visit counts, all at 4x the LR grid; `finalize` returns the
    arg-max class map (uint8) and the visit-normalised height (uint16, only where visited).
    Integer widths and wrap-around are the reference's (numpy `+=` on uint16 / uint8)."""

    def __init__(self, lr_height: int, lr_width: int, chans_build: int, upscale: int = 4):
        import numpy as np
        self.np = np
        self.upscale = upscale
        h, w = lr_height * upscale, lr_width * upscale
        self.height = np.zeros((h, w), dtype=np.uint16)
        self.build = np.zeros((chans_build, h, w), dtype=np.uint16)
        self.weight = np.zeros((h, w), dtype=np.uint8)

    def add(self, ypred, build, positions) -> None:
        """ypred [n,1,H,W], build [n,K,H,W]: the integer-valued outputs of `predict_shard` (any
        integer dtype / tensor); positions: n x (xoff, yoff, xcount, ycount) in LR pixels."""
        np = self.np
        yp = np.asarray(ypred.cpu() if hasattr(ypred, "cpu") else ypred).astype(np.uint16)
        bp = np.asarray(build.cpu() if hasattr(build, "cpu") else build).astype(np.uint16)
        for i, pos in enumerate(positions):
            xoff, yoff, xcount, ycount = (int(v) * self.upscale for v in pos)
            self.height[yoff:yoff + ycount, xoff:xoff + xcount] += yp[i, 0, :ycount, :xcount]
            self.build[:, yoff:yoff + ycount, xoff:xoff + xcount] += bp[i, :, :ycount, :xcount]
            self.weight[yoff:yoff + ycount, xoff:xoff + xcount] += 1

    def merge(self, other: "CityMosaic") -> None:
        """Sum the partial rasters of another rank (sharded sweep: every tile is in exactly one)."""
        self.height += other.height
        self.build += other.build
        self.weight += other.weight

    def finalize(self):
        np = self.np
        build = np.argmax(self.build, axis=0).astype(np.uint8)
        height = self.height.copy()
        mask = self.weight > 0
        height[mask] = np.round(height[mask] / self.weight[mask]).astype(np.uint16)
        return build, height

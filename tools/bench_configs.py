#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs 3-5 (bench.py carries config 2, the headline).

  config 3: full SR + feature-aggregation head, fwd+bwd, batch 32, weighted losses, 1 GPU
  config 4: the same step as a data-parallel training loop (32 tiles / GPU, Adam, one NCCL
            all-reduce of the flat gradient bucket per step) — run under torchrun
  config 5: inference sweep, 10k synthetic 64x64 grids, batch 128, sharded rank::world
  config 6: SR fine-tune step: RealESRGAN(is_train=True).optimize_parameters — generator forward,
            L1 pixel loss, tensor-core backward (rrdbnet_train.py), Adam, EMA — batch 12 (the upstream YAML's),
            64x64 -> 256x256; the discriminator / VGG terms are stock PyTorch and not plugged in

Prints one JSON line per config (rank 0).  CUDA events, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5, 6])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--grids", type=int, default=10000)
    ap.add_argument("--numerics", default="exact")
    ap.add_argument("--num-block", type=int, default=23)
    args = ap.parse_args()

    import torch
    import bhsr  # noqa: F401
    from bhsr import dp
    from bhsr.models import SRRegress_Cls_feature
    from bhsr.rrdbnet import RRDBNet
    import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    if args.config == 6:
        from bhsr.rrdbnet import RealESRGAN
        torch.manual_seed(1337)
        B = args.batch or 12
        m = RealESRGAN(device=str(dev), num_block=args.num_block, is_train=True, ema_decay=0.999)
        m.feed_data({"lq": torch.rand(B, 3, 64, 64), "gt": torch.rand(B, 3, 256, 256)})
        m.use_cuda_graph = os.environ.get("BHSR_SR_GRAPH", "1") != "0"
        if m.use_cuda_graph:
            for _ in range(3):       # two eager warm-up steps, then the capture
                m.optimize_parameters()
        out = {}
        for _ in range(args.warmup):
            out = m.optimize_parameters()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = m.optimize_parameters()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"config": 6, "metric": "SR fine-tune generator step (fwd + L1 + tensor-core bwd + Adam + EMA)",
                          "value": B / ms * 1e3, "unit": "tiles/s", "ms_per_step": ms, "batch": B,
                          "num_block": args.num_block, "l_g_pix": out.get("l_g_pix"), "numerics": "exact", "launch": out.get("launch", "eager"),
                          "note": "reference-derived figure (BASELINE.md §2): 0.688 s/iter for the full G+D step, batch 12, unknown GPU"}),
              flush=True)
        return

    torch.manual_seed(1337)
    net_g = RRDBNet(3, 3, scale=4, num_feat=64, num_block=args.num_block, num_grow_ch=32).to(dev).eval()
    net_g.numerics = args.numerics
    for p in net_g.parameters():
        p.requires_grad = False
    isaggre = args.config != 5  # the predict script builds the head without the aggregation branch
    net = SRRegress_Cls_feature("efficientnet-b4", encoder_weights=None, in_channels=8, super_in=64, super_mid=16,
                                upscale=4, isaggre=isaggre, chans_build=7).to(dev)

    if os.environ.get("BHSR_SMP_NHWC", "1") != "0":
        net.smp_channels_last()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    if args.config in (3, 4):
        B = args.batch or 32
        net.train()
        dp.broadcast_module(net)
        crit = [dp.MSE_adapt_weight(0.0, dev), dp.MSE_adapt_weight(0.0, dev), dp.CE_DICE_adapt_weight(0.0, dev)]
        params = list(net.parameters()) + [c.log_var for c in crit]
        opt = torch.optim.Adam([{"params": list(net.parameters())}, {"params": [c.log_var for c in crit], "name": "lossweight"}],
                               lr=1e-3, weight_decay=1e-4)
        bucket = dp.FlatGradAllReduce(params)
        x = torch.from_numpy(synth.tiles(B, 8, seed=1337 + rank)).to(dev)
        h, h_aggre, build, w, w_aggre = dp.synthetic_labels(B, dev, seed=rank)
        loss_box = {}

        def step(i):
            loss_box["loss"] = dp.train_step(net_g, net, crit, opt, bucket, x, h, h_aggre, build, w, w_aggre)

        for i in range(args.warmup):
            step(i)
        ms = timed(step, args.steps)
        loss = float(loss_box["loss"].item())
        line = {"config": args.config, "metric": "tiles/sec fwd+bwd (RRDBNet-23 features + head + losses + Adam)",
                "value": B * world * args.steps / ms * 1e3, "unit": "tiles/s", "n_gpus": world, "batch_per_gpu": B,
                "ms_per_step": ms / args.steps, "numerics": args.numerics, "loss": loss,
                "grad_bucket_floats": bucket.numel, "collective": "1 x NCCL all-reduce / step" if world > 1 else "none (1 GPU)"}
    else:
        B = args.batch or 128
        net.eval()
        n_grids = args.grids
        mine = list(dp.shard_indices((n_grids + B - 1) // B, rank, world))  # batches of this rank
        x = torch.from_numpy(synth.tiles(B, 8, seed=7 + rank)).to(dev)
        host_h = torch.empty((B, 1, 256, 256), dtype=torch.uint16).pin_memory()

        def sweep(_):
            for _b in mine:
                ypred, build = dp.predict_shard(net_g, net, x)
                host_h.copy_(ypred, non_blocking=True)   # the mosaic accumulation is host-side (predict...:181-185)
            torch.cuda.current_stream().synchronize()

        sweep(0)
        ms = timed(sweep, 1)
        line = {"config": 5, "metric": "tiles/sec inference sweep (features + head + uint16 post-processing)",
                "value": len(mine) * B * world / ms * 1e3 if world == 1 else None, "unit": "tiles/s", "n_gpus": world,
                "batch": B, "grids": n_grids, "ms_total": ms, "numerics": args.numerics}
        if world > 1:
            line["value"] = n_grids / ms * 1e3
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

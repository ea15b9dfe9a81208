"""Host-side logic of the data-parallel harness on CPU: world_size-2 gloo processes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bhsr  # noqa: F401
    from bhsr import dp
    torch.manual_seed(100 + rank)          # different init per rank on purpose
    net = nn.Sequential(nn.Linear(6, 5), nn.BatchNorm1d(5), nn.Linear(5, 1))
    dp.broadcast_module(net, src=0)
    bucket = dp.FlatGradAllReduce(net.parameters())
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    crit = dp.MSE_adapt_weight(0.0, device="cpu")
    torch.manual_seed(7)
    x_all, y_all = torch.randn(8, 6), torch.randn(8, 1)
    idx = list(dp.shard_indices(8, rank, world))
    for _ in range(3):
        bucket.zero_()
        loss = crit(net(x_all[idx]), y_all[idx], torch.ones(len(idx), 1))
        loss.backward()
        bucket.all_reduce()
        opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        out.put(([g.numpy() for g in gathered], bucket.numel, idx))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_all_reduce_keeps_replicas_in_sync():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    params, numel, idx0 = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(params[0], params[1]), "replicas diverged: gradients were not averaged identically"
    assert numel == 6 * 5 + 5 + 5 + 5 + 5 + 1 and idx0 == [0, 2, 4, 6]


def test_losses_match_their_closed_forms():
    import bhsr  # noqa: F401
    from bhsr import dp
    from oracle import ref_numpy as R
    rng = np.random.RandomState(0)
    pred, tgt, w = (rng.rand(2, 8, 8).astype(np.float32) for _ in range(3))
    crit = dp.MSE_adapt_weight(0.3, device="cpu")
    got = crit(torch.from_numpy(pred), torch.from_numpy(tgt), torch.from_numpy(w)).item()
    assert abs(got - R.weighted_mse_adapt(pred, tgt, w, 0.3)) < 1e-6
    logits = torch.from_numpy(rng.standard_normal((2, 7, 8, 8)).astype(np.float32))
    labels = torch.from_numpy(rng.randint(0, 7, (2, 8, 8)))
    ce = dp.CE_DICE_adapt_weight(0.0, device="cpu")
    val = ce(logits, labels, torch.from_numpy(w))
    p = logits.softmax(1)[:, 1:].sum(1)
    dice = 1 - (2 * (p * (labels > 0)).sum() + 1) / (p.sum() + (labels > 0).sum() + 1)
    ref = (torch.nn.functional.cross_entropy(logits, labels, reduction="none") * torch.from_numpy(w)).mean() + dice
    assert abs(val.item() - ref.item()) < 1e-6


@pytest.mark.parametrize("c,log_var", [(7, 0.3), (2, -0.5), (16, 0.0)])
def test_ce_dice_oracle_gradients_match_torch_autograd(c, log_var):
    """oracle.ce_dice_adapt_weight (the closed-form gradients the fused `bhsr_ce_dice` kernel implements,
    selfloss.py:145-168) against stock autograd of the reference's formula in fp64, and the host path of
    dp.CE_DICE_adapt_weight against both."""
    import bhsr  # noqa: F401
    from bhsr import dp
    from oracle import ref_numpy as R
    rng = np.random.RandomState(c)
    z = (rng.standard_normal((3, c, 9, 5)) * 3).astype(np.float32)
    t = (rng.randint(0, c, (3, 9, 5)) * (rng.rand(3, 9, 5) > 0.5)).astype(np.int64)
    w = (0.1 + 3 * rng.rand(3, 9, 5)).astype(np.float32)
    zd = torch.from_numpy(z).double().requires_grad_(True)
    lv = torch.tensor(float(log_var), dtype=torch.float64, requires_grad=True)
    ce = (torch.nn.functional.cross_entropy(zd, torch.from_numpy(t), reduction="none") * torch.from_numpy(w).double()).mean()
    p = zd.softmax(dim=1)[:, 1:].sum(dim=1)
    m2 = (torch.from_numpy(t) > 0).double()
    ref = (ce + 1 - (2.0 * (p * m2).sum() + 1.0) / (p.sum() + m2.sum() + 1.0)) * torch.exp(-lv) + lv
    ref.backward()
    loss, grad, glv = R.ce_dice_adapt_weight(z, t, w, log_var)
    assert abs(loss - ref.item()) < 1e-12
    np.testing.assert_allclose(grad, zd.grad.numpy(), rtol=1e-10, atol=1e-15)
    assert abs(glv - lv.grad.item()) < 1e-12
    crit = dp.CE_DICE_adapt_weight(log_var, device="cpu")
    zc = torch.from_numpy(z).requires_grad_(True)
    val = crit(zc, torch.from_numpy(t), torch.from_numpy(w))
    val.backward()
    assert abs(val.item() - loss) < 1e-5 * max(1.0, abs(loss))
    np.testing.assert_allclose(zc.grad.numpy(), grad, rtol=1e-4, atol=1e-7)


def test_bucket_guards_against_detached_grads():
    import bhsr  # noqa: F401
    from bhsr import dp
    lin = nn.Linear(3, 2)
    b = dp.FlatGradAllReduce(lin.parameters())
    lin(torch.ones(1, 3)).sum().backward()
    assert b.flat.abs().sum() > 0 and lin.weight.grad.data_ptr() == b.flat.data_ptr()
    torch.optim.SGD(lin.parameters(), lr=0.1).zero_grad(set_to_none=True)
    with pytest.raises(RuntimeError):
        b.zero_()


def test_lr_schedule_rescales_every_group_like_the_reference():
    """train.py:66-80: 1e-3 -> 1e-4 after epoch 10 -> 1e-5 after epoch 20, and (reference quirk) the
    'lossweight' group is rescaled too."""
    import torch
    from bhsr import dp
    net = torch.nn.Linear(4, 2)
    opt, crit = dp.build_training_state(net, init_lr=1e-3, isaggre=True, device="cpu")
    assert len(crit) == 3 and len(opt.param_groups) == 2
    assert opt.param_groups[1]['name'] == 'lossweight' and opt.param_groups[1]['weight_decay'] == 1e-4
    seen = {e: dp.adjust_learning_rate(1e-3, e, opt) for e in (1, 10, 11, 20, 21, 30)}
    assert seen == {1: 1e-3, 10: 1e-3, 11: 1e-4, 20: 1e-4, 21: 1e-5, 30: 1e-5}
    assert [g['lr'] for g in opt.param_groups] == [1e-5, 1e-5]


def test_checkpoint_schema_and_resume(tmp_path):
    """train.py:150-168, 198-212: checkpoint.tar keys, 5-epoch copies, best-model rule, resume."""
    import os
    import torch
    from bhsr import dp
    net = torch.nn.Linear(4, 2)
    best = 0.0                                   # the reference's initial value: never "best"
    best, is_best = dp.save_checkpoint(str(tmp_path), 4, net, [0.1, 0.2, 0.3], best, val_rmse=7.0)
    assert not is_best and best == 0.0 and not os.path.exists(tmp_path / "model_best.tar")
    best, is_best = dp.save_checkpoint(str(tmp_path), 5, net, [0.1, 0.2, 0.3], float("inf"), val_rmse=7.0)
    assert is_best and best == 7.0
    assert os.path.exists(tmp_path / "model_best.tar") and os.path.exists(tmp_path / "checkpoint5.tar")
    ckpt = torch.load(tmp_path / "checkpoint.tar")
    assert sorted(ckpt) == ['best_acc', 'epoch', 'log_vars', 'state_dict'] and ckpt['epoch'] == 5
    other = torch.nn.Linear(4, 2)
    epoch, best_acc, log_vars = dp.load_checkpoint(str(tmp_path), other)
    assert (epoch, best_acc) == (5, 7.0) and log_vars == pytest.approx([0.1, 0.2, 0.3])
    assert torch.equal(other.weight, net.weight)
    assert dp.load_checkpoint(str(tmp_path / "missing"), other) == (0, None, [0.0, 0.0, 0.0])


def test_grid_positions_and_city_mosaic_match_the_reference_arithmetic():
    """BH_loader.py:908-929 and predict_realesanet_feature_globe.py:157-204 restated directly."""
    from bhsr import dp
    transform = (500000.0, 10.0, 0.0, 4100000.0, 0.0, -10.0)        # 10 m Sentinel-2 grid
    bounds = [(500000.0, 4099360.0, 500640.0, 4100000.0),           # 64x64 cell at the origin
              (500320.0, 4099040.0, 500960.0, 4099680.0),           # overlaps the first one
              (500645.0, 4099365.0, 501285.0, 4100005.0)]           # off-grid: rounding
    pos = dp.grid_positions(bounds, transform)
    assert pos[0] == (0, 0, 64, 64) and pos[1] == (32, 32, 64, 64)
    assert pos[2] == (round(64.5), round(-0.5), 64, 64) == (64, 0, 64, 64)   # half-to-even, as Python's round
    rng = np.random.RandomState(3)
    n, k = 2, 7
    yp = rng.randint(0, 900, size=(n, 1, 256, 256))
    bp = rng.randint(0, 256, size=(n, k, 256, 256))
    m = dp.CityMosaic(128, 128, k)
    m.add(torch.from_numpy(yp).to(torch.int32), torch.from_numpy(bp).to(torch.int32), pos[:2])
    # direct restatement
    h = np.zeros((512, 512), np.uint16); b = np.zeros((k, 512, 512), np.uint16); w = np.zeros((512, 512), np.uint8)
    for i, (xo, yo, xc, yc) in enumerate(np.array(pos[:2]) * 4):
        h[yo:yo + yc, xo:xo + xc] += yp[i, 0, :yc, :xc].astype(np.uint16)
        b[:, yo:yo + yc, xo:xo + xc] += bp[i, :, :yc, :xc].astype(np.uint16)
        w[yo:yo + yc, xo:xo + xc] += 1
    build, height = m.finalize()
    assert np.array_equal(build, np.argmax(b, axis=0).astype(np.uint8))
    ref_h = h.copy(); mask = w > 0
    ref_h[mask] = np.round(h[mask] / w[mask]).astype(np.uint16)
    assert np.array_equal(height, ref_h) and int(w.max()) == 2 and height[-1, -1] == 0
    # two ranks, one tile each, merged == one rank with both
    a, c = dp.CityMosaic(128, 128, k), dp.CityMosaic(128, 128, k)
    a.add(yp[:1], bp[:1], pos[:1]); c.add(yp[1:], bp[1:], pos[1:2]); a.merge(c)
    b2, h2 = a.finalize()
    assert np.array_equal(b2, build) and np.array_equal(h2, height)


def test_hierweight_known_answer_and_label_pipeline():
    """BH_loader.py:1116-1124: the shipped globe histogram gives the published level weights; the
    label pipeline maps heights to levels / weights / 4x4 aggregates like the loader (:327-392)."""
    from bhsr import dp
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
    w = dp.hierweight(golden["bh_stats_globe"], (0, 3, 12, 21, 30, 60, 90, 255))
    np.testing.assert_allclose(w.numpy(), golden["hierweight_kat"], rtol=0, atol=5e-9)
    assert dp.hierweight(golden["bh_stats_globe"], (0, 3, 12, 21, 30, 60, 90, 255), "equal").tolist() == [1.0] * 7
    ws = dp.hierweight(golden["bh_stats_globe"], (0, 3, 12, 21, 30, 60, 90, 255), "simple")
    assert abs(float(ws.sum()) - 7.0) < 1e-9 and bool((ws[1:] >= ws[:-1]).all())   # rarer levels weigh more
    h = torch.tensor([[0, 2, 3, 11], [12, 29, 30, 59], [60, 89, 90, 255], [0, 0, 0, 0]], dtype=torch.float32)
    h = h.repeat_interleave(4, 0).repeat_interleave(4, 1).unsqueeze(0)              # [1,16,16], 4x4 blocks
    build, weight, h_aggre, w_aggre = dp.make_labels(h, w, scale=0.25)
    assert build[0, ::4, ::4].tolist() == [[0, 0, 1, 1], [2, 3, 4, 4], [5, 5, 6, 6], [0, 0, 0, 0]]
    np.testing.assert_allclose(weight[0, ::4, ::4].numpy(), w.float()[build[0, ::4, ::4]].numpy())
    np.testing.assert_allclose(h_aggre[0].numpy(), h[0, ::4, ::4].numpy(), atol=1e-5)   # constant blocks
    np.testing.assert_allclose(w_aggre[0].numpy(), weight[0, ::4, ::4].numpy(), atol=1e-6)

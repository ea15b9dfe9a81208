"""Parity against the UNMODIFIED reference run LIVE from the staged copy baseline/_ref (SURVEY.md §8c "Staging").

`oracle/stage_reference.py` copies the reference's own source files of the hot path into baseline/_ref (git-ignored,
shipped to the GPU box with the working tree).  Where it is present these tests run the reference modules themselves —
no golden subsample, no port in between:

  not gpu : the oracle ports (oracle/ref_torch.py, oracle/ref_numpy.py) against the live reference on fresh seeds the
            committed goldens do not contain (bit-exact for the torch port: same ATen calls in the same order);
  gpu     : the CUDA path against the live reference, FULL tensors at the north-star tolerance
            |got - ref| <= 1e-4 + 1e-3 |ref| — the 23-block network on tiles of a B=64 batch (reference in fp64), the
            shipped RealESRGAN_x4plus checkpoint on the reference's own test images (SR/rrdbnet_arch.py:648-667),
            HRfeature / HRfuse_residual in eval mode, the aggregation operators and the uncertainty-weighted losses.

On a clone without the staged copy every test here skips (the committed goldens still pin the same paths).
"""
import glob
import os

import numpy as np
import pytest
import torch

import synth
from conftest import ROOT, assert_close
from oracle import stage_reference

needs_ref = pytest.mark.skipif(not stage_reference.available(),
                               reason="no staged reference under baseline/_ref (oracle/stage_reference.py)")


@pytest.fixture(scope="module")
def ref():
    return stage_reference.load()


def _tsd(sd, dtype=torch.float32):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) if v.dtype.kind == "f"
            else torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def _ref_rrdbnet(ref, sd, num_block, scale=4, dtype=torch.float32):
    net = ref.arch.RRDBNet(3, 3, scale, 64, num_block, 32)
    net.load_state_dict(_tsd(sd), strict=True)
    return net.to(dtype).eval()


# ---------------------------------------------------------------------------------------------- CPU: oracle vs live
@needs_ref
def test_staged_files_are_the_manifest(ref):
    """The staged copy is byte-identical to what the manifest recorded at staging time (nothing edited since)."""
    import hashlib
    import json
    with open(os.path.join(ref.root, "MANIFEST.json")) as f:
        man = json.load(f)["sha256"]
    assert "SR/rrdbnet_arch.py" in man and "SR/HRfuse.py" in man
    for rel, sha in man.items():
        with open(os.path.join(ref.root, rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == sha, rel


@needs_ref
def test_staged_reference_does_not_shadow_the_drop_in_import_paths(ref):
    """Loading the staged reference leaves `SR.*`, `mymodels` and `aggregate_utils` resolving to this repo's drop-in
    modules (the reference is imported under private names), so product code can never pick it up by accident."""
    import importlib
    import sys
    ours = importlib.import_module("SR.rrdbnet_arch")
    assert os.path.realpath(ours.__file__).startswith(os.path.realpath(ROOT) + os.sep + "SR")
    assert os.path.realpath(ref.arch.__file__).startswith(os.path.realpath(ref.root))
    assert ours.RRDBNet is not ref.arch.RRDBNet
    assert importlib.import_module("aggregate_utils").__file__ != ref.aggregate.__file__
    pkg = importlib.import_module("bhsr")
    for name, mod in list(sys.modules.items()):
        f = getattr(mod, "__file__", None) or ""
        if name.startswith(pkg.__name__ + ".") or name.startswith("bhsr."):
            assert "baseline" + os.sep + "_ref" not in f and os.sep + "oracle" + os.sep not in f, name


@needs_ref
@pytest.mark.parametrize("scale,num_block,seed", [(4, 2, 11), (2, 1, 12), (1, 1, 13)])
def test_oracle_rrdbnet_vs_live_reference(ref, scale, num_block, seed):
    """oracle/ref_torch.py == the reference modules, bit for bit (forward_feature and forward, every scale's
    pixel-unshuffle path: SR/rrdbnet_arch.py:101-110, 227-254); oracle/ref_numpy.py within fp32 re-association."""
    from oracle import ref_numpy as R
    from oracle import ref_torch as T
    sd = synth.rrdbnet_state(num_in_ch=3, num_out_ch=3, scale=scale, num_block=num_block, seed=seed)
    net = _ref_rrdbnet(ref, sd, num_block, scale)
    hw = 16 * (4 // scale)
    x = torch.from_numpy(synth.tiles(2, 3, hw, hw, seed=seed))
    with torch.no_grad():
        fea = net.forward_feature(x)
        img = net(x)
    tsd = _tsd(sd)
    assert torch.equal(T.rrdbnet_forward_feature(x, tsd, scale=scale), fea)
    assert torch.equal(T.rrdbnet_forward(x, tsd, scale=scale), img)
    if scale == 4:
        got = R.rrdbnet_forward_feature(x.numpy(), sd)
        assert_close(got, fea.numpy(), rtol=1e-4, atol=1e-5, what="numpy oracle vs live reference")


@needs_ref
@pytest.mark.parametrize("training", [False, True])
def test_oracle_head_blocks_vs_live_reference(ref, training):
    """HRfeature / HRfuse_residual (SR/HRfuse.py:161-231): the torch oracle against the live modules, eval and train."""
    from oracle import ref_torch as T
    sd = synth.hrfeature_state(seed=21)
    m = ref.hrfuse.HRfeature(in_chans=64, mid_chans=16, out_chans=16)
    m.load_state_dict(_tsd(sd), strict=True)
    m.train(training)
    x = torch.from_numpy(synth.features(2, 64, 32, 32, seed=5))
    with torch.no_grad():
        want = m(x)
    got = T.hrfeature(x, _tsd(sd), training=training)
    assert_close(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-6, what="HRfeature oracle vs live")
    for out in (1, 7):
        sd = synth.hrfuse_residual_state(out=out, seed=22 + out)
        m = ref.hrfuse.HRfuse_residual(hr_chans=16, lr_chans=16, mid_chans=16, out_chans=out, upscale=4)
        m.load_state_dict(_tsd(sd), strict=True)
        m.train(training)
        lr = torch.from_numpy(synth.features(2, 16, 8, 8, seed=6))
        hr = torch.from_numpy(synth.features(2, 16, 32, 32, seed=7))
        with torch.no_grad():
            want = m(lr, hr)
        got = T.hrfuse_residual(lr, hr, _tsd(sd), training=training)
        assert_close(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-6, what=f"HRfuse_residual(out={out}) oracle vs live")


# ---------------------------------------------------------------------------------------------- GPU: CUDA vs live
@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda:0")


def _load(module, sd, dev):
    module.load_state_dict(_tsd(sd), strict=True)
    return module.to(dev).eval()


@needs_ref
@pytest.mark.gpu
def test_cuda_rrdbnet_23block_vs_live_reference_full_tensor(ref, dev):
    """BASELINE config 2 shape (B=64, 23 blocks, exact numerics): every output element of four tiles of the batch
    against the UNMODIFIED reference `RRDBNet.forward_feature` run live in fp64 on the host."""
    from bhsr import rrdbnet
    sd = synth.rrdbnet_state(num_block=23, seed=77)
    net = _load(rrdbnet.RRDBNet(3, 3, scale=4, num_feat=64, num_block=23, num_grow_ch=32), sd, dev)
    x = synth.tiles(64, 3, seed=4242)
    picks = [0, 21, 42, 63]
    with torch.no_grad():
        got = net.forward_feature(torch.from_numpy(x).to(dev))[picks].cpu().numpy()
    rnet = _ref_rrdbnet(ref, sd, 23, dtype=torch.float64)
    with torch.no_grad():
        want = rnet.forward_feature(torch.from_numpy(x[picks]).double()).numpy()
    assert got.shape == want.shape == (4, 64, 256, 256)
    assert_close(got, want, what="CUDA forward_feature vs live reference (23 blocks, B=64)")


@needs_ref
@pytest.mark.gpu
def test_cuda_x4plus_on_reference_test_images_vs_live_reference(ref, dev):
    """The checkpoint the reference ships on the images the reference ships, the way its own demo feeds them
    (SR/rrdbnet_arch.py:648-667: cv2.imread, /255, HWC -> CHW): SR image (`forward`) and feature map
    (`forward_feature`), full tensors, against the live reference in fp64."""
    import cv2
    from bhsr import rrdbnet
    ckpt = os.path.join(ROOT, "oracle", "_ref", "RealESRGAN_x4plus.pth")
    imgs = sorted(glob.glob(os.path.join(ref.root, "SR", "testimg", "*.jpg")))
    if not os.path.exists(ckpt) or not imgs:
        pytest.skip("RealESRGAN_x4plus.pth / SR/testimg not staged")
    state = torch.load(ckpt, map_location="cpu")["params_ema"]
    tiles = np.stack([cv2.imread(p).astype(np.float32) / 255.0 for p in imgs]).transpose(0, 3, 1, 2)
    assert tiles.shape[1:] == (3, 64, 64)
    net = rrdbnet.RRDBNet(3, 3, scale=4, num_feat=64, num_block=23, num_grow_ch=32)
    net.load_state_dict(state, strict=True)
    net = net.to(dev).eval()
    rnet = ref.arch.RRDBNet(3, 3, 4, 64, 23, 32)
    rnet.load_state_dict(state, strict=True)
    rnet = rnet.double().eval()
    xt = torch.from_numpy(np.ascontiguousarray(tiles))
    with torch.no_grad():
        fea = net.forward_feature(xt.to(dev)).cpu().numpy()
        img = net(xt.to(dev)).cpu().numpy()
        want_fea = rnet.forward_feature(xt.double()).numpy()
        want_img = rnet(xt.double()).numpy()
    assert_close(fea, want_fea, what="x4plus forward_feature on SR/testimg vs live reference")
    assert_close(img, want_img, what="x4plus forward (SR image) on SR/testimg vs live reference")


@needs_ref
@pytest.mark.gpu
def test_cuda_head_blocks_vs_live_reference(ref, dev):
    """HRfeature and HRfuse_residual (SR/HRfuse.py:161-231) in eval mode at the head's real map size
    (64x64 -> 256x256 would be 4 GB in fp64 on the host: 32x32 -> 128x128 here), full tensors vs the live modules."""
    from bhsr import hrfuse
    sd = synth.hrfeature_state(seed=31)
    m = _load(hrfuse.HRfeature(in_chans=64, mid_chans=16, out_chans=16), sd, dev)
    r = ref.hrfuse.HRfeature(in_chans=64, mid_chans=16, out_chans=16)
    r.load_state_dict(_tsd(sd), strict=True)
    r = r.double().eval()
    x = synth.features(3, 64, 128, 128, seed=8)
    with torch.no_grad():
        got = m(torch.from_numpy(x).to(dev)).cpu().numpy()
        want = r(torch.from_numpy(x).double()).numpy()
    assert_close(got, want, what="HRfeature (eval) vs live reference")
    for out in (1, 7):
        sd = synth.hrfuse_residual_state(out=out, seed=40 + out)
        m = _load(hrfuse.HRfuse_residual(hr_chans=16, lr_chans=16, mid_chans=16, out_chans=out, upscale=4), sd, dev)
        r = ref.hrfuse.HRfuse_residual(hr_chans=16, lr_chans=16, mid_chans=16, out_chans=out, upscale=4)
        r.load_state_dict(_tsd(sd), strict=True)
        r = r.double().eval()
        lr = synth.features(3, 16, 32, 32, seed=9)
        hr = synth.features(3, 16, 128, 128, seed=10)
        with torch.no_grad():
            got = m(torch.from_numpy(lr).to(dev), torch.from_numpy(hr).to(dev)).cpu().numpy()
            want = r(torch.from_numpy(lr).double(), torch.from_numpy(hr).double()).numpy()
        assert_close(got, want, what=f"HRfuse_residual(out={out}, eval) vs live reference")


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("training", [True, False])
def test_cuda_head_training_step_vs_live_reference_autograd(ref, dev, training):
    """Row K10: HRfeature -> HRfuse_residual under autograd (tcgen05 forward / dgrad / wgrad, BatchNorm forward and
    backward kernels) against the live reference modules differentiated by stock PyTorch autograd in fp64
    (SR/HRfuse.py:143-190 as driven by train.py:246-257): output, both input gradients, every parameter gradient and
    the BatchNorm running statistics."""
    from bhsr import hrfuse
    oc = 7
    sdf = synth.hrfeature_state(seed=81)
    sdr = synth.hrfuse_residual_state(out=oc, seed=82)
    hr = synth.features(2, 64, 32, 32, seed=14)
    lr = synth.features(2, 16, 8, 8, seed=15)
    g = np.random.RandomState(16).standard_normal((2, oc, 32, 32)).astype(np.float32)

    rf = ref.hrfuse.HRfeature(in_chans=64, mid_chans=16, out_chans=16)
    rf.load_state_dict(_tsd(sdf), strict=True)
    rr = ref.hrfuse.HRfuse_residual(hr_chans=16, lr_chans=16, mid_chans=16, out_chans=oc, upscale=4)
    rr.load_state_dict(_tsd(sdr), strict=True)
    rf, rr = rf.double().train(training), rr.double().train(training)
    a = torch.from_numpy(lr).double().requires_grad_(True)
    b = torch.from_numpy(hr).double().requires_grad_(True)
    y = rr(a, rf(b))
    (y * torch.from_numpy(g).double()).sum().backward()

    mf = _load(hrfuse.HRfeature(64, 16, 16), sdf, dev).train(training)
    mr = _load(hrfuse.HRfuse_residual(16, 16, 16, oc, 4), sdr, dev).train(training)
    ac = torch.from_numpy(lr).to(dev).requires_grad_(True)
    bc = torch.from_numpy(hr).to(dev).requires_grad_(True)
    yc = mr(ac, mf(bc))
    (yc * torch.from_numpy(g).to(dev)).sum().backward()

    assert_close(yc.detach().cpu().numpy(), y.detach().numpy(), what="forward")
    scale = lambda r: 1e-4 * max(1.0, float(np.abs(r).max()))
    for got, want, what in ((ac.grad, a.grad, "grad x_lr"), (bc.grad, b.grad, "grad hr_fea")):
        assert_close(got.cpu().numpy(), want.numpy(), rtol=2e-3, atol=scale(want.numpy()), what=what)
    for mod, rmod, tag in ((mf, rf, "hrfeat"), (mr, rr, "fuse")):
        rgrads = dict(rmod.named_parameters())
        for name, prm in mod.named_parameters():
            assert prm.grad is not None, name
            want = rgrads[name].grad.numpy()
            assert_close(prm.grad.cpu().numpy(), want, rtol=2e-3, atol=scale(want), what=f"grad {tag}.{name}")
        rbufs = dict(rmod.named_buffers())
        for name, buf in mod.named_buffers():
            if "running" in name:
                np.testing.assert_allclose(buf.cpu().numpy(), rbufs[name].numpy(), rtol=1e-4, atol=1e-5, err_msg=name)
            elif "num_batches_tracked" in name:
                assert int(buf) == int(rbufs[name]), name


@needs_ref
@pytest.mark.gpu
def test_cuda_aggregate_vs_live_reference(ref, dev):
    """aggregate_torch / aggregate_torch_gpu (aggregate_utils.py:29-59): 4x4 block sums over the valid-pixel count of
    256x256 height labels (82 % zeros, like the loader's), against the live reference functions (a stock conv on the
    same device)."""
    import aggregate_utils as ours
    g = torch.Generator().manual_seed(5)
    h = torch.rand(4, 1, 256, 256, generator=g) * 60
    h = torch.where(torch.rand(4, 1, 256, 256, generator=g) < 0.82, torch.zeros(()), h).to(dev)
    want = ref.aggregate.aggregate_torch_gpu(h, 0.25, device=dev)
    got = ours.aggregate_torch_gpu(h, 0.25, device=dev)
    assert got.shape == want.shape
    assert_close(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-6, atol=1e-6, what="aggregate_torch_gpu vs live reference")
    want = ref.aggregate.aggregate_torch(h.cpu()[:1], 0.25)          # the loader calls it per sample on host tensors
    got = ours.aggregate_torch(h[:1], 0.25)
    assert got.shape == want.shape
    assert_close(got.cpu().numpy(), want.numpy(), rtol=1e-6, atol=1e-6, what="aggregate_torch vs live reference")


@needs_ref
@pytest.mark.gpu
def test_cuda_weighted_losses_vs_live_reference(ref, dev):
    """Row N4: the fused weighted-MSE forward+backward kernel (dp.MSE_adapt_weight -> bhsr_weighted_mse) and
    CE_DICE_adapt_weight against the live reference classes (losses_pytorch/selfloss.py:81-90, 145-168, which create
    their log-variance parameter on "cuda"): loss value, d loss / d prediction and d loss / d log_var, at the training
    step's shapes."""
    from bhsr import dp
    g = torch.Generator().manual_seed(3)
    for log_var in (0.0, 0.7):
        pred = (torch.randn(4, 1, 256, 256, generator=g) * 10).to(dev).requires_grad_(True)
        tgt = (torch.rand(4, 1, 256, 256, generator=g) * 30).to(dev)
        wgt = (torch.rand(4, 1, 256, 256, generator=g) + 0.5).to(dev)
        a = dp.MSE_adapt_weight(log_var, device=dev)
        b = ref.selfloss.MSE_adapt_weight(log_var)
        la, lb = a(pred, tgt, wgt), b(pred, tgt, wgt)
        ga = torch.autograd.grad(la, [pred, a.log_var])
        gb = torch.autograd.grad(lb, [pred, b.log_var])
        assert torch.allclose(la, lb, rtol=1e-4, atol=0), (la.item(), lb.item())
        assert torch.allclose(ga[0], gb[0], rtol=1e-5, atol=1e-9)
        assert torch.allclose(ga[1], gb[1], rtol=1e-4, atol=1e-6)
        logits = torch.randn(4, 7, 64, 64, generator=g).to(dev).requires_grad_(True)
        cls = torch.randint(0, 7, (4, 64, 64), generator=g).to(dev)
        cw = (torch.rand(4, 64, 64, generator=g) + 0.5).to(dev)
        a = dp.CE_DICE_adapt_weight(log_var, device=dev)
        b = ref.selfloss.CE_DICE_adapt_weight(log_var)
        la, lb = a(logits, cls, cw), b(logits, cls, cw)
        ga = torch.autograd.grad(la, [logits, a.log_var])
        gb = torch.autograd.grad(lb, [logits, b.log_var])
        assert torch.allclose(la, lb, rtol=1e-4, atol=0), (la.item(), lb.item())
        assert torch.allclose(ga[0], gb[0], rtol=1e-4, atol=1e-9)
        assert torch.allclose(ga[1], gb[1], rtol=1e-4, atol=1e-6)

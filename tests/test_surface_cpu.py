"""CPU-only checks of the drop-in boundary: C-ABI exports and struct layouts, reference import
paths, state_dict surface, init parity with the reference constructors, error behaviour."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, assert_close

REF = "/root/reference"
HEADER = os.path.join(ROOT, "include", "bhsr.h")


def test_library_exports_every_declared_symbol():
    import bhsr
    from bhsr import _lib
    lib = _lib.load()
    src = open(HEADER).read()
    declared = set(re.findall(r"\b(bhsr_[a-z0-9_]+)\s*\(", src))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bhsr_version() == 100
    # size queries are host-only and must work without a GPU
    # exact: 3 chunks of 32 channels x 9 taps x (hi+lo) 64 rows x 32 ch; fast: 2 chunks of 64 x 9 x 32 rows x 64
    assert lib.bhsr_packed_conv_weight_bytes(32, 96, 9, 0) == 3 * 9 * 64 * 32 * 2
    assert lib.bhsr_packed_conv_weight_bytes(32, 96, 9, 1) == 2 * 9 * 32 * 64 * 2
    assert lib.bhsr_rrdbnet_bias_floats(23) == 23 * 3 * (4 * 32 + 64) + 4 * 64
    assert lib.bhsr_rrdbnet_workspace_bytes(1, 64, 64, 1) > 0


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    from bhsr import _lib
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "bhsr.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(BhsrConvTcDesc), offsetof(BhsrConvTcDesc, desc_mode), sizeof(BhsrRrdbNetDesc),'
                    'offsetof(BhsrRrdbNetDesc, mblocks), sizeof(BhsrHeadConvDesc), offsetof(BhsrHeadConvDesc, accumulate));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_lib.ConvTcDesc), _lib.ConvTcDesc.desc_mode.offset,
            ctypes.sizeof(_lib.RrdbNetDesc), _lib.RrdbNetDesc.mblocks.offset,
            ctypes.sizeof(_lib.HeadConvDesc), _lib.HeadConvDesc.accumulate.offset]
    assert got == want


def test_reference_import_paths_resolve():
    from SR.rrdbnet_arch import RealESRGAN, RRDBNet, pixel_unshuffle  # train.py:14
    from SR.HRfuse import HRfuse, HRfuse_x2, HRfeature, HRfuse_residual, Refine_residual, GeoNet, HRupsample  # mymodels.py:13
    from mymodels import SRRegress_Cls_feature  # train.py:16
    from aggregate_utils import aggregate_torch  # BH_loader.py:9
    import SR.RRDBNet as old
    assert callable(RealESRGAN) and callable(SRRegress_Cls_feature) and callable(aggregate_torch)
    assert len(old.RRDBNet(4, 3, 64, 1).state_dict()) == 42


def test_rrdbnet_state_dict_surface():
    from bhsr.rrdbnet import RRDBNet
    net = RRDBNet(3, 3, scale=4, num_feat=64, num_block=23, num_grow_ch=32)
    sd = net.state_dict()
    assert len(sd) == 702 and sum(v.numel() for v in sd.values()) == 16697987  # rrdbnet_arch.py:658
    assert sd["body.22.rdb3.conv5.weight"].shape == (64, 192, 3, 3)
    assert sd["body.0.rdb1.conv2.weight"].shape == (32, 96, 3, 3)
    assert sd["conv_last.weight"].shape == (3, 64, 3, 3)
    assert RRDBNet(3, 3, scale=2, num_block=1).conv_first.weight.shape == (64, 12, 3, 3)
    assert RRDBNet(3, 3, scale=1, num_block=1).conv_first.weight.shape == (64, 48, 3, 3)


def test_x4plus_checkpoint_loads_strict():
    from bhsr.rrdbnet import RRDBNet
    ckpt = os.path.join(ROOT, "oracle", "_ref", "RealESRGAN_x4plus.pth")
    if not os.path.exists(ckpt):
        pytest.skip("checkpoint not staged")
    net = RRDBNet(3, 3, scale=4, num_feat=64, num_block=23, num_grow_ch=32)
    net.load_state_dict(torch.load(ckpt, map_location="cpu")["params_ema"], strict=True)


def test_head_state_dict_surface():
    from bhsr.models import SRRegress_Cls_feature
    net = SRRegress_Cls_feature("efficientnet-b4", encoder_weights=None, in_channels=8, super_in=64,
                                super_mid=16, upscale=4, isaggre=True, chans_build=7)
    own = [(n, p) for n, p in net.named_parameters() if n.split(".")[0] in ("reg", "seg", "hrfeat", "aggre_height")]
    assert len(own) == 77 and sum(p.numel() for _, p in own) == 94137
    enc = sum(p.numel() for p in net.encoder.parameters())
    assert abs(enc / 1e6 - 17.55) < 0.01  # mymodels.py:765 "encoder 17.55 M"
    sd = net.state_dict()
    for k in ("reg.upsampler.0.weight", "reg.upsampler.2.bias", "reg.fuse.0.downsample.0.weight",
              "seg.fuse.2.bn2.num_batches_tracked", "hrfeat.0.downsample.1.running_var", "seg.conv_last.bias",
              "aggre_height.weight", "encoder._blocks.31._project_conv.weight", "decoder1.blocks.4.conv2.1.weight"):
        assert k in sd, k
    assert sd["seg.conv_last.weight"].shape == (7, 16, 3, 3)
    assert sd["reg.fuse.0.conv1.weight"].shape == (16, 32, 3, 3)
    assert "aggre_height.weight" not in SRRegress_Cls_feature("efficientnet-b4", encoder_weights=None).state_dict()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_surface_and_init_match_the_reference_constructors():
    """Same seed -> same tensors as the reference classes: key order, shapes, and the RNG-order
    dependent init sequence (kaiming*0.1 in RDBs, default init elsewhere)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    arch, old, hrf, agg = make_golden.import_reference()
    try:
        from bhsr import hrfuse, rrdbnet
        torch.manual_seed(7)
        ref = arch.RRDBNet(3, 3, scale=4, num_feat=64, num_block=2, num_grow_ch=32)
        torch.manual_seed(7)
        mine = rrdbnet.RRDBNet(3, 3, scale=4, num_feat=64, num_block=2, num_grow_ch=32)
        rs, ms = ref.state_dict(), mine.state_dict()
        assert list(rs.keys()) == list(ms.keys())
        for k in rs:
            assert torch.equal(rs[k], ms[k]), k
        torch.manual_seed(3)
        ref = old.RRDBNet(4, 3, 64, 1, 32)
        torch.manual_seed(3)
        mine = rrdbnet.OldRRDBNet(4, 3, 64, 1, 32)
        assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
        for cls, args in (("HRfeature", (64, 16, 16)), ("HRfuse_residual", (16, 16, 16, 7, 4)), ("Upsampler", ()),
                          ("HRupsample", (4, 3, 4)), ("GeoNet", (4, 16)), ("Refine_residual", (16, 16, 16, 3)),
                          ("HRfuse", ()), ("HRfuse_x2", ())):
            torch.manual_seed(11)
            r = getattr(hrf, cls)(*args)
            torch.manual_seed(11)
            m = getattr(hrfuse, cls)(*args)
            rs, ms = r.state_dict(), m.state_dict()
            assert list(rs.keys()) == list(ms.keys()), cls
            for k in rs:
                assert torch.equal(rs[k], ms[k]), (cls, k)
        # HRfuse.py:233-241 smoke: HRupsample(lr_chans=4, out_chans=3, upscale=4) has 1,295 parameters
        assert sum(p.numel() for p in hrfuse.HRupsample(4, 3, 4).parameters()) == \
            sum(p.numel() for p in hrf.HRupsample(4, 3, 4).parameters()) == 1295
    finally:
        for k in [k for k in sys.modules if k == "SR" or k.startswith("SR.") or k in ("aggregate_utils",)]:
            del sys.modules[k]
        sys.path[:] = [p for p in sys.path if p != REF]


def test_no_cpu_fallback_and_autograd_guard():
    from bhsr._lib import BhsrError
    from bhsr.hrfuse import HRfeature
    from bhsr.rrdbnet import RRDBNet
    net = RRDBNet(3, 3, num_block=1)
    with pytest.raises(BhsrError, match="CUDA tensor"):
        net.forward_feature(torch.zeros(1, 3, 64, 64))
    with pytest.raises(BhsrError, match="CUDA tensor"):
        net(torch.zeros(1, 3, 64, 64))
    with pytest.raises(BhsrError, match="CUDA tensor"):
        HRfeature(64, 16, 16)(torch.zeros(1, 64, 8, 8))
    with pytest.raises(AssertionError):
        from bhsr.rrdbnet import pixel_unshuffle
        pixel_unshuffle(torch.zeros(1, 1, 5, 4), 2)  # rrdbnet_arch.py:106
    with pytest.raises(ValueError):
        from bhsr.hrfuse import BasicBlock
        BasicBlock(4, 4, groups=2)  # HRfuse.py:125-126
    with pytest.raises(NotImplementedError):
        from bhsr.hrfuse import Upsampler
        Upsampler(scale=5)  # HRfuse.py:41-42


def test_pixel_unshuffle_bit_exact(golden):
    from bhsr.rrdbnet import pixel_unshuffle
    x = torch.from_numpy(golden["pixel_unshuffle_in"])
    assert np.array_equal(pixel_unshuffle(x, 2).numpy(), golden["pixel_unshuffle_s2"])
    assert np.array_equal(pixel_unshuffle(x, 4).numpy(), golden["pixel_unshuffle_s4"])


def test_aggregate_host_path_vs_golden(golden):
    from bhsr import aggregate
    x = torch.from_numpy(golden["aggregate_in"])
    y = aggregate.aggregate_torch(x, 0.25)
    assert y.shape == (64, 64)
    assert_close(y.numpy(), golden["aggregate_torch"], rtol=1e-6, atol=1e-6, what="aggregate_torch")
    assert_close(aggregate.aggregate(golden["aggregate_in"][0, 0], 0.25), golden["aggregate_loop"], rtol=1e-9, atol=1e-9)
    y2 = aggregate._block_aggregate(x, 4, 1.0, True)
    assert_close(y2.numpy(), golden["aggregate_torch_gpu"], rtol=1e-5, atol=1e-3, what="aggregate_torch_gpu")


def test_realesrgan_shell_has_no_side_effects():
    from bhsr.rrdbnet import RealESRGAN
    m = RealESRGAN(device="cpu", num_block=1)   # the reference needs CUDA + a VGG19 download here
    assert len(m.net_g.state_dict()) == 42 and m.net_g.training
    assert m.net_d is None and m.cri_perceptual is None          # stock PyTorch pieces: pluggable, not built
    with pytest.raises(AttributeError, match="fine-tuning"):
        m.optimizer_g                                            # only with is_train=True
    t = RealESRGAN(device="cpu", num_block=1, is_train=True, ema_decay=0.9)   # rrdbnet_arch.py:459-505
    assert isinstance(t.optimizer_g, torch.optim.Adam) and t.optimizer_g.defaults["betas"] == (0.9, 0.99)
    assert all(not p.requires_grad for p in t.net_g_ema.parameters())
    for (k, a), (_, b) in zip(t.net_g.state_dict().items(), t.net_g_ema.state_dict().items()):
        assert torch.equal(a, b), k                              # model_ema(0) copies the generator (:476)


def test_packed_weight_caches_are_invalidated():
    """caches keyed on (data_ptr, _version) miss writes through `.data`; modules must
    drop them on load_state_dict / .to() / explicit invalidate_cache(), and the strict mode must see the
    content change."""
    import torch
    from bhsr import _lib, hrfuse, rrdbnet
    net = rrdbnet.RRDBNet(3, 3, num_block=1)
    blk = hrfuse.BasicBlock(32, 16)
    fuse = hrfuse.HRfuse_residual(16, 16, 16, 1, 4)
    for m in (net, blk, fuse):
        m.__dict__["_tc_cache"] = ("k", "v")
        m.__dict__["_tc_own"] = ("k", "v")
    net._packed = ("key", None, None)
    net.load_state_dict(net.state_dict())
    assert net._packed is None and "_tc_cache" not in net.__dict__
    blk.load_state_dict(blk.state_dict())
    assert "_tc_cache" not in blk.__dict__
    fuse.__dict__["_tc_own"] = ("k", "v")
    fuse.double()                       # _apply
    assert "_tc_own" not in fuse.__dict__ and "_tc_cache" not in fuse.fuse[0].__dict__
    blk.__dict__["_tc_cache"] = ("k", "v")
    parent = torch.nn.Sequential(blk)
    _lib.invalidate_cache(parent)       # explicit, from any ancestor
    assert "_tc_cache" not in blk.__dict__
    # version keys: in-place ops on the parameter change the key, `.data` writes do not (documented) ...
    w = blk.conv1.weight
    k0 = _lib.tensor_key([w])
    with torch.no_grad():
        w.mul_(1.0)
    assert _lib.tensor_key([w]) != k0
    k1 = _lib.tensor_key([w])
    w.data.mul_(2.0)
    assert _lib.tensor_key([w]) == k1
    # ... unless the strict (fingerprint) mode is on
    old = _lib.STRICT_CACHE
    try:
        _lib.STRICT_CACHE = True
        k2 = _lib.tensor_key([w])
        w.data.mul_(2.0)
        assert _lib.tensor_key([w]) != k2
    finally:
        _lib.STRICT_CACHE = old
    # kernels that write running statistics through raw pointers bump the version
    rm = blk.bn1.running_mean
    v = rm._version
    _lib.bump_version(rm)
    assert rm._version > v


def test_basic_block_rejects_unsupported_batchnorm_modes():
    import pytest
    import torch
    from bhsr import hrfuse
    blk = hrfuse.BasicBlock(16, 16)
    blk.train()
    blk.bn1.eval()                      # "freeze BN only": not supported by the fused kernels -> loud error
    with pytest.raises(NotImplementedError):
        blk(torch.zeros(1, 16, 4, 4))
    blk2 = hrfuse.BasicBlock(16, 16)
    blk2.bn2.momentum = None
    with pytest.raises(NotImplementedError):
        blk2(torch.zeros(1, 16, 4, 4))

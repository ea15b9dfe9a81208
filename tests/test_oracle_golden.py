"""Pin the CPU oracles (oracle/ref_numpy.py, oracle/ref_torch.py) against the golden vectors the
REFERENCE modules produced (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import synth
from conftest import assert_close
from oracle import ref_numpy as R
from oracle import ref_torch as T

TIGHT = dict(rtol=2e-5, atol=2e-5)


def tdict(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def test_rrdbnet_2block(golden):
    sd = synth.rrdbnet_state(num_block=2, seed=11)
    x = synth.tiles(2, 3, seed=1337)
    fea = R.rrdbnet_forward_feature(x, sd)
    assert fea.shape == (2, 64, 256, 256)
    assert_close(synth.subsample(fea), golden["rrdb2_feature_sub"], what="numpy forward_feature", **TIGHT)
    assert_close(fea[0, :8, :20, :20], golden["rrdb2_feature_corner"], what="corner", **TIGHT)
    assert_close(fea[1, 56:, 236:, 236:], golden["rrdb2_feature_edge"], what="edge", **TIGHT)
    np.testing.assert_allclose(synth.stats(fea), golden["rrdb2_feature_stats"], rtol=1e-5)
    img = R.rrdbnet_forward(x, sd)
    assert_close(synth.subsample(img, 1, 4), golden["rrdb2_forward_sub"], what="numpy forward", **TIGHT)
    feat = T.rrdbnet_forward_feature(torch.from_numpy(x), tdict(sd)).numpy()
    assert_close(synth.subsample(feat), golden["rrdb2_feature_sub"], what="torch forward_feature", **TIGHT)
    imgt = T.rrdbnet_forward(torch.from_numpy(x), tdict(sd)).numpy()
    np.testing.assert_allclose(synth.stats(imgt), golden["rrdb2_forward_stats"], rtol=1e-5)


def test_rrdbnet_23block(golden):
    sd = synth.rrdbnet_state(num_block=23, seed=23)
    assert len(sd) == 702 and sum(v.size for v in sd.values()) == 16697987  # rrdbnet_arch.py:658
    x = synth.tiles(1, 3, seed=4242)
    # deep stack, outputs up to ~1e2: compare in the north-star tolerance, fp64-accumulating oracle
    fea = R.rrdbnet_forward_feature(x, sd, acc_dtype=np.float64)
    assert_close(synth.subsample(fea), golden["rrdb23_feature_sub"], rtol=1e-4, atol=1e-4, what="numpy 23 blocks")
    feat = T.rrdbnet_forward_feature(torch.from_numpy(x), tdict(sd)).numpy()
    assert_close(synth.subsample(feat), golden["rrdb23_feature_sub"], rtol=1e-4, atol=1e-4, what="torch 23 blocks")


def test_old_rrdbnet_and_scales(golden):
    sd = synth.rrdbnet_state(num_in_ch=4, num_block=1, seed=5)
    x = synth.tiles(2, 4, seed=99)
    y = R.old_rrdbnet_forward(x, synth.to_old_rrdbnet_keys(sd))
    assert_close(synth.subsample(y, 1, 4), golden["old_rrdb1_forward_sub"], what="old class", **TIGHT)
    for sc, hw in ((2, 128), (1, 256)):
        sd = synth.rrdbnet_state(num_in_ch=3, scale=sc, num_block=1, seed=50 + sc)
        x = synth.tiles(1, 3, hw, hw, seed=60 + sc)
        y = R.rrdbnet_forward(x, sd, scale=sc)
        assert_close(synth.subsample(y, 1, 4), golden[f"rrdb1_scale{sc}_forward_sub"], what=f"scale {sc}", **TIGHT)
        yt = T.rrdbnet_forward(torch.from_numpy(x), tdict(sd), scale=sc).numpy()
        assert_close(synth.subsample(yt, 1, 4), golden[f"rrdb1_scale{sc}_forward_sub"], what=f"torch scale {sc}", **TIGHT)


def test_index_paths_bit_exact(golden):
    assert np.array_equal(R.pixel_unshuffle(golden["pixel_unshuffle_in"], 2), golden["pixel_unshuffle_s2"])
    assert np.array_equal(R.pixel_unshuffle(golden["pixel_unshuffle_in"], 4), golden["pixel_unshuffle_s4"])
    assert np.array_equal(R.pixel_shuffle(golden["pixel_shuffle_in"], 2), golden["pixel_shuffle_r2"])
    assert np.array_equal(R.nearest_up2(golden["nearest_in"]), golden["nearest_x2"])
    with pytest.raises(AssertionError):  # rrdbnet_arch.py:106
        R.pixel_unshuffle(np.zeros((1, 1, 5, 4), np.float32), 2)


@pytest.mark.parametrize("training", [False, True])
def test_head_pieces(golden, training):
    tag = "train" if training else "eval"
    hr = synth.features(2, 64, 64, 64, seed=21)
    lr = synth.features(2, 16, 16, 16, seed=22)
    hr16 = synth.features(2, 16, 64, 64, seed=23)
    sd = synth.hrfeature_state(seed=31)
    new_stats = {}
    y = R.hrfeature(hr, sd, training=training, new_stats=new_stats)
    assert_close(y, golden[f"hrfeature_{tag}"], rtol=1e-4, atol=1e-4, what=f"HRfeature {tag}")
    td = tdict(sd)
    yt = T.hrfeature(torch.from_numpy(hr), td, training=training).numpy()
    assert_close(yt, golden[f"hrfeature_{tag}"], rtol=1e-4, atol=1e-4, what=f"torch HRfeature {tag}")
    if training:  # running-stat updates (momentum 0.1, unbiased variance)
        for k, v in new_stats.items():
            ref = golden["hrfeature_train_buf." + k]
            np.testing.assert_allclose(v, ref, rtol=1e-4, atol=1e-5, err_msg=k)
            if "num_batches" not in k:
                np.testing.assert_allclose(td[k].numpy(), ref, rtol=1e-4, atol=1e-5, err_msg=k)
    for oc in (1, 7):
        sd = synth.hrfuse_residual_state(out=oc, seed=40 + oc)
        y = R.hrfuse_residual(lr, hr16, sd, training=training)
        assert y.shape == (2, oc, 64, 64)
        assert_close(y, golden[f"hrfuse_out{oc}_{tag}"], rtol=1e-4, atol=1e-4, what=f"HRfuse_residual {oc} {tag}")
        yt = T.hrfuse_residual(torch.from_numpy(lr), torch.from_numpy(hr16), tdict(sd), training=training).numpy()
        assert_close(yt, golden[f"hrfuse_out{oc}_{tag}"], rtol=1e-4, atol=1e-4, what=f"torch HRfuse {oc} {tag}")


def test_upsampler(golden):
    sd = {}
    synth.upsampler_state(np.random.RandomState(77), sd, "u", 16, 4)
    lr = synth.features(2, 16, 16, 16, seed=22)
    assert_close(R.upsampler(lr, sd, "u", 4), golden["upsampler"], what="Upsampler", **TIGHT)


def test_aggregate(golden):
    x = golden["aggregate_in"]
    assert_close(R.aggregate_torch(x, 0.25), golden["aggregate_torch"], rtol=1e-6, atol=1e-6, what="aggregate_torch")
    assert R.aggregate_torch(x, 0.25).shape == (64, 64)
    assert_close(R.aggregate_torch_gpu(x, 0.25), golden["aggregate_torch_gpu"], rtol=1e-5, atol=1e-3, what="aggregate_torch_gpu")
    assert_close(R.aggregate(x[0, 0], 0.25), golden["aggregate_loop"], rtol=1e-9, atol=1e-9, what="aggregate loop")


def test_hierweight_known_answer(golden):
    """BH_loader.py:1116-1124: the reference prints these weights for the globe statistics."""
    w = R.hierweight(golden["bh_stats_globe"], (0, 3, 12, 21, 30, 60, 90, 255))
    np.testing.assert_allclose(w, golden["hierweight_kat"], rtol=0, atol=5e-9)


def build_srregress(isaggre, device="cpu"):
    """The drop-in SRRegress_Cls_feature exactly as make_golden.py built the reference class: seeded
    construction, synth.head_state for the reference-owned parameters, perturbed stand-in smp part."""
    from bhsr.models import SRRegress_Cls_feature
    torch.manual_seed(synth.SRREGRESS_SEED)
    net = SRRegress_Cls_feature("efficientnet-b4", encoder_weights=None, in_channels=8, super_in=64,
                                super_mid=16, upscale=4, isaggre=isaggre, chans_build=7)
    sd = synth.head_state(64, 16, 7, isaggre, seed=100)
    missing, unexpected = net.load_state_dict(tdict(sd), strict=False)
    assert not unexpected and all(k.split(".")[0] in ("encoder", "decoder1", "decoder2") for k in missing)
    synth.perturb_smp_state(net)
    return net.to(device).eval(), sd


def check_srregress_outputs(golden, tag, height, build, aggre=None, **tol):
    """Compare full outputs with every a16 golden array of configuration `tag`."""
    for name, r in (("height", height), ("build", build)):
        assert_close(synth.subsample(r, 1, 4), golden[f"srregress_{tag}_{name}_sub"], what=f"{tag} {name}", **tol)
        assert_close(r[0, :, :24, :24], golden[f"srregress_{tag}_{name}_corner"], what=f"{tag} {name} corner", **tol)
        assert_close(r[1, :, 232:, 232:], golden[f"srregress_{tag}_{name}_edge"], what=f"{tag} {name} edge", **tol)
        np.testing.assert_allclose(synth.stats(r)[1:3], golden[f"srregress_{tag}_{name}_stats"][1:3], rtol=1e-3)
    if aggre is not None:
        assert_close(aggre, golden[f"srregress_{tag}_height_aggre"], what=f"{tag} height_aggre", **tol)


@pytest.mark.parametrize("isaggre", [True, False])
def test_srregress_wiring_vs_reference_golden(golden, isaggre):
    """a16: the oracle's restatement of SRRegress_Cls_feature.forward (mymodels.py:270-293) against the
    outputs of the REFERENCE class (exec'd source slice, tests/golden/make_golden.py).  The decoder
    features come from the stand-in smp modules run on CPU — the same modules the generator bound into
    the reference class; their rebuilt tensors are proven identical by the stored checksum."""
    tag = "aggre" if isaggre else "noaggre"
    net, sd = build_srregress(isaggre)
    np.testing.assert_allclose(synth.smp_checksum(net), golden[f"srregress_{tag}_smp_checksum"], rtol=1e-12)
    x = torch.from_numpy(synth.tiles(2, 8, seed=3))
    sf = torch.from_numpy(synth.features(2, 64, 256, 256, seed=4))
    with torch.no_grad():
        enc = net.encoder(x)
        hfea, bfea = net.decoder1(*enc), net.decoder2(*enc)
        out = T.srregress_head(hfea, bfea, sf, tdict(sd), isaggre)
    check_srregress_outputs(golden, tag, out[0].numpy(), out[1].numpy(), out[2].numpy() if isaggre else None,
                            rtol=1e-4, atol=1e-4)
    if isaggre:   # forward_unsup = height squeezed, forward_nobuild = (height, height_aggre)  (:295-337)
        assert_close(out[0].numpy()[:, 0, ::4, ::4], golden["srregress_aggre_unsup_sub"], rtol=1e-4, atol=1e-4)
        assert_close(synth.subsample(out[0].numpy(), 1, 4), golden["srregress_aggre_nobuild_height_sub"], rtol=1e-4, atol=1e-4)
        assert_close(out[2].numpy(), golden["srregress_aggre_nobuild_height_aggre"], rtol=1e-4, atol=1e-4)
        # numpy oracle on the same decoder features
        ref = R.srregress_head(hfea.numpy(), bfea.numpy(), sf.numpy(), sd, True)
        check_srregress_outputs(golden, tag, ref[0], ref[1], ref[2], rtol=1e-4, atol=1e-4)
    else:
        assert_close(synth.subsample(out[0].numpy(), 1, 4), golden["srregress_noaggre_nobuild_height_sub"], rtol=1e-4, atol=1e-4)

"""Generate the golden vectors under tests/golden/ by running the REFERENCE modules on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The reference is imported from where it lies (nothing is copied): SR/rrdbnet_arch.py,
SR/RRDBNet.py, SR/HRfuse.py and aggregate_utils.py import after stubbing their unused
matplotlib / rasterio imports; mymodels.py does not parse (IndentationError at line 467), so
its hot class is exec'd from the source slice lines 7-14 + 231-337 with the smp names bound to
this repo's stand-in encoder/decoder (smp is a third-party dependency absent from the reference
tree) — exec_reference_srregress() / srregress_goldens() below.  Inputs and parameters come from tests/golden/synth.py.

Also stages the one real checkpoint the reference ships (SR/pretrained/RealESRGAN_x4plus.pth)
into oracle/_ref/ (git-ignored; travels to the GPU box) for the realistic-weights parity test.
"""
import os
import shutil
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("BHSR_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import synth  # noqa: E402


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "rasterio"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    sys.path.insert(0, REF)
    # make sure `SR` resolves to the reference package, not to this repo's drop-in of the same name
    for k in [k for k in sys.modules if k == "SR" or k.startswith("SR.") or k in ("aggregate_utils", "mymodels")]:
        del sys.modules[k]
    import importlib.util

    def load(modname, relpath, package_dir=None):
        spec = importlib.util.spec_from_file_location(
            modname, os.path.join(REF, relpath),
            submodule_search_locations=[package_dir] if package_dir else None)
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m

    sr = types.ModuleType("SR")
    sr.__path__ = [os.path.join(REF, "SR")]
    sys.modules["SR"] = sr
    load("SR.srloss", "SR/srloss.py")
    arch = load("SR.rrdbnet_arch", "SR/rrdbnet_arch.py")
    old = load("SR.RRDBNet", "SR/RRDBNet.py")
    hrf = load("SR.HRfuse", "SR/HRfuse.py")
    agg = load("aggregate_utils", "aggregate_utils.py")
    load("SR.edsr", "SR/edsr.py")   # imported by mymodels.py:10
    return arch, old, hrf, agg


def exec_reference_srregress():
    """The reference `SRRegress_Cls_feature` (mymodels.py:231-337).  mymodels.py cannot be imported
    (IndentationError at :467), so the class is exec'd from the source slice lines 7-14 (imports) +
    231-337 (the class), SURVEY.md §8(c), with `segmentation_models_pytorch{,.encoders,.decoders.unet}`
    bound to this repo's stand-in (bhsr.smp_compat: the third-party package is absent from the
    reference tree).  Everything else the class touches — SR.HRfuse.{HRfeature, HRfuse_residual},
    nn.Conv2d — is the reference's own code imported by import_reference()."""
    import bhsr  # noqa: F401  (package alias)
    from bhsr import smp_compat
    smp = types.ModuleType("segmentation_models_pytorch")
    enc = types.ModuleType("segmentation_models_pytorch.encoders")
    dec = types.ModuleType("segmentation_models_pytorch.decoders")
    unet = types.ModuleType("segmentation_models_pytorch.decoders.unet")
    enc.get_encoder = smp_compat.get_encoder
    unet.UnetDecoder = smp_compat.UnetDecoder
    smp.encoders, smp.decoders, dec.unet = enc, dec, unet
    for m in (smp, enc, dec, unet):
        sys.modules[m.__name__] = m
    with open(os.path.join(REF, "mymodels.py")) as f:
        lines = f.readlines()
    src = "".join(lines[6:14]) + "\n" + "".join(lines[230:337])
    ns = {"__name__": "mymodels_slice"}
    exec(compile(src, os.path.join(REF, "mymodels.py") + "[7-14,231-337]", "exec"), ns)
    return ns["SRRegress_Cls_feature"]


def srregress_goldens(out):
    """a16: forward / forward_unsup / forward_nobuild of the reference class, isaggre True and False,
    eval mode, on CPU.  The reference-owned parameters come from synth.head_state; the stand-in
    encoder / decoders are built under torch.manual_seed(synth.SRREGRESS_SEED) and perturbed by
    synth.perturb_smp_state (a checksum of them is stored so the tests can prove they rebuilt the
    same tensors)."""
    cls = exec_reference_srregress()
    x = synth.tiles(2, 8, seed=3)
    sf = synth.features(2, 64, 256, 256, seed=4)
    for isaggre in (True, False):
        torch.manual_seed(synth.SRREGRESS_SEED)
        net = cls("efficientnet-b4", encoder_weights=None, in_channels=8, super_in=64, super_mid=16,
                  upscale=4, isaggre=isaggre, chans_build=7)
        sd = synth.head_state(64, 16, 7, isaggre, seed=100)
        missing, unexpected = net.load_state_dict({k: t(v) for k, v in sd.items()}, strict=False)
        assert not unexpected and all(k.split(".")[0] in ("encoder", "decoder1", "decoder2") for k in missing)
        synth.perturb_smp_state(net)
        net.eval()
        tag = "aggre" if isaggre else "noaggre"
        out[f"srregress_{tag}_smp_checksum"] = synth.smp_checksum(net)
        with torch.no_grad():
            res = net(t(x), t(sf))
            names = ("height", "build", "height_aggre")[:len(res)]
            for n, r in zip(names, res):
                r = r.numpy()
                out[f"srregress_{tag}_{n}_stats"] = synth.stats(r)
                if n == "height_aggre":
                    out[f"srregress_{tag}_{n}"] = r
                else:
                    out[f"srregress_{tag}_{n}_sub"] = synth.subsample(r, 1, 4)
                    out[f"srregress_{tag}_{n}_corner"] = np.ascontiguousarray(r[0, :, :24, :24])
                    out[f"srregress_{tag}_{n}_edge"] = np.ascontiguousarray(r[1, :, 232:, 232:])
            if isaggre:
                u = net.forward_unsup(t(x), t(sf)).numpy()
                assert u.shape == (2, 256, 256)
                out["srregress_aggre_unsup_sub"] = np.ascontiguousarray(u[:, ::4, ::4])
                nb = net.forward_nobuild(t(x), t(sf))
                assert len(nb) == 2
                out["srregress_aggre_nobuild_height_sub"] = synth.subsample(nb[0].numpy(), 1, 4)
                out["srregress_aggre_nobuild_height_aggre"] = nb[1].numpy()
            else:
                nb = net.forward_nobuild(t(x), t(sf))
                assert isinstance(nb, torch.Tensor)
                out["srregress_noaggre_nobuild_height_sub"] = synth.subsample(nb.numpy(), 1, 4)
        print(f"srregress {tag}: height range", float(res[0].min()), float(res[0].max()),
              "build range", float(res[1].min()), float(res[1].max()))


def hrfuse_ablation_goldens(out, hrf):
    """HRfuse / HRfuse_x2 (SR/HRfuse.py:47-89), eval and train mode."""
    lr = synth.features(2, 16, 16, 16, seed=22)
    hr_lr = synth.features(2, 16, 16, 16, seed=24)      # HRfuse concatenates at the LOW resolution
    hr16 = synth.features(2, 16, 64, 64, seed=23)
    for training in (False, True):
        tag = "train" if training else "eval"
        with torch.no_grad():
            m = load_sd(hrf.HRfuse(16, 16, 16, 3, 4), synth.hrfuse_plain_state(seed=81))
            m.train(training)
            out[f"hrfuse_plain_{tag}"] = m(t(lr), t(hr_lr)).numpy()
            m = load_sd(hrf.HRfuse_x2(16, 16, 16, 3, 4), synth.hrfuse_plain_state(seed=82))
            m.train(training)
            out[f"hrfuse_x2_{tag}"] = m(t(lr), t(hr16)).numpy()


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def load_sd(module, sd):
    module.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
    return module


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    arch, old, hrf, agg = import_reference()
    out = {}
    if "--only-hrfuse-ablation" in sys.argv:
        path = os.path.join(HERE, "reference_vectors.npz")
        with np.load(path) as z:
            out = {k: z[k] for k in z.files if not k.startswith(("hrfuse_plain_", "hrfuse_x2_"))}
        hrfuse_ablation_goldens(out, hrf)
        np.savez_compressed(path, **out)
        print("updated reference_vectors.npz", os.path.getsize(path) / 1e6, "MB;", len(out), "arrays")
        return
    if "--only-srregress" in sys.argv:   # add / refresh the a16 vectors, keep every other array as it is
        path = os.path.join(HERE, "reference_vectors.npz")
        with np.load(path) as z:
            out = {k: z[k] for k in z.files if not k.startswith("srregress_")}
        srregress_goldens(out)
        np.savez_compressed(path, **out)
        print("updated reference_vectors.npz", os.path.getsize(path) / 1e6, "MB;", len(out), "arrays")
        return

    # ---------------------------------------------------------------- RRDBNet, 2 blocks
    with torch.no_grad():
        sd = synth.rrdbnet_state(num_block=2, seed=11)
        net = load_sd(arch.RRDBNet(3, 3, scale=4, num_feat=64, num_block=2, num_grow_ch=32), sd).eval()
        x = synth.tiles(2, 3, seed=1337)
        fea = net.forward_feature(t(x)).numpy()
        img = net.forward(t(x)).numpy()
        out["rrdb2_feature_sub"] = synth.subsample(fea)
        out["rrdb2_feature_stats"] = synth.stats(fea)
        out["rrdb2_forward_sub"] = synth.subsample(img, 1, 4)
        out["rrdb2_forward_stats"] = synth.stats(img)
        # one full-resolution tile corner so index-path errors cannot hide in the subsampling
        out["rrdb2_feature_corner"] = np.ascontiguousarray(fea[0, :8, :20, :20])
        out["rrdb2_feature_edge"] = np.ascontiguousarray(fea[1, 56:, 236:, 236:])
        print("rrdb2 feature range", fea.min(), fea.max())

        # ------------------------------------------------------------ RRDBNet, 23 blocks
        sd = synth.rrdbnet_state(num_block=23, seed=23)
        net = load_sd(arch.RRDBNet(3, 3, scale=4, num_feat=64, num_block=23, num_grow_ch=32), sd).eval()
        nparams = sum(p.numel() for p in net.parameters())
        assert nparams == 16697987, nparams  # rrdbnet_arch.py:658 "generator 16.70M"
        x = synth.tiles(1, 3, seed=4242)
        fea = net.forward_feature(t(x)).numpy()
        out["rrdb23_feature_sub"] = synth.subsample(fea)
        out["rrdb23_feature_stats"] = synth.stats(fea)
        print("rrdb23 feature range", fea.min(), fea.max())

        # ------------------------------------------------------------ real checkpoint
        ckpt = os.path.join(REF, "SR", "pretrained", "RealESRGAN_x4plus.pth")
        if os.path.exists(ckpt):
            dst = os.path.join(ROOT, "oracle", "_ref")
            os.makedirs(dst, exist_ok=True)
            if not os.path.exists(os.path.join(dst, "RealESRGAN_x4plus.pth")):
                shutil.copy(ckpt, os.path.join(dst, "RealESRGAN_x4plus.pth"))
            w = torch.load(ckpt, map_location="cpu")["params_ema"]
            net.load_state_dict(w, strict=True)
            import cv2
            imgs = []
            tdir = os.path.join(REF, "SR", "testimg")
            for f in sorted(os.listdir(tdir))[:2]:
                im = cv2.cvtColor(cv2.imread(os.path.join(tdir, f)), cv2.COLOR_BGR2RGB)
                imgs.append(im)
            u8 = np.stack(imgs)  # [2,64,64,3] uint8, rrdbnet_arch.py:660-663
            xin = torch.from_numpy(u8).float().permute(0, 3, 1, 2) / 255.0
            fea = net.forward_feature(xin).numpy()
            out["x4plus_input_u8"] = u8
            out["x4plus_feature_sub"] = synth.subsample(fea)
            out["x4plus_feature_stats"] = synth.stats(fea)
            print("x4plus feature range", fea.min(), fea.max())

        # ------------------------------------------------------------ old-style class, scale variants
        sd = synth.rrdbnet_state(num_in_ch=4, num_block=1, seed=5)
        net = load_sd(old.RRDBNet(in_nc=4, out_nc=3, nf=64, nb=1, gc=32), synth.to_old_rrdbnet_keys(sd)).eval()
        x = synth.tiles(2, 4, seed=99)  # SR/RRDBNet.py:82 smoke shape
        out["old_rrdb1_forward_sub"] = synth.subsample(net(t(x)).numpy(), 1, 4)
        for sc, hw in ((2, 128), (1, 256)):
            sd = synth.rrdbnet_state(num_in_ch=3, scale=sc, num_block=1, seed=50 + sc)
            net = load_sd(arch.RRDBNet(3, 3, scale=sc, num_feat=64, num_block=1, num_grow_ch=32), sd).eval()
            x = synth.tiles(1, 3, hw, hw, seed=60 + sc)
            out[f"rrdb1_scale{sc}_forward_sub"] = synth.subsample(net(t(x)).numpy(), 1, 4)
        xs = synth.features(2, 3, 8, 12, seed=3)
        out["pixel_unshuffle_in"] = xs
        out["pixel_unshuffle_s2"] = arch.pixel_unshuffle(t(xs), 2).numpy()
        out["pixel_unshuffle_s4"] = arch.pixel_unshuffle(t(xs), 4).numpy()

    # ---------------------------------------------------------------- HR fusion head pieces
    hr = synth.features(2, 64, 64, 64, seed=21)
    lr = synth.features(2, 16, 16, 16, seed=22)
    hr16 = synth.features(2, 16, 64, 64, seed=23)
    for training in (False, True):
        tag = "train" if training else "eval"
        with torch.no_grad():
            m = load_sd(hrf.HRfeature(64, 16, 16), synth.hrfeature_state(seed=31))
            m.train(training)
            out[f"hrfeature_{tag}"] = m(t(hr)).numpy()
            if training:
                for k, v in m.state_dict().items():
                    if "running" in k or "num_batches" in k:
                        out[f"hrfeature_train_buf.{k}"] = v.numpy()
            for oc in (1, 7):
                m = load_sd(hrf.HRfuse_residual(16, 16, 16, oc, 4), synth.hrfuse_residual_state(out=oc, seed=40 + oc))
                m.train(training)
                out[f"hrfuse_out{oc}_{tag}"] = m(t(lr), t(hr16)).numpy()
    with torch.no_grad():
        m = hrf.Upsampler(scale=4, n_feats=16)
        sd = {}
        synth.upsampler_state(np.random.RandomState(77), sd, "u", 16, 4)
        load_sd(m, {k[2:]: v for k, v in sd.items()})
        out["upsampler"] = m(t(lr)).numpy()
        ps_in = synth.features(1, 8, 3, 5, seed=9)
        out["pixel_shuffle_in"] = ps_in
        out["pixel_shuffle_r2"] = torch.nn.PixelShuffle(2)(t(ps_in)).numpy()
        up_in = synth.features(1, 2, 3, 4, seed=10)
        out["nearest_in"] = up_in
        out["nearest_x2"] = torch.nn.functional.interpolate(t(up_in), scale_factor=2, mode="nearest").numpy()

    # ---------------------------------------------------------------- SRRegress_Cls_feature (a16)
    srregress_goldens(out)
    hrfuse_ablation_goldens(out, hrf)

    # ---------------------------------------------------------------- aggregation
    h256 = (np.random.RandomState(8).rand(1, 1, 256, 256) * 60).astype(np.float32)
    h256[h256 < 45] = 0  # mostly-zero height label, like the dataset
    out["aggregate_in"] = h256
    out["aggregate_torch"] = agg.aggregate_torch(t(h256), 0.25).numpy()
    out["aggregate_torch_gpu"] = agg.aggregate_torch_gpu(t(h256), 0.25, device="cpu").detach().numpy()
    out["aggregate_loop"] = agg.aggregate(h256[0, 0], 0.25)

    # ---------------------------------------------------------------- loss-weight KAT (BH_loader.py:1116-1124)
    stats = np.loadtxt(os.path.join(REF, "datasetglobe", "bh_stats_globe.txt"))
    out["bh_stats_globe"] = stats
    out["hierweight_kat"] = np.array([0.08743518, 0.26821995, 0.32067124, 0.73515255, 0.98135007,
                                      1.60267172, 3.0044993])

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    sz = os.path.getsize(os.path.join(HERE, "reference_vectors.npz"))
    print("wrote reference_vectors.npz", sz / 1e6, "MB;", len(out), "arrays")


if __name__ == "__main__":
    main()

"""Development probe for the tcgen05 conv kernel (run on a B200 through gpurun).

Usage: python tools/probe_conv_tc.py <case> [desc_mode]
Each case runs in its own process (a device trap poisons the CUDA context), prints one JSON
line with the max error against a float64 torch conv of the same (decoded) operands.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import bhsr  # noqa: F401
from bhsr import ops
from bhsr._lib import NUMERICS_EXACT, NUMERICS_FAST


def split(v):
    hi = v.to(torch.float16)
    lo = ((v - hi.float()) * 2048.0).to(torch.float16)
    return hi, lo


def to_planes(x, ctot, choff=0):
    """fp32 NCHW -> (hi, lo) NHWC planes with ctot channels (others random finite)."""
    nb, c, h, w = x.shape
    full = torch.randn(nb, h, w, ctot, device=x.device) * 0.5
    full[..., choff:choff + c] = x.permute(0, 2, 3, 1)
    return split(full)


def decode(hi, lo, exact):
    v = hi.double()
    if exact and lo is not None:
        v = v + lo.double() / 2048.0
    return v.permute(0, 3, 1, 2).contiguous()


def run_case(name, desc_mode):
    torch.manual_seed(1234)
    dev = "cuda"
    cfg = dict(nb=2, h=64, w=64, cin=64, ctot=192, cout=32, exact=False, mb=1, lrelu=True,
               res=0, up=False, nchw=False, time=False)
    cases = {
        "fast32": {},
        "fast32_mb2": dict(mb=2),
        "exact32": dict(exact=True),
        "fast64_c192": dict(cout=64, cin=192, lrelu=False, res=2),
        "exact64_c192": dict(cout=64, cin=192, exact=True, lrelu=False, res=2),
        "exact32_c96": dict(cin=96, exact=True),
        "fast32_c160_mb2": dict(cin=160, mb=2),
        "up_exact": dict(cout=64, cin=64, ctot=64, exact=True, up=True),
        "up_fast_mb2": dict(cout=64, cin=64, ctot=64, mb=2, up=True),
        "hr_exact_nchw": dict(nb=1, h=256, w=256, cout=64, cin=64, ctot=64, exact=True, lrelu=False, nchw=True),
        "hr_fast_nchw_mb2": dict(nb=1, h=256, w=256, cout=64, cin=64, ctot=64, mb=2, lrelu=False, nchw=True),
        "odd_h": dict(nb=3, h=40, w=128, cin=128, exact=True),
        "time_fast32": dict(nb=64, mb=2, time=True),
        "time_exact32": dict(nb=64, exact=True, time=True),
        "time_fast64_c192": dict(nb=64, cout=64, cin=192, mb=2, lrelu=False, res=1, time=True),
        "time_exact64_c192": dict(nb=64, cout=64, cin=192, exact=True, lrelu=False, res=1, time=True),
        "time_fast32_mb1": dict(nb=64, mb=1, time=True),
    }
    cases["time_exact32_mb2"] = dict(nb=64, exact=True, mb=2, time=True)
    cases["time_exact64_c192_mb2"] = dict(nb=64, cout=64, cin=192, exact=True, mb=2, lrelu=False, res=1, time=True)
    cases["time_exact32_c160_mb2"] = dict(nb=64, cin=160, exact=True, mb=2, time=True)
    cases["time_exact32_c160"] = dict(nb=64, cin=160, exact=True, mb=1, time=True)
    cases["time_fast32_c160_mb2"] = dict(nb=64, cin=160, mb=2, time=True)
    cases["exact32_mb2"] = dict(exact=True, mb=2)
    cases["exact32_c160_mb2"] = dict(cin=160, exact=True, mb=2, nb=3)
    cases["odd_h_mb2"] = dict(nb=3, h=40, w=128, cin=128, exact=True, mb=2)
    cases["fast32_c96_mb2"] = dict(cin=96, mb=2, nb=3)
    cases["time_exact32_mb2_nb16"] = dict(nb=16, exact=True, mb=2, time=True)
    cases["time_exact32_c160_mb2_nb16"] = dict(nb=16, cin=160, exact=True, mb=2, time=True)
    cases["time_exact32_mb2_nb32"] = dict(nb=32, exact=True, mb=2, time=True)
    cases["time_exact32_mb2_ct64"] = dict(nb=64, exact=True, mb=2, ctot=64, time=True)
    cases["time_exact64_c192_mb2_nb16"] = dict(nb=16, cout=64, cin=192, exact=True, mb=2, lrelu=False, res=1, time=True)
    cases["time_fast32_ct64"] = dict(nb=64, mb=2, ctot=64, time=True)
    cases["time_exact32_c32_ct32"] = dict(nb=64, cin=32, ctot=32, exact=True, mb=2, time=True)
    cases["time_exact32_c32_ct192"] = dict(nb=64, cin=32, ctot=192, exact=True, mb=2, time=True)
    cases["time_fast64_ct64"] = dict(nb=64, cout=64, cin=64, ctot=64, mb=2, lrelu=False, time=True)
    cases["time_fast64_ct192"] = dict(nb=64, cout=64, cin=64, ctot=192, mb=2, lrelu=False, time=True)
    cases["time_exact32_c96"] = dict(nb=64, cin=96, exact=True, mb=1, time=True)
    cases["time_exact32_c96_mb2"] = dict(nb=64, cin=96, exact=True, mb=2, time=True)
    cases["time_exact32_c128"] = dict(nb=64, cin=128, exact=True, mb=1, time=True)
    cases["time_exact32_c128_mb2"] = dict(nb=64, cin=128, exact=True, mb=2, time=True)
    cases["exact64_c192_mb2"] = dict(cout=64, cin=192, exact=True, mb=2, lrelu=False, res=2)
    cases["exact64_c64_mb2_nb6"] = dict(nb=6, cout=64, cin=64, exact=True, mb=2, lrelu=False, res=1, max_ctas=4)
    cases["exact32_c160_mb2_nb4"] = dict(cin=160, exact=True, mb=2, nb=4)
    cases["exact32_c160_mb2_nb6s"] = dict(cin=160, exact=True, mb=2, nb=6, max_ctas=80)
    cases["exact32_c96_w130_nb2"] = dict(cin=96, exact=True, mb=2, nb=2, h=23, w=130)
    for mbv in (3, 4):   # tall tiles of the single-accumulator dx kernel (conv_dxs.cuh)
        cases[f"fast32_mb{mbv}"] = dict(mb=mbv)
        cases[f"fast32_c160_mb{mbv}"] = dict(cin=160, mb=mbv, nb=3, max_ctas=5)
        cases[f"fast32_c96_w130_mb{mbv}"] = dict(cin=96, mb=mbv, nb=2, h=23, w=130)
        cases[f"fast32_c32_h7_mb{mbv}"] = dict(cin=32, ctot=32, mb=mbv, nb=5, h=7, w=40, max_ctas=3)
        for cc in (64, 96, 128, 160):
            cases[f"time_fast32_c{cc}_mb{mbv}"] = dict(nb=64, cin=cc, mb=mbv, time=True)
    cases["time_fast32_c96_mb2"] = dict(nb=64, cin=96, mb=2, time=True)
    cases["time_fast32_c128_mb2"] = dict(nb=64, cin=128, mb=2, time=True)
    for cc in (16, 32, 64, 96):     # 64-output exact convs: CTA-pair kernel on even batches (cin % 32 == 0), per-tap otherwise
        for nbv in (2, 3):
            cases[f"exact64_c{cc}_nb{nbv}"] = dict(cout=64, cin=cc, ctot=96 if cc == 96 else 64, exact=True, mb=2, nb=nbv, lrelu=False, h=16, w=16)
            cases[f"exact64_c{cc}_nb{nbv}_nchw"] = dict(cout=64, cin=cc, ctot=96 if cc == 96 else 64, exact=True, mb=2, nb=nbv, lrelu=False, h=16, w=16, nchw=True)
    # 16-output variant of the exact dx kernel (the head's 16-channel convs) vs the same layer padded to 32 outputs
    cases["exact16_c16"] = dict(cout=16, cin=16, ctot=32, exact=True, mb=2, nb=3, h=40, w=130)
    cases["exact16_c64"] = dict(cout=16, cin=64, ctot=64, exact=True, mb=2, nb=2, lrelu=False)
    cases["exact16_c32_small"] = dict(cout=16, cin=32, ctot=32, exact=True, mb=2, nb=5, h=7, w=40, max_ctas=3)
    cases["time_exact16_c16_256"] = dict(cout=16, cin=16, ctot=32, exact=True, mb=2, nb=32, h=256, w=256, time=True)
    cases["time_exact32_c16_256"] = dict(cout=32, cin=16, ctot=32, exact=True, mb=2, nb=32, h=256, w=256, time=True)
    # 16-channel planes: conv_dx_kernel<..., C16> (16-channel chunks, SWIZZLE_32B) reads and writes contiguous 32-byte records
    cases["exact16_t16"] = dict(cout=16, cin=16, ctot=16, exact=True, mb=2, nb=3, h=40, w=130, octot=16, ochoff=0)
    cases["exact16_t16_small"] = dict(cout=16, cin=16, ctot=16, exact=True, mb=2, nb=5, h=7, w=40, max_ctas=3, octot=16, ochoff=0)
    cases["exact16_t48_off16"] = dict(cout=16, cin=16, ctot=16, exact=True, mb=2, nb=2, octot=48, ochoff=16)
    cases["time_exact16_t16_256"] = dict(cout=16, cin=16, ctot=16, exact=True, mb=2, nb=32, h=256, w=256, time=True, octot=16, ochoff=0)
    cases["time_exact16_c16_256_o32"] = dict(cout=16, cin=16, ctot=32, exact=True, mb=2, nb=32, h=256, w=256, time=True, octot=32, ochoff=0)
    # output layout: 64-byte results into 384-byte records (the RDB concat buffer) vs a compact 32-channel plane
    for cc in (64, 96, 128, 160):
        cases[f"time_exact32_c{cc}_o32"] = dict(nb=64, cin=cc, exact=True, mb=2, time=True, octot=32, ochoff=0)
    cases["time_fast32_c64_o32"] = dict(nb=64, cin=64, mb=2, time=True, octot=32, ochoff=0)
    cases["small_multi"] = dict(nb=8, max_ctas=4)
    cases["small_multi_mb2"] = dict(nb=8, max_ctas=4, mb=2)
    cases["small_multi_exact"] = dict(nb=8, max_ctas=4, exact=True)
    cfg["max_ctas"] = 0
    cfg.update(cases[name])
    if os.environ.get("PROBE_MAX_CTAS"):
        cfg["max_ctas"] = int(os.environ["PROBE_MAX_CTAS"])
    def note(msg):
        print("#", msg, file=sys.stderr, flush=True)
    note("start " + name)
    nb, h, w, cin, ctot, cout = (cfg[k] for k in ("nb", "h", "w", "cin", "ctot", "cout"))
    exact = cfg["exact"]
    numerics = NUMERICS_EXACT if exact else NUMERICS_FAST
    x = torch.rand(nb, cin, h, w, device=dev) * 2 - 0.5
    wt = torch.randn(cout, cin, 3, 3, device=dev) * (0.3 / (cin ** 0.5))
    bias = torch.randn(cout, device=dev) * 0.1
    in_hi, in_lo = to_planes(x, ctot)
    xd = decode(in_hi, in_lo, exact)[:, :cin]
    wd = wt.double() if exact else wt.to(torch.float16).double()
    res = []
    for i in range(cfg["res"]):
        r = torch.randn(nb, 64, h, w, device=dev)
        res.append(to_planes(r, 64))

    if cfg["up"]:
        oh, ow = 2 * h, 2 * w
        ref = F.conv2d(F.interpolate(xd, scale_factor=2, mode="nearest"), wt.double() if exact else None,
                       bias.double(), padding=1) if exact else None
        out_hi = torch.zeros(nb, oh, ow, cout, dtype=torch.float16, device=dev)
        out_lo = torch.zeros_like(out_hi)
        refs = torch.zeros(nb, cout, oh, ow, dtype=torch.float64, device=dev)
        for a in range(2):
            for b in range(2):
                wp = ops.pack_conv_weights(wt, numerics, fold_phase=2 * a + b)
                ops.conv_tc(in_hi, in_lo, 0, cin, wp, cout, bias, ops.phase_taps(a, b), out_hi, out_lo,
                            out_scale=2, out_oy=a, out_ox=b, lrelu=cfg["lrelu"], numerics=numerics,
                            mblocks=cfg["mb"], desc_mode=desc_mode)
        if not exact:
            # fast mode folds fp32 weight sums then rounds to fp16: rebuild that reference per phase
            for a in range(2):
                for b in range(2):
                    taps = ops.phase_taps(a, b)
                    acc = torch.zeros(nb, cout, h, w, dtype=torch.float64, device=dev)
                    xp = F.pad(xd, (1, 1, 1, 1))
                    for iy in range(2):
                        for ix in range(2):
                            kys = ([0], [1, 2]) if a == 0 else ([0, 1], [2])
                            kxs = ([0], [1, 2]) if b == 0 else ([0, 1], [2])
                            wf = sum(wt[:, :, ky, kx] for ky in kys[iy] for kx in kxs[ix])
                            wf = wf.to(torch.float16).double()
                            dy, dx = taps[iy * 2 + ix]
                            patch = xp[:, :, 1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
                            acc += torch.einsum("oc,nchw->nohw", wf, patch)
                    refs[:, :, a::2, b::2] = acc + bias.double().view(1, -1, 1, 1)
            ref = refs
        if cfg["lrelu"]:
            ref = F.leaky_relu(ref, 0.2)
        got = decode(out_hi, out_lo, True)
    else:
        ref = F.conv2d(xd, wd, bias.double(), padding=1)
        if cfg["lrelu"]:
            ref = F.leaky_relu(ref, 0.2)
        kw = {}
        if cfg["res"] >= 1:
            ref = ref * 0.2 + decode(*res[0], True)
            kw.update(res1=(res[0][0], res[0][1], 0), alpha1=0.2)
        if cfg["res"] >= 2:
            ref = ref * 0.2 + decode(*res[1], True)
            kw.update(res2=(res[1][0], res[1][1], 0), alpha2=0.2)
        wp = ops.pack_conv_weights(wt, numerics)
        if cfg["nchw"]:
            out = torch.zeros(nb, cout, h, w, device=dev)
            call = lambda: ops.conv_tc(in_hi, in_lo, 0, cin, wp, cout, bias, ops.PLAIN_TAPS, None, None,
                                       out_f32=out, lrelu=cfg["lrelu"], numerics=numerics,
                                       mblocks=cfg["mb"], desc_mode=desc_mode, **kw)
        else:
            octot = cfg.get("octot", 192)
            ochoff = cfg.get("ochoff", 64)
            out_hi = torch.zeros(nb, h, w, octot, dtype=torch.float16, device=dev)
            out_lo = torch.zeros_like(out_hi)
            call = lambda: ops.conv_tc(in_hi, in_lo, 0, cin, wp, cout, bias, ops.PLAIN_TAPS, out_hi, out_lo,
                                       out_choff=ochoff, lrelu=cfg["lrelu"], numerics=numerics,
                                       mblocks=cfg["mb"], desc_mode=desc_mode, max_ctas=cfg["max_ctas"], **kw)
        torch.cuda.synchronize()
        note("inputs ready")
        call()
        torch.cuda.synchronize()
        note("first call done")
        if cfg["nchw"]:
            got = out.double()
        else:
            got = decode(out_hi, out_lo, True)[:, ochoff:ochoff + cout]
            untouched = (float(out_hi[..., :ochoff].abs().max()) if ochoff else 0.0) + (float(out_hi[..., ochoff + cout:].abs().max()) if ochoff + cout < octot else 0.0)
        if cfg["time"]:
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            iters = 20
            for _ in range(iters):
                call()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            flops = 2.0 * nb * h * w * cout * cin * 9
            print(json.dumps({"case": name, "ms": ms, "tflops": flops / ms / 1e9}))
            if os.environ.get("BHSR_DEBUG_TIMING") == "1":
                import ctypes, numpy as np
                from bhsr import _lib
                buf = np.zeros((148, 8), dtype=np.int64)
                _lib.check(_lib.load().bhsr_debug_timing(buf.ctypes.data, 148))
                tiles = buf[:, 4].clip(min=1)
                print(json.dumps({"case": name, "mma_warp_cycles_per_tile": {
                    "total": float((buf[:, 0] / tiles).mean()), "wait_tmem_empty": float((buf[:, 1] / tiles).mean()),
                    "wait_act": float((buf[:, 2] / tiles).mean()), "wait_weights": float((buf[:, 3] / tiles).mean()),
                    "tiles_per_cta": float(tiles.mean()),
                    "prologue_cycles": float(buf[:, 5].mean()), "epi_wait_per_tile": float((buf[:, 6] / tiles).mean()),
                    "kernel_cycles_mean": float(buf[:, 7].mean()), "kernel_cycles_max": float(buf[:, 7].max()),
                    "mma_loop_max": float(buf[:, 0].max())}}))
    torch.cuda.synchronize()
    err = (got - ref).abs()
    tol = 1e-5 + 1e-4 * ref.abs()  # loose here: fp16 hi/lo output quantisation is ~2^-22 relative
    out = {"case": name, "desc_mode": desc_mode, "max_abs_err": float(err.max()),
           "max_ref": float(ref.abs().max()), "frac_bad": float((err > tol).double().mean()),
           "rel_l2": float((got - ref).norm() / ref.norm())}
    if not cfg["up"] and not cfg["nchw"]:
        out["untouched_abs"] = untouched
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    run_case(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)

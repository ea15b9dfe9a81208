"""TEST / BENCH INFRASTRUCTURE — stages the UNMODIFIED reference modules for the GPU box (SURVEY.md §8c "Staging").

`/root/reference` exists in the build container only.  `stage()` copies the handful of reference source files the
hot path lives in — byte for byte, nothing edited — into `baseline/_ref/` (git-ignored, NOT gpurun-ignored: it travels
to the GPU box with the working tree but never enters the history), and `load()` imports them from there.  With the
staged copy present

  * `bench.py --impl reference` and the `cpu_baseline` leg time the reference's OWN `RRDBNet.forward_feature`
    (SR/rrdbnet_arch.py:227-240) on the host cores (`cpu_baseline.kind = "reference"`), and
  * `tests/test_reference_live.py` checks the oracle port and the CUDA path against the reference run live.

Without it (a fresh clone) both fall back to the oracle port `oracle/ref_torch.py` (`kind = "port"`).  Only `tests/`,
`bench.py`'s reference / cpu_baseline legs and `__graft_entry__.build()` (staging = building the checker) may import
this module; the product package never does.
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")
# the files SURVEY.md §8c lists (mymodels.py does not parse — IndentationError at :467 — and is staged only so that
# tests can exec its hot class from the source slice, as tests/golden/make_golden.py does)
FILES = (
    "SR/rrdbnet_arch.py", "SR/srloss.py", "SR/HRfuse.py", "SR/RRDBNet.py", "SR/edsr.py",
    "aggregate_utils.py", "mymodels.py", "losses_pytorch/selfloss.py",
)
DIRS = ("SR/testimg",)


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def stage(ref: str = "/root/reference", dst: str = STAGED) -> bool:
    """Copy the reference files into `dst`; returns False (and does nothing) when `ref` is absent."""
    if not os.path.isdir(ref):
        return False
    manifest = {}
    for rel in FILES:
        src = os.path.join(ref, rel)
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(src, out)
        manifest[rel] = _sha(out)
    for rel in DIRS:
        src = os.path.join(ref, rel)
        if os.path.isdir(src):
            out = os.path.join(dst, rel)
            os.makedirs(out, exist_ok=True)
            for name in sorted(os.listdir(src)):
                if os.path.isfile(os.path.join(src, name)):
                    shutil.copyfile(os.path.join(src, name), os.path.join(out, name))
                    manifest[f"{rel}/{name}"] = _sha(os.path.join(out, name))
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": ref, "sha256": manifest}, f, indent=1, sort_keys=True)
    return True


def available(dst: str = STAGED) -> bool:
    return os.path.isfile(os.path.join(dst, "SR", "rrdbnet_arch.py"))


_loaded = {}


def load(dst: str = STAGED):
    """Import the staged reference modules under PRIVATE module names (`_bhsr_ref.*`), so that neither this repo's
    drop-in `SR` package nor `aggregate_utils` shim is shadowed.  Returns a namespace with `.arch`
    (SR/rrdbnet_arch.py), `.old` (SR/RRDBNet.py), `.hrfuse` (SR/HRfuse.py), `.aggregate` (aggregate_utils.py) and
    `.selfloss` (losses_pytorch/selfloss.py).  The reference's `from SR.srloss import ...` is satisfied by binding the
    name `SR` to the staged package for the duration of the import only."""
    if dst in _loaded:
        return _loaded[dst]
    if not available(dst):
        raise FileNotFoundError(f"no staged reference under {dst} (run oracle/stage_reference.py in the build container)")
    for name in ("matplotlib", "matplotlib.pyplot", "rasterio"):     # imported, never used on this path
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)

    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "SR" or k.startswith("SR.")}
    try:
        pkg = types.ModuleType("SR")
        pkg.__path__ = [os.path.join(dst, "SR")]
        sys.modules["SR"] = pkg

        def imp(modname, rel):
            spec = importlib.util.spec_from_file_location(modname, os.path.join(dst, rel))
            m = importlib.util.module_from_spec(spec)
            sys.modules[modname] = m
            spec.loader.exec_module(m)
            return m

        imp("SR.srloss", "SR/srloss.py")
        ns = types.SimpleNamespace(
            arch=imp("SR.rrdbnet_arch", "SR/rrdbnet_arch.py"),
            old=imp("SR.RRDBNet", "SR/RRDBNet.py"),
            hrfuse=imp("SR.HRfuse", "SR/HRfuse.py"),
            aggregate=imp("_bhsr_ref_aggregate_utils", "aggregate_utils.py"),
            selfloss=imp("_bhsr_ref_selfloss", "losses_pytorch/selfloss.py"),
            root=dst,
        )
    finally:
        for k in [k for k in sys.modules if k == "SR" or k.startswith("SR.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    _loaded[dst] = ns
    return ns


if __name__ == "__main__":
    ok = stage(*(sys.argv[1:2] or ["/root/reference"]))
    print(f"staged -> {STAGED}" if ok else "reference tree absent: nothing staged")

#!/usr/bin/env python
"""Joules per launch of the RDB conv kernels under the box's power cap, next to cuBLAS bf16 (Finding 8 diagnostic).

Each case is launched back to back for ~3 s while nvidia-smi samples power and SM clock every 100 ms; the first half of
the samples is discarded (ramp, averaging window).  Reported per case: us / launch at the sustained clock, W, MHz,
mJ / launch, and pJ per EXECUTED tensor flop (exact numerics executes 3 fp16 products per algorithmic MAC) — the
figure to hold against cuBLAS's.  Diagnostic only; not part of the product path or of bench.py."""
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bhsr  # noqa: E402,F401
from bhsr import ops  # noqa: E402
from bhsr._lib import NUMERICS  # noqa: E402

SECONDS = float(os.environ.get("PROBE_SECONDS", "3.0"))


class Sampler:
    def __init__(self):
        q = "power.draw.instant,power.draw,clocks.sm"
        self.proc = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.rows = []
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for ln in self.proc.stdout:
            try:
                f = [float(x) for x in ln.strip().split(",")]
                self.rows.append((time.perf_counter(), f))
            except ValueError:
                pass

    def window(self, t0, t1):
        r = [f for t, f in self.rows if t0 <= t <= t1]
        if not r:
            return None
        n = len(r)
        return {"w_instant": sum(x[0] for x in r) / n, "w_avg": sum(x[1] for x in r) / n, "mhz": sum(x[2] for x in r) / n, "samples": n}


def run_case(name, call, flop_exec, sampler):
    for _ in range(10):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start = time.perf_counter()
    n = 0
    e0.record()
    while time.perf_counter() - t_start < SECONDS:
        for _ in range(200):
            call()
        n += 200
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    us = e0.elapsed_time(e1) / n * 1e3
    w = sampler.window(t_start + 0.5 * (t_end - t_start), t_end)
    rec = {"case": name, "us_per_launch": us, "launches": n, "tflops_executed": flop_exec / (us * 1e-6) / 1e12}
    if w:
        watts = max(w["w_instant"], w["w_avg"])
        rec.update(w)
        rec["mj_per_launch"] = watts * us * 1e-3 * 1e-3 * 1e3 / 1e3
        rec["pj_per_executed_flop"] = watts * us * 1e-6 / flop_exec * 1e12
    print(json.dumps(rec), flush=True)
    time.sleep(1.0)


def main():
    dev = torch.device("cuda:0")
    B = 64
    sampler = Sampler()
    time.sleep(0.5)
    g = torch.Generator(device="cpu").manual_seed(7)
    hi = (torch.randn((B, 64, 64, 192), generator=g) * 0.5).to(torch.float16).to(dev)
    lo = (torch.randn((B, 64, 64, 192), generator=g) * 0.5).to(torch.float16).to(dev)
    out_hi, out_lo = torch.zeros_like(hi), torch.zeros_like(hi)
    # idle floor
    t0 = time.perf_counter(); time.sleep(1.5)
    print(json.dumps({"case": "idle", **(sampler.window(t0 + 0.7, time.perf_counter()) or {})}), flush=True)
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
    b = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
    c = torch.empty(n, n, device=dev, dtype=torch.bfloat16)
    run_case("cublas_bf16_8192", lambda: torch.matmul(a, b, out=c), 2.0 * n ** 3, sampler)
    for numerics, layers in (("exact", range(5)), ("fast", (0, 3, 4))):
        num = NUMERICS[numerics]
        for ci in layers:
            cin, cout = 64 + 32 * ci, (32 if ci < 4 else 64)
            w = torch.randn((cout, cin, 3, 3), generator=g).to(dev) * 0.01
            bias = torch.zeros(cout, device=dev)
            wp = ops.pack_conv_weights(w, num)
            kw = dict(lrelu=True) if ci < 4 else dict(res1=(hi, lo, 0), alpha1=0.2)

            def call(cin=cin, cout=cout, wp=wp, bias=bias, kw=kw, ci=ci, num=num):
                ops.conv_tc(hi, lo, 0, cin, wp, cout, bias, ops.PLAIN_TAPS, out_hi, out_lo,
                            out_choff=(cin if ci < 4 else 0), numerics=num, **kw)
            flop = 2.0 * B * 64 * 64 * cout * cin * 9 * (3 if numerics == "exact" else 1)
            run_case(f"{numerics}.rdb.conv{ci + 1}", call, flop, sampler)
    sampler.proc.terminate()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Build profiles/<round>_summary.md (round = $BHSR_ROUND, default r02) from the artefacts a gpurun profile call brought back
(gpurun_out/launches_*.csv, gpurun_out/prof_rdb_*.ncu-rep) plus the saved bench lines."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
ROUND = os.environ.get("BHSR_ROUND", "r02")


def launch_table(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    by, tot, seq = collections.OrderedDict(), 0.0, []
    for r in data:
        v = float(r[ival].replace(",", ""))
        v = v / 1000 if r[iunit] == "ns" else (v * 1000 if r[iunit] == "ms" else v)
        name = re.sub(r"\(.*", "", r[iname]).replace("void bhsr::", "").replace("void ", "")
        by.setdefault(name, [0, 0.0])
        by[name][0] += 1
        by[name][1] += v
        tot += v
        seq.append((name, v))
    return by, tot, seq


def ncu_raw(rep):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
            "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
    idx = {w: hdr.index(w) for w in want if w in hdr}
    out = []
    for r in data:
        out.append({w: (r[i], units[i]) for w, i in idx.items()})
    return out


def main():
    lines = [f"# Round {int(ROUND[1:])} profile summary (B200, sm_100a)", "",
             "All numbers below come from `gpurun` calls on a B200; ncu times are cold-cache and serialised "
             "(compare shares, not absolutes); bench lines are CUDA-event timings outside any profiler.", ""]
    for mode in ("exact", "fast"):
        path = os.path.join(OUT, f"launches_{mode}.csv")
        if not os.path.exists(path):
            continue
        by, tot, seq = launch_table(path)
        lines += [f"## Launch list of one `forward_feature` step, B=64, numerics={mode}",
                  f"`ncu --metrics gpu__time_duration.sum --clock-control none` on `bench.py --steps 1 --warmup 3 --numerics {mode} --no-graph`; "
                  f"{sum(n for n, _ in by.values())} launches, {tot / 1000:.2f} ms total.", "",
                  "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
        for k, (n, t) in by.items():
            lines.append(f"| `{k}` | {n} | {t:.0f} | {100 * t / tot:.1f}% | {t / n:.1f} |")
        first = [f"{v:.1f}" for _, v in seq[1:6]]
        lines += ["", f"First RDB (conv1..conv5, us): {', '.join(first)}; conv_hr (last launch): {seq[-1][1]:.1f} us.", ""]
        rep = os.path.join(OUT, f"prof_rdb_{mode}.ncu-rep")
        if os.path.exists(rep):
            lines += [f"### `ncu --set full` capture of 10 consecutive trunk convs ({mode})", "",
                      "| layer | kernel | us | tensor pipe active % | smem tensor-read wavefronts % | smem bank writes % | DRAM read MB | DRAM write MB | L2->SM MB | L2 hit % | SM GHz | regs |",
                      "|---|---|---|---|---|---|---|---|---|---|---|---|"]
            caps = ncu_raw(rep)
            names = [re.sub(r"\(CUtensor.*", "", k.get("Kernel Name", ("", ""))[0]).replace("void bhsr::", "").replace("void ", "") for k in caps]
            # conv5 is the only 64-output conv inside the trunk (CTA-pair kernel): layers follow in order
            i5 = next((i for i, n in enumerate(names) if n.startswith("conv_pair_kernel") or n.startswith("conv_tc_kernel<64")), None)
            layer_of = {}
            if i5 is not None:
                for i in range(len(caps)):
                    layer_of[i] = "rdb.conv%d" % ((i - i5 - 1) % 5 + 1)
            kjson = {}
            for i, k in enumerate(caps):
                g = lambda w: k.get(w, ("", ""))[0]
                lay = layer_of.get(i, "?")
                lines.append(f"| {lay} | `{names[i]}` | {g('gpu__time_duration.sum')} | {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')} | "
                             f"{g('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed')} | {g('l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed')} | "
                             f"{g('dram__bytes_read.sum')} | {g('dram__bytes_write.sum')} | {g('l1tex__m_xbar2l1tex_read_bytes.sum')} | {g('lts__t_sector_hit_rate.pct')} | "
                             f"{g('sm__cycles_elapsed.avg.per_second')} | {g('launch__registers_per_thread')} |")
                try:
                    unit_r = k.get("dram__bytes_read.sum", ("", ""))[1]
                    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(unit_r, 1e6)
                    dram = (float(g("dram__bytes_read.sum").replace(",", "")) + float(g("dram__bytes_write.sum").replace(",", ""))) * scale
                    if lay != "?" and lay not in kjson:
                        kjson[lay] = {"dram_bytes": dram, "us_under_ncu": float(g("gpu__time_duration.sum").replace(",", "")),
                                      "tensor_pipe_active_pct": float(g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")),
                                      "source": f"profiles/{ROUND}_summary.md (ncu --set full, {mode}, cold L2 per replay)"}
                except ValueError:
                    pass
            lines.append("")
            allk = {}
            kpath = os.path.join(PROF, f"{ROUND}_ncu_kernels.json")
            if os.path.exists(kpath):
                allk = json.load(open(kpath))
            allk[mode] = kjson
            json.dump(allk, open(kpath, "w"), indent=1)
    for name in sorted(os.listdir(PROF)):
        if name.startswith(f"{ROUND}_bench") and name.endswith(".json"):
            try:
                d = json.loads(open(os.path.join(PROF, name)).read().strip().split("\n")[0])
            except Exception:
                continue
            if "value" in d:
                extra = d.get("other_numerics", {})
                lines.append(f"* `{name}`: {d.get('config', {}).get('numerics', d.get('impl', ''))} {d['value']:.0f} tiles/s "
                             f"({d.get('ms_per_step', 0):.2f} ms/step, roofline frac {d.get('roofline', {}).get('frac', 0):.3f})"
                             + (f"; {extra.get('numerics')} {extra.get('value', 0):.0f} tiles/s" if extra else ""))
    open(os.path.join(PROF, f"{ROUND}_summary.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()

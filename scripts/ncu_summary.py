#!/usr/bin/env python
"""Turn one gpurun round (gpurun_out/<tag>_launches.csv + <tag>_prof.ncu-rep + <tag>_bench_1gpu.json) into the
tracked files under profiles/: <tag>_launches.csv (copy), <tag>_traffic.json (per-launch DRAM bytes of every captured
kernel, what bench.py reports as roofline.traffic) and <tag>_summary.md.

    python scripts/ncu_summary.py r01w ["free-text title"]
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

# kernel name prefix -> algorithmic bytes for config 2
N, M = 10_000_000, 1_000_000


def launch_shares(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[ui] == "ns":
            v /= 1000.0
        k = r[ki].split("(")[0][:72]
        a = agg.setdefault(k, [0.0, 0])
        a[0] += v
        a[1] += 1
    tot = sum(v[0] for v in agg.values())
    return [(k, v / tot * 100, n, v / n) for k, (v, n) in sorted(agg.items(), key=lambda x: -x[1][0])]


def full_capture(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def g(r, name, scale_unit=None):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return None
        v = float(r[i].replace(",", ""))
        u = units[i]
        if scale_unit == "bytes":
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        if scale_unit == "us":
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)  # ncu names: ns/us/ms
        return v

    seen = collections.OrderedDict()
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        seen.setdefault(name, []).append(r)
    out = collections.OrderedDict()
    for name, rs in seen.items():
        def avg(metric, su=None):
            vals = [g(r, metric, su) for r in rs]
            vals = [v for v in vals if v is not None]
            return sum(vals) / len(vals) if vals else None
        out[name] = {
            "launches_captured": len(rs),
            "time_us": avg("gpu__time_duration.sum", "us"),
            "dram_read_bytes": avg("dram__bytes_read.sum", "bytes"),
            "dram_write_bytes": avg("dram__bytes_write.sum", "bytes"),
            "dram_pct": avg("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "l2_hit_pct": avg("lts__t_sector_hit_rate.pct"),
            "lts_pct": avg("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            "l1tex_pct": avg("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            "l1tex2xbar_req_pct": avg("l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "sm_pct": avg("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": avg("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "regs": avg("launch__registers_per_thread"),
            "grid": avg("launch__grid_size"),
            "block": avg("launch__block_size"),
            "lts_requests_from_tex": avg("lts__t_requests_srcunit_tex.sum"),
        }
        o = out[name]
        o["traffic_bytes"] = (o["dram_read_bytes"] or 0) + (o["dram_write_bytes"] or 0)
    return out


def main():
    tag = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else ""
    if os.path.exists(os.path.join(OUT, f"{tag}_config3_launches.csv")) or os.path.exists(os.path.join(OUT, f"{tag}_config3_prof.ncu-rep")):
        return main_config3(tag, title)
    os.makedirs(PROF, exist_ok=True)
    lines = [f"# {tag}: {title}".rstrip(": "), ""]
    lcsv = os.path.join(OUT, f"{tag}_launches.csv")
    if os.path.exists(lcsv):
        shutil.copy(lcsv, os.path.join(PROF, f"{tag}_launches.csv"))
        lines += ["Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 "
                  "--warmup 3 --no-cpu-baseline --skip-e2e` (config 2; 5 steps captured incl. warm-up). Per-launch times are "
                  "cold-cache and serialised: compare SHARES, not absolutes.", "",
                  "| share | launches | avg us | kernel |", "|---:|---:|---:|---|"]
        for k, sh, n, a in launch_shares(lcsv):
            lines.append(f"| {sh:.2f}% | {n} | {a:.2f} | `{k}` |")
        lines.append("")
    rep = os.path.join(OUT, f"{tag}_prof.ncu-rep")
    if os.path.exists(rep):
        cap = full_capture(rep)
        json.dump(cap, open(os.path.join(PROF, f"{tag}_traffic.json"), "w"), indent=1)
        lines += [f"## `ncu --set full --clock-control none --import-source on` ({tag}_prof.ncu-rep; averages over the captured "
                  "launches; ncu flushes caches between replays, so DRAM traffic is the cold-cache worst case)", "",
                  "| kernel | n | time us | dram rd MB | dram wr MB | dram % | L2 hit % | lts % | l1tex % | l1tex->xbar req % | sm % | warps act % | regs | grid x block |",
                  "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
        f = lambda v, d=1: "-" if v is None else f"{v:.{d}f}"
        for k, o in cap.items():
            lines.append(f"| `{k}` | {o['launches_captured']} | {f(o['time_us'])} | {f((o['dram_read_bytes'] or 0)/1e6)} | "
                         f"{f((o['dram_write_bytes'] or 0)/1e6)} | {f(o['dram_pct'])} | {f(o['l2_hit_pct'])} | {f(o['lts_pct'])} | "
                         f"{f(o['l1tex_pct'])} | {f(o['l1tex2xbar_req_pct'])} | {f(o['sm_pct'])} | {f(o['warps_active_pct'])} | "
                         f"{f(o['regs'],0)} | {f(o['grid'],0)} x {f(o['block'],0)} |")
        lines.append("")
    bj = os.path.join(OUT, f"{tag}_bench_1gpu.json")
    if os.path.exists(bj):
        txt = open(bj).read().strip().splitlines()
        if txt:
            d = json.loads(txt[-1])
            shutil.copy(bj, os.path.join(PROF, f"{tag}_bench_1gpu.json"))
            r = d["roofline"]
            lines += [f"## bench.py of the same build (not under ncu; `profiles/{tag}_bench_1gpu.json`)", "",
                      f"* value {d['value']/1e9:.2f} G pairs/s, {d['ms_per_step']:.3f} ms/step, {d['gpu_launches']} launches / "
                      f"{d['steps']} steps, SM clock {d['clocks']['sm_mhz']} MHz, reasons {d['clocks']['reasons']}",
                      f"* e2e {d['e2e']['value']/1e6:.0f} M pairs/s ({d['e2e'].get('ms_per_step', 0):.2f} ms/step; H2D "
                      f"{d['e2e']['h2d_bytes_per_step']/1e6:.0f} MB, D2H {d['e2e']['d2h_bytes_per_step']/1e6:.0f} MB per step)",
                      f"* cpu_baseline {d['cpu_baseline']['value']/1e6:.1f} M pairs/s on {d['cpu_baseline']['cores']} threads ({d['cpu_baseline']['kind']})",
                      f"* roofline: {r['kernel']}: {r['achieved']:.0f} GB/s algorithmic of {r['peak']:.0f} GB/s = {r['frac']:.3f}"]
            for k, v in r.get("all_kernels", {}).items():
                lines.append(f"  * {k}: {v['ms']*1000:.1f} us, {v['GBps']:.0f} GB/s algorithmic")
            lines.append(f"  * index build {r.get('index_build_ms', 0)*1000:.0f} us, offset scan {r.get('offset_scan_ms', 0)*1000:.1f} us")
            lines.append("")
    rj = os.path.join(OUT, f"{tag}_bench_ref.json")
    if os.path.exists(rj):
        shutil.copy(rj, os.path.join(PROF, f"{tag}_bench_ref.json"))
    open(os.path.join(PROF, f"{tag}_summary.md"), "w").write("\n".join(lines))
    print("\n".join(lines))


def main_config3(tag, title):
    """Round 2: captures of tests/tools/profile_config3.py (one index build + count_overlaps + two-pass overlap on
    BASELINE config 3) -> profiles/<tag>_config3_{launches.csv,traffic.json,summary.md}."""
    os.makedirs(PROF, exist_ok=True)
    lines = [f"# {tag} (config 3: 100 M reads x 90 M variants, one B200): {title}".rstrip(": "), ""]
    lcsv = os.path.join(OUT, f"{tag}_config3_launches.csv")
    if os.path.exists(lcsv):
        shutil.copy(lcsv, os.path.join(PROF, f"{tag}_config3_launches.csv"))
        lines += ["Command: `PB_REPS=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python "
                  "tests/tools/profile_config3.py` (two repetitions of: index build, count_overlaps, overlap pass 1 + 2). Per-launch "
                  "times are cold-cache and serialised: compare SHARES, not absolutes.", "",
                  "| share | launches | avg us | kernel |", "|---:|---:|---:|---|"]
        for k, sh, n, a in launch_shares(lcsv):
            lines.append(f"| {sh:.2f}% | {n} | {a:.2f} | `{k}` |")
        lines.append("")
    rep = os.path.join(OUT, f"{tag}_config3_prof.ncu-rep")
    if os.path.exists(rep):
        cap = full_capture(rep)
        meta = {"_meta": {"captured_unix": int(os.path.getmtime(rep)), "tag": tag}}  # bench.py reports the newest capture
        json.dump({**cap, **meta}, open(os.path.join(PROF, f"{tag}_config3_traffic.json"), "w"), indent=1)
        lines += [f"## `ncu --set full --clock-control none --import-source on` ({tag}_config3_prof.ncu-rep; averages over the captured "
                  "launches; ncu flushes caches between replays, so DRAM traffic is the cold-cache worst case)", "",
                  "| kernel | n | time us | dram rd MB | dram wr MB | dram % | L2 hit % | lts % | l1tex % | l1tex->xbar req % | sm % | warps act % | regs | grid x block |",
                  "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
        f = lambda v, d=1: "-" if v is None else f"{v:.{d}f}"
        for k, o in cap.items():
            lines.append(f"| `{k[:70]}` | {o['launches_captured']} | {f(o['time_us'])} | {f((o['dram_read_bytes'] or 0)/1e6)} | "
                         f"{f((o['dram_write_bytes'] or 0)/1e6)} | {f(o['dram_pct'])} | {f(o['l2_hit_pct'])} | {f(o['lts_pct'])} | "
                         f"{f(o['l1tex_pct'])} | {f(o['l1tex2xbar_req_pct'])} | {f(o['sm_pct'])} | {f(o['warps_active_pct'])} | "
                         f"{f(o['regs'],0)} | {f(o['grid'],0)} x {f(o['block'],0)} |")
        lines.append("")
    for extra in ("bench_1gpu", "bench_dev"):
        bj = os.path.join(OUT, f"{tag}_{extra}.json")
        if os.path.exists(bj) and open(bj).read().strip():
            shutil.copy(bj, os.path.join(PROF, f"{tag}_{extra}.json"))
    open(os.path.join(PROF, f"{tag}_config3_summary.md"), "w").write("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main()

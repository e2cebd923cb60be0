"""Turns gpurun_out ncu artefacts into the small, committed summaries under profiles/.
    python tools/summarize_profiles.py <tag> <launches.csv> [<full.ncu-rep>]   # from the raw `ncu --csv --log-file` launch list
    python tools/summarize_profiles.py regen <tag>                               # re-derive launches_<tag>.md and ncu_traffic.json from the
                                                                                 # TRACKED compact per-launch table profiles/launches_<tag>.csv
The launch list comes from (B200_PROFILING.md):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
        --clock-control none -s <warm-up launches> -c <launches of one step> --csv --log-file gpurun_out/launches.csv python bench.py --no-extras ...
"""
import collections
import csv
import subprocess
import sys


def launches(path, out):
    """launch list with gpu__time_duration (+ optional dram bytes) per launch; writes the markdown summary and, when dram metrics are
    present, profiles/ncu_traffic.json (dram bytes per conv_tc launch, read by bench.py for roofline.traffic)."""
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ii, ki, mi, vi, ui = H.index("ID"), H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit")
    per = collections.OrderedDict()
    for r in data:
        d = per.setdefault(r[ii], {"k": r[ki].split("(")[0].replace("void ", "").replace("myolo::", "")})
        v = float(r[vi].replace(",", ""))
        if r[mi].startswith("gpu__time_duration"):
            d["us"] = v / 1000 if r[ui] in ("ns", "nsecond") else (v * 1000 if r[ui] in ("ms", "msecond") else v)
        elif r[mi].startswith("dram__bytes_read"):
            d["rd"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1)
        elif r[mi].startswith("dram__bytes_write"):
            d["wr"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1)
        elif r[mi].startswith("sm__pipe_tensor_cycles_active"):
            d["tc"] = v
    # the compact per-launch table is what gets committed: everything below can be regenerated from it (`regen`)
    compact = out.replace(".md", ".csv")
    with open(compact, "w") as f:
        f.write("launch,kernel,duration_us,dram_read_bytes,dram_write_bytes,tensor_pipe_pct\n")
        for n, d in enumerate(per.values()):
            f.write(f"{n},{d['k'].replace(', ', ';').replace(',', ';')},{d.get('us', 0):.3f},{d.get('rd', -1):.0f},{d.get('wr', -1):.0f},{d.get('tc', -1):.2f}\n")
    summarize(per, out, compact)


def regen(tag):
    per = collections.OrderedDict()
    for r in csv.DictReader(open(f"profiles/launches_{tag}.csv")):
        d = {"k": r["kernel"], "us": float(r["duration_us"])}
        if float(r["dram_read_bytes"]) >= 0:
            d["rd"], d["wr"] = float(r["dram_read_bytes"]), float(r["dram_write_bytes"])
        if float(r["tensor_pipe_pct"]) >= 0:
            d["tc"] = float(r["tensor_pipe_pct"])
        per[r["launch"]] = d
    summarize(per, f"profiles/launches_{tag}.md", f"profiles/launches_{tag}.csv")


def summarize(per, out, path):
    agg, tot = collections.OrderedDict(), 0.0
    for d in per.values():
        a = agg.setdefault(d["k"], [0, 0.0, 0.0]); a[0] += 1; a[1] += d.get("us", 0.0); a[2] += d.get("rd", 0.0) + d.get("wr", 0.0)
        tot += d.get("us", 0.0)
    has_dram = any("rd" in d for d in per.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised): {len(per)} launches, {tot:.1f} us\n\n")
        f.write("| kernel | launches | total us | share |" + (" dram MB (rd+wr) |" if has_dram else "") + "\n|---|---:|---:|---:|" + ("---:|" if has_dram else "") + "\n")
        for k, (n, t, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {t:.1f} | {100 * t / tot:.1f}% |" + (f" {by / 1e6:.1f} |" if has_dram else "") + "\n")
        f.write("\nper-launch durations in stream order (us" + (", tensor-pipe % in brackets" if any("tc" in d for d in per.values()) else "") + "):\n\n")
        f.write(" ".join(f"{d['k'][:12]}:{d.get('us', 0):.1f}" + (f"[{d['tc']:.0f}]" if d.get("tc", -1) >= 0 and d["k"].startswith("conv_tc") else "")
                         for d in per.values()) + "\n")
    if has_dram:
        import json
        conv = [d for d in per.values() if d["k"].startswith("conv_tc")]
        json.dump({"cfg": "s_psp", "batch": 16, "launches": len(conv),
                   "dram_bytes_per_launch": sum(d.get("rd", 0) + d.get("wr", 0) for d in conv) / max(1, len(conv)),
                   "dram_read_MB_total": sum(d.get("rd", 0) for d in conv) / 1e6, "dram_write_MB_total": sum(d.get("wr", 0) for d in conv) / 1e6,
                   "source": path}, open("profiles/ncu_traffic.json", "w"))


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = rows[0]
    want = {"gpu__time_duration.sum": "dur_us", "dram__bytes_read.sum": "dram_rd_MB", "dram__bytes_write.sum": "dram_wr_MB",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_%",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_%", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_%",
            "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum": "tma_ld_MB", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_%",
            "smsp__inst_executed.sum": "warp_insts", "launch__grid_size": "grid", "launch__registers_per_thread": "regs",
            "launch__shared_mem_per_block_dynamic": "smem_KB", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_%"}
    idx = {v: H.index(k) for k, v in want.items() if k in H}
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none (one row per captured launch, stream order)\n\n")
        f.write("| # | " + " | ".join(idx) + " |\n|---|" + "---:|" * len(idx) + "\n")
        tot_rd = tot_wr = 0.0
        nrows = 0
        for n, r in enumerate(rows[2:]):
            f.write(f"| {n} | " + " | ".join(r[i] for i in idx.values()) + " |\n")
            try:
                tot_rd += float(r[idx["dram_rd_MB"]]); tot_wr += float(r[idx["dram_wr_MB"]]); nrows += 1
            except Exception:
                pass


def last_step(tag, path, n):
    """keep only the last n launches of a raw ncu launch list (= one step of tools/one_step.py) and summarise them"""
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ids = sorted({int(r[0]) for r in data})
    keep = set(ids[-n:])
    tmp = path + ".laststep.csv"
    with open(tmp, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(H)
        for r in data:
            if int(r[0]) in keep:
                w.writerow(r)
    launches(tmp, f"profiles/launches_{tag}.md")


def conv_sections(tag, path, perop=None):
    """`ncu --section SpeedOfLight ... --page raw --csv` of EVERY conv_tc launch of one step (tools/capture_profiles.sh) -> a compact tracked
    table profiles/conv_launches_<tag>.csv/.md: duration, tensor-pipe %, issue %, DRAM %, L2 %, grid, smem per launch (+ the layer shape when
    the `#conv` lines of `bench.py --profile-ops` are given)"""
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 2:]
    col = {k: H.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                                   "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                                   "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
                                   "launch__registers_per_thread")}
    unit = rows[hdr + 1][col["gpu__time_duration.sum"]]
    scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
    shapes = []
    if perop:
        shapes = [" ".join(l.split()[2:6]) for l in open(perop) if l.startswith("#conv") and "[1," in l]
    out_csv, out_md = f"profiles/conv_launches_{tag}.csv", f"profiles/conv_launches_{tag}.md"
    with open(out_csv, "w") as f, open(out_md, "w") as g:
        f.write("launch,kernel,layer,duration_us,tensor_pipe_pct,issue_active_pct,dram_pct,l2_pct,grid,smem_bytes,regs\n")
        g.write("# every conv_tc launch of one forward (ncu SpeedOfLight sections, --clock-control none, cold-cache / serialised)\n\n"
                "| # | layer | us | tensor pipe % | issue % | DRAM % | L2 % | grid | smem KB |\n|---:|---|---:|---:|---:|---:|---:|---:|---:|\n")
        tot = 0.0
        for n, r in enumerate(data):
            if len(r) <= max(col.values()):
                continue
            v = lambda k: r[col[k]].replace(",", "")  # noqa: E731
            us = float(v("gpu__time_duration.sum")) * scale
            tot += us
            layer = shapes[n] if n < len(shapes) else ""
            kern = v("Kernel Name").split("(")[0].replace("void ", "").replace(", ", ";")
            f.write(f"{n},{kern},{layer},{us:.2f},{float(v('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')):.1f},"
                    f"{float(v('sm__issue_active.avg.pct_of_peak_sustained_elapsed')):.1f},{float(v('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f},"
                    f"{float(v('lts__throughput.avg.pct_of_peak_sustained_elapsed')):.1f},{v('launch__grid_size')},{v('launch__shared_mem_per_block_dynamic')},{v('launch__registers_per_thread')}\n")
            g.write(f"| {n} | {layer} | {us:.1f} | {float(v('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')):.1f} | "
                    f"{float(v('sm__issue_active.avg.pct_of_peak_sustained_elapsed')):.1f} | {float(v('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
                    f"{float(v('lts__throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | {v('launch__grid_size')} | {float(v('launch__shared_mem_per_block_dynamic')) / 1024:.0f} |\n")
        g.write(f"\nsum of durations {tot:.1f} us\n")
    print("wrote", out_md)


if __name__ == "__main__":
    if sys.argv[1] == "last-step":
        last_step(sys.argv[2], sys.argv[3], int(sys.argv[4]))
        sys.exit(0)
    if sys.argv[1] == "conv-sections":
        conv_sections(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
        sys.exit(0)
    if sys.argv[1] == "regen":
        regen(sys.argv[2])
        sys.exit(0)
    tag = sys.argv[1]
    launches(sys.argv[2], f"profiles/launches_{tag}.md")
    if len(sys.argv) > 3:
        full(sys.argv[3], f"profiles/ncu_full_conv_{tag}.md")

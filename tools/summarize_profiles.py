"""Turns gpurun_out ncu artefacts into the small, committed summaries under profiles/.
    python tools/summarize_profiles.py <tag> <launches.csv> [<full.ncu-rep>]
"""
import collections
import csv
import subprocess
import sys


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    per = []
    for r in data:
        t = float(r[vi].replace(",", ""))
        t = t / 1000 if r[ui] == "ns" else (t * 1000 if r[ui] == "ms" else t)
        k = r[ki].split("(")[0].replace("void ", "").replace("myolo::", "")
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
        per.append((k, t))
    with open(out, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised): {len(data)} launches, {tot:.1f} us\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {t:.1f} | {100 * t / tot:.1f}% |\n")
        f.write("\nper-launch durations in stream order (us):\n\n")
        f.write(" ".join(f"{k[:12]}:{t:.1f}" for k, t in per) + "\n")


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
    if nrows:
        import json
        json.dump({"cfg": "s_psp", "batch": 16, "launches": nrows, "dram_bytes_per_launch": (tot_rd + tot_wr) * 1e6 / nrows,
                   "dram_read_MB_total": tot_rd, "dram_write_MB_total": tot_wr, "source": rep}, open("profiles/ncu_traffic.json", "w"))


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(sys.argv[2], f"profiles/launches_{tag}.md")
    if len(sys.argv) > 3:
        full(sys.argv[3], f"profiles/ncu_full_conv_{tag}.md")

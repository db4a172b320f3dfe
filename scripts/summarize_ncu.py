"""Summarise ncu reports (gpurun_out/*.ncu-rep, launches.csv) into small text files under profiles/.

    python scripts/summarize_ncu.py r01            # prefix for the output files
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__inst_executed.avg.per_cycle_elapsed"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def summarize_report(rep, fh):
    rows = ncu_csv(rep, "raw")
    if len(rows) < 3:
        fh.write(f"{rep}: empty\n")
        return
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        fh.write(f"== {r[idx['Kernel Name']]}  (report {os.path.basename(rep)}, ncu --set full --clock-control none)\n")
        for k in KEEP:
            if k in idx:
                fh.write(f"   {k:85s} {r[idx[k]]:>16s} {units[idx[k]]}\n")
        rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
        fh.write(f"   traffic = dram read + write = {rd + wr:.3f} {units[idx['dram__bytes_read.sum']]} per launch\n")
        stalls = {h: r[i] for h, i in idx.items() if h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h}
        top = sorted(stalls.items(), key=lambda kv: -float(kv[1] or 0))[:6]
        fh.write("   top stall reasons (pc samples): " + ", ".join(f"{k.split('stalled_')[1]}={v}" for k, v in top) + "\n")
    src = ncu_csv(rep, "source")
    if len(src) > 2:
        hdr = src[1]
        idx = {h: i for i, h in enumerate(hdr)}
        sec = []
        for r in src[2:]:
            if len(r) < 10 or r[0] == "Kernel Name":
                break
            sec.append(r)
        tot = sum(int(r[idx["# Samples"]] or 0) for r in sec) or 1
        fh.write(f"   hottest SASS instructions of the first launch ({tot} samples):\n")
        for r in sorted(sec, key=lambda r: -int(r[idx["# Samples"]] or 0))[:8]:
            fh.write(f"     {int(r[idx['# Samples']] or 0) / tot * 100:5.1f}%  {r[idx['Source']].strip()[:90]}\n")
        ops = collections.Counter()
        for r in sec:
            toks = r[idx["Source"]].split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            ops[op.split(".")[0]] += int(r[idx["Instructions Executed"]] or 0)
        fh.write("   proof-of-path mnemonics executed: " + ", ".join(f"{k}={ops[k]}" for k in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS") if ops.get(k)) + "\n")
    fh.write("\n")


def summarize_launches(path, fh):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[idx["Kernel Name"]].split("(")[0]
        val = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        us = val / 1e3 if unit.startswith("ns") else (val if unit.startswith("us") else val * 1e3)
        agg[name][0] += 1
        agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    fh.write("launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)\n")
    fh.write(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}\n")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"{name[:70]:70s} {n:8d} {us:12.1f} {us / n:10.2f} {us / tot * 100:6.1f}%\n")
    fh.write("\n")


def main():
    """python scripts/summarize_ncu.py <tag> [launches.csv] [report.ncu-rep ...]  (paths under gpurun_out/)"""
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    g = os.path.join(ROOT, "gpurun_out")
    files = sys.argv[2:]
    lc = next((os.path.join(g, f) for f in files if f.endswith(".csv")), os.path.join(g, "launches.csv"))
    reps = [os.path.join(g, f) for f in files if f.endswith(".ncu-rep")] or \
           [os.path.join(g, f) for f in sorted(os.listdir(g)) if f.endswith(".ncu-rep")]
    with open(os.path.join(OUT, f"{tag}_ncu_summary.txt"), "w") as fh:
        fh.write("ncu summaries on one B200.  Launch list: the first launches of `python bench.py --steps 2 --warmup 1 --no-slab`\n"
                 "(config 2: 256x256, batch 64; set-up, warm-up, the timed steps and the per-class profiling passes).  The solver\n"
                 "kernels are launched per batch block -- 2 blocks with the CNN, 4 for bare solver steps, on separate streams -- so\n"
                 "their per-launch figures are per block of 32 / 16 trajectories.  Full captures: scripts/prof_step.py at the\n"
                 "stated shape (conv kernels: 256 x 256 x 64; solver kernels: 1024 x 1024 x 16, one batch block).\n\n")
        if os.path.exists(lc):
            summarize_launches(lc, fh)
        for rep in reps:
            summarize_report(rep, fh)
    print(open(os.path.join(OUT, f"{tag}_ncu_summary.txt")).read()[:9000])


if __name__ == "__main__":
    main()

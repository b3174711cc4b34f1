"""Condense .ncu-rep captures (brought back in gpurun_out/) into small tracked summaries under profiles/.

    python tools/summarise_ncu.py gpurun_out/prof_match_r1c.ncu-rep [more.ncu-rep ...] --tag r1c

Writes profiles/<tag>_<report>.metrics.csv (one row per captured kernel, selected `--page raw` metrics) and
profiles/<tag>_<report>.hotspots.txt (top SASS lines by stall samples, opcode histogram)."""
import argparse
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True, check=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--tag", required=True)
    a = ap.parse_args()
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    for rep in a.reports:
        base = os.path.splitext(os.path.basename(rep))[0]
        rows = list(csv.reader(ncu(["-i", rep, "--page", "raw", "--csv"]).splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        cols = [m for m in METRICS if m in idx]
        with open(os.path.join(out_dir, f"{a.tag}_{base}.metrics.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(cols)
            w.writerow([units[idx[c]] for c in cols])
            for r in rows[2:]:
                w.writerow([r[idx[c]] for c in cols])
        src = list(csv.reader(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]).splitlines()))
        kernels, cur = [], None
        for r in src:
            if r and r[0] == "Kernel Name":
                cur = {"name": r[1], "hdr": None, "rows": []}
                kernels.append(cur)
            elif cur is not None:
                if cur["hdr"] is None:
                    cur["hdr"] = r
                else:
                    cur["rows"].append(r)
        with open(os.path.join(out_dir, f"{a.tag}_{base}.hotspots.txt"), "w") as f:
            for k in kernels:
                h = k["hdr"]
                i_src, i_s, i_ex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
                tot_s = sum(int(r[i_s]) for r in k["rows"]) or 1
                tot_e = sum(int(r[i_ex]) for r in k["rows"]) or 1
                f.write(f"==== {k['name']}\n  SASS lines {len(k['rows'])}, warp instructions executed {tot_e}, stall samples {tot_s}\n")
                f.write("  top lines by stall samples:\n")
                for r in sorted(k["rows"], key=lambda r: -int(r[i_s]))[:15]:
                    f.write(f"    {100 * int(r[i_s]) / tot_s:5.1f}%  exec={r[i_ex]:>9s}  {r[i_src].strip()[:110]}\n")
                hist = collections.Counter()
                for r in k["rows"]:
                    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_src])
                    hist[m.group(2).split(".")[0] if m else "?"] += int(r[i_ex])
                f.write("  opcode mix (executed warp instructions):\n")
                for op, c in hist.most_common(14):
                    f.write(f"    {op:12s} {100 * c / tot_e:5.1f}%\n")
        print("wrote", f"{a.tag}_{base}.metrics.csv / .hotspots.txt")


if __name__ == "__main__":
    sys.exit(main())

"""Summarise an .ncu-rep (read on the build box, no GPU needed) into the handful of numbers the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls N] > profiles/rNN_xxx.txt
"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__cycles_elapsed.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def main():
    rep = sys.argv[1]
    nstall = int(sys.argv[sys.argv.index("--stalls") + 1]) if "--stalls" in sys.argv else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"Kernel Name  {d['Kernel Name'][:110]}")
        for k in KEYS:
            if k in d:
                print(f"  {k:80s} {d[k]} {u.get(k, '')}")
        print()
    if nstall:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(src.splitlines()))
        hdr = rows[1]
        i_src, i_s, i_n = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        data = []
        for r in rows[2:]:
            if r and r[0] == "Kernel Name":
                break
            try:
                data.append((int(r[i_s]), int(r[i_n]), r[i_src]))
            except (ValueError, IndexError):
                pass
        tot = sum(d[0] for d in data)
        print(f"first kernel: {tot} stall samples, {sum(d[1] for d in data)} warp instructions; top {nstall} SASS lines by samples:")
        for s_, n, txt in sorted(data, reverse=True)[:nstall]:
            print(f"  {s_:6d} samples  {n:9d} exec  {txt.strip()[:120]}")


if __name__ == "__main__":
    main()

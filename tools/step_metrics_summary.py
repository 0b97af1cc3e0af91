"""Per-kernel table of one graph replay from the ncu metrics list (tools/ncu_r02.sh step 1): launches, summed time, tensor-pipe
activity and DRAM bytes per kernel -- cold-cache and serialised, so compare SHARES and per-kernel rates, not the sum.

    python tools/step_metrics_summary.py gpurun_out/ncu/r02_step_metrics.csv > profiles/r02_step_kernels.txt
"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr = i
            break
    H = rows[hdr]
    ki, mi, vi, idi, gi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID"), H.index("Grid Size")
    launches = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        d = launches.setdefault(r[idi], {"name": r[ki], "grid": r[gi]})
        try:
            d[r[mi]] = float(r[vi].replace(",", ""))
        except ValueError:
            pass
    ls = list(launches.values())
    packs = [i for i, d in enumerate(ls) if "pack_weights_batched" in d["name"]]
    if len(packs) >= 2:
        ls = ls[packs[0]:packs[1]]
    agg = collections.OrderedDict()
    for d in ls:
        n = re.sub(r"\(CUtensorMap.*|\(.*", "", d["name"]).replace("void stfb::", "").replace("stfb::", "").replace("void ", "")
        n = re.sub(r"at::native::.*", "at::native::*", n)
        a = agg.setdefault(n, {"n": 0, "ns": 0.0, "tens": 0.0, "rd": 0.0, "wr": 0.0, "dram": 0.0})
        t = d.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1
        a["ns"] += t
        a["tens"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        a["dram"] += t * d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["ns"] for a in agg.values())
    print(f"# one graph replay under ncu (--graph-profiling node, cold and serialised): {len(ls)} launches, {tot / 1e6:.3f} ms summed")
    print("#  launches   sum_us  share  tensor_pipe%  dram%   DRAM rd+wr MB   achieved GB/s   kernel")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        if a["ns"] < 3000:
            continue
        mb = (a["rd"] + a["wr"]) / 1e6
        print(f"  {a['n']:5d} {a['ns'] / 1e3:9.1f} {100 * a['ns'] / tot:5.1f}%  {a['tens'] / a['ns']:10.1f}  {a['dram'] / a['ns']:6.1f}  {mb:12.1f}  {mb * 1e6 / a['ns']:12.0f}   {n[:90]}")


if __name__ == "__main__":
    main()

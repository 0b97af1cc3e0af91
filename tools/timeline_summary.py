"""Summarise gpurun_out/timeline.csv (tools/timeline.py): one graph replay, per-stream busy time, concurrency, and the time
each kernel family spends running ALONE (exposed time = what a faster kernel would actually take off the step).

    python tools/timeline_summary.py [gpurun_out/timeline.csv] > profiles/r02_timeline_summary.txt
"""
import collections
import csv
import re
import sys


def short(name):
    n = re.sub(r"\(.*", "", name).replace("void stfb::", "").replace("stfb::", "").replace("void ", "")
    n = re.sub(r"at::native::.*", "at::native::*", n)
    return n[:70]


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.csv"
    rows = [r for r in csv.DictReader(open(path))]
    ev = [(float(r["start_us"]), float(r["dur_us"]), r["stream"], r["name"]) for r in rows]
    packs = [i for i, e in enumerate(ev) if "pack_weights_batched" in e[3]]
    if len(packs) >= 2:
        ev = ev[packs[0]:packs[1]]
    t0 = ev[0][0]
    ev = [(s - t0, d, st, n) for s, d, st, n in ev]
    end = max(s + d for s, d, _, _ in ev)
    print(f"# one step: {len(ev)} device activities, {end / 1e3:.3f} ms from the weight pack to the end of the optimizer")
    busy = collections.defaultdict(float)
    cnt = collections.Counter()
    for s, d, st, n in ev:
        busy[st] += d
        cnt[st] += 1
    print("# per stream: busy ms / launches")
    for st, b in sorted(busy.items(), key=lambda kv: -kv[1]):
        print(f"  stream {st:>6}: {b / 1e3:7.3f} ms  {cnt[st]:4d}")
    # sweep: concurrency profile and exclusive time per kernel
    pts = []
    for i, (s, d, st, n) in enumerate(ev):
        pts.append((s, 1, i))
        pts.append((s + d, -1, i))
    pts.sort(key=lambda p: (p[0], p[1]))
    active = set()
    last = 0.0
    conc = collections.defaultdict(float)
    excl = collections.defaultdict(float)
    shared = collections.defaultdict(float)
    for t, k, i in pts:
        dt = t - last
        if dt > 0:
            conc[min(len(active), 4)] += dt
            if len(active) == 1:
                excl[short(ev[next(iter(active))][3])] += dt
            elif len(active) > 1:
                for j in active:
                    shared[short(ev[j][3])] += dt / len(active)
        last = t
        if k == 1:
            active.add(i)
        else:
            active.discard(i)
    print("# concurrency (ms with k kernels in flight): " + ", ".join(f"{k}{'+' if k == 4 else ''}: {v / 1e3:.3f}" for k, v in sorted(conc.items())))
    tot = collections.defaultdict(float)
    n_l = collections.Counter()
    for s, d, st, n in ev:
        tot[short(n)] += d
        n_l[short(n)] += 1
    print("# kernel: launches, summed duration, time running ALONE, its share of concurrent time (ms)")
    for k, v in sorted(tot.items(), key=lambda kv: -(excl[kv[0]] + shared[kv[0]])):
        if v < 20:
            continue
        print(f"  {n_l[k]:4d} {v / 1e3:7.3f} {excl[k] / 1e3:7.3f} {shared[k] / 1e3:7.3f}  {k}")


if __name__ == "__main__":
    main()

"""Per-kernel counts of the Blackwell-native SASS mnemonics in libstfb200.so (cuobjdump -sass; no GPU needed).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG = TMA tensor loads, UTCBAR = tcgen05.commit, UTCATOMSWS =
tcgen05.alloc/dealloc, SYNCS = mbarrier ops, UCGABAR = cluster barriers; HMMA (legacy mma.sync) must not appear.
"""
import collections
import os
import re
import subprocess
import sys

SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stf_unet_b200", "libstfb200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTCBAR", "UTCBAR.2CTA", "UTCATOMSWS", "SYNCS", "UCGABAR", "HMMA", "FFMA", "MUFU", "RED", "ATOM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], stdout=subprocess.PIPE, text=True, check=True).stdout
    per = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            name = re.sub(r"\(CUtensorMap_st.*|\(.*", "", name).replace("void stfb::", "").replace("stfb::", "")
            per[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        c = per[name]
        c["_total"] += 1
        base = op.split(".")[0]
        if base in ("UTCHMMA", "UTCBAR") and ".2CTA" in op:
            c[base + ".2CTA"] += 1
        for pref in ("UCGABAR", "RED", "ATOM"):              # UCGABAR_ARV / UCGABAR_WAIT, REDG / REDUX, ATOMG / ATOMS ...
            if base.startswith(pref) and base != pref:
                c[pref] += 1
        if base in KEYS:
            c[base] += 1
    print(f"# {os.path.basename(SO)}: per-kernel SASS mnemonic counts (cuobjdump -sass, sm_100a)")
    print("# " + " ".join(f"{k:>12s}" for k in KEYS) + "  total  kernel")
    tot = collections.Counter()
    for k, c in sorted(per.items(), key=lambda kv: -kv[1]["UTCHMMA"]):
        if not any(c[x] for x in ("UTCHMMA", "LDTM", "UTMALDG", "UCGABAR")):
            continue
        print("  " + " ".join(f"{c[x]:12d}" for x in KEYS) + f" {c['_total']:6d}  {k}")
        tot.update(c)
    print("# kernels without tensor / TMA instructions (FFMA, MUFU, RED/ATOM counts):")
    for k, c in sorted(per.items()):
        if any(c[x] for x in ("UTCHMMA", "LDTM", "UTMALDG", "UCGABAR")):
            continue
        print(f"  FFMA {c['FFMA']:5d} MUFU {c['MUFU']:4d} RED {c['RED']:3d} ATOM {c['ATOM']:3d} total {c['_total']:6d}  {k}")
        tot.update(c)
    print("# library totals: " + ", ".join(f"{x}={tot[x]}" for x in KEYS))
    assert tot["HMMA"] == 0, "legacy mma.sync path found"


if __name__ == "__main__":
    main()

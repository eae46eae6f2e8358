"""SASS evidence for the Blackwell-native kernels: per kernel, how many tcgen05 / TMA instructions the in-tree
libgraphslim_b200.so contains (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor, UBLKCP =
cp.async.bulk, UTCBAR = tcgen05.commit), followed by the instruction lines themselves for the fused PGE kernels.

    python profiles/sass_evidence.py > profiles/r2_sass_evidence.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "graphslim_b200", "libgraphslim_b200.so")
KEYS = ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "HMMA")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    counts, lines, fn = collections.OrderedDict(), collections.defaultdict(list), None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        for k in KEYS:
            if re.search(r"\b" + k + r"\b", ln) or (k in ln and k != "HMMA"):
                counts[fn][k] += 1
                lines[fn].append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln.strip()))
                break
    print("# tcgen05 / TMA instruction counts per kernel (cuobjdump -sass graphslim_b200/libgraphslim_b200.so)\n")
    tot = collections.Counter()
    for fn, c in counts.items():
        if not c:
            continue
        tot.update(c)
        print(f"{demangle(fn)[:110]:110s} " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
    print("\nTOTAL " + "  ".join(f"{k}={v}" for k, v in sorted(tot.items())))
    for fn in counts:
        if "pge_l2" in fn and "ILi256ELi3" in fn:
            print(f"\n## {demangle(fn)}")
            for ln in lines[fn]:
                print("   " + ln)


if __name__ == "__main__":
    main()

"""Instruction histogram of the in-tree library's SASS (cuobjdump -sass), per kernel family: the mnemonics that prove the
hot path is hand-written tcgen05 / TMEM / TMA code (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor
loads, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops) and what the aux kernels are made of.  Runs on CPU.
usage: python tools/sass_histogram.py > profiles/r02_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "v2x-sim_b200", "v2x_b200", "libv2x_b200.so")
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "HMMA", "FFMA",
       "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "REDG", "MUFU", "BAR", "ELECT", "ACQBULK", "MEMBAR", "ERRBAR", "NANOSLEEP"]


def family(name):
    name = re.sub(r"\(.*", "", name)
    m = re.search(r"v2x::(\w+)", name) or re.search(r"(\w+_kernel)", name)
    return m.group(1) if m else name[:60]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur][op] += 1
            kernels[cur]["__total__"] += 1
    names = list(kernels)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    fam = collections.OrderedDict()
    for mangled, d in zip(names, dem):
        f = family(d)
        agg = fam.setdefault(f, {"variants": 0, "ops": collections.Counter()})
        agg["variants"] += 1
        agg["ops"].update(kernels[mangled])
    print("# SASS instruction histogram of libv2x_b200.so (`cuobjdump -sass`, sm_100a)\n")
    print("Per kernel family (template instantiations summed).  UTCHMMA = `tcgen05.mma`, LDTM = `tcgen05.ld`, UTMALDG = TMA tensor "
          "loads, UTCBAR = `tcgen05.commit`, SYNCS = mbarrier arrive / try_wait, LDGSTS = `cp.async`, MEMBAR = fences "
          "(`fence.sys` of the peer exchange), ATOMG / RED = global atomics.\n")
    cols = [k for k in KEY if any(a["ops"].get(k, 0) or any(o.startswith(k) for o in a["ops"]) for a in fam.values())]
    print("| kernel family | variants | SASS instructions | " + " | ".join(cols) + " |")
    print("|---|---|---|" + "---|" * len(cols))
    tot = collections.Counter()
    for f, a in fam.items():
        def cnt(k):
            return sum(v for o, v in a["ops"].items() if o == k or o.startswith(k + "_") or (k in ("LDG", "STG", "LDS", "STS") and o == k))
        row = [cnt(k) for k in cols]
        for k, v in zip(cols, row):
            tot[k] += v
        print("| `%s` | %d | %d | " % (f, a["variants"], a["ops"]["__total__"]) + " | ".join(str(v) if v else "" for v in row) + " |")
    print("| **total** | %d | %d | " % (sum(a["variants"] for a in fam.values()), sum(a["ops"]["__total__"] for a in fam.values()))
          + " | ".join(str(tot[k]) for k in cols) + " |")


if __name__ == "__main__":
    main()

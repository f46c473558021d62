"""Aggregates an `ncu --csv` log (long format: one row per kernel launch and metric) by kernel family into a table of launch
counts, time, DRAM and L2 throughput against the measured peaks.
   python tools/aux_report.py gpurun_out/aux.csv profiles/r02_aux_kernels.md"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def family(name):
    name = re.sub(r"\(.*", "", name)
    m = re.search(r"(\w+_kernel)", name)
    base = m.group(1) if m else name[:50]
    t = re.search(r"_kernel<([^>]*)>", name)
    return base + ("<%s>" % t.group(1) if t and base.startswith(("conv_wgrad", "channel_reduce", "resample2", "nms_map")) else "")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm = float(peaks["hbm_gbs"])
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    head = rows[0]
    ix = {h: i for i, h in enumerate(head)}
    per = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) < len(head):
            continue
        k = (r[ix["ID"]], r[ix["Kernel Name"]])
        v = r[ix["Metric Value"]].replace(",", "")
        try:
            val = float(v)
        except ValueError:
            continue
        unit = r[ix["Metric Unit"]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per.setdefault(k, {})[r[ix["Metric Name"]]] = val * scale
    fam = collections.OrderedDict()
    for (_, name), m in per.items():
        f = fam.setdefault(family(name), {"n": 0, "us": 0.0, "dram": 0.0, "l2": 0.0})
        f["n"] += 1
        f["us"] += m.get("gpu__time_duration.sum", 0.0)
        f["dram"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        f["l2"] += m.get("lts__t_bytes.sum", 0.0)
    tot = sum(f["us"] for f in fam.values())
    out = ["# Kernels outside the bench configs' launch lists (tools/prof_aux.py under ncu, %s)" % os.path.basename(src), "",
           "One pass each of: DiscoNet / AgentWise / Max fusion forward (8 scenes), uint8-input V2VNet forward + on-device NMS, "
           "voxel-row forward, and one V2VNet training step (4 scenes, forward + backward).  Durations are ncu's (serialised, "
           "cold clocks); GB/s = bytes / duration summed over the family's launches; HBM peak %.0f GB/s (MEASURED_PEAKS.json); "
           "L2 = lts__t_bytes (the LTS cap is ~12.4 TB/s at 1965 MHz)." % hbm, "",
           "| kernel family | launches | total us | share | DRAM MB | DRAM GB/s | % HBM peak | L2 MB | L2 GB/s |",
           "|---|---|---|---|---|---|---|---|---|"]
    for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        us = max(f["us"], 1e-9)
        out.append("| `%s` | %d | %.1f | %.1f%% | %.1f | %.0f | %.0f%% | %.1f | %.0f |" % (
            name, f["n"], f["us"], 100 * f["us"] / tot, f["dram"] / 1e6, f["dram"] / us / 1e3, 100 * f["dram"] / us / 1e3 / hbm,
            f["l2"] / 1e6, f["l2"] / us / 1e3))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:14]))


if __name__ == "__main__":
    main()

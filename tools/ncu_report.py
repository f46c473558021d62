"""Join an `ncu --set full` capture of tools/prof_step.py with the plan's launch list -> per-launch roofline table.
   python tools/ncu_report.py gpurun_out/prof_v2v_det_mixed.ncu-rep gpurun_out/prof_v2v_det_mixed_launches.json profiles/r02_v2v_det_mixed
Writes <prefix>_ncu_full_summary.csv and <prefix>_per_layer_roofline.md and records the conv launches' DRAM bytes in
profiles/conv_traffic.json (bench.py's roofline.traffic).  Reads the report here (no GPU needed).

Columns of the table: executed TFLOP/s of the launch (its multiply-accumulates x tensor-core passes = what the tensor
pipe sustains) against the measured burst bf16 peak, DRAM GB/s (dram__bytes_read + write over the launch duration)
against the measured copy bandwidth, and which of the two the launch is nearer to.  ncu durations are serialised /
cold-cache, so absolute times run a few percent slower than the live CUDA-event timing of bench.py; the per-launch SHARE
is what must agree (profiles/README.md)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3,
         "msecond": 1e3}


def num(v, unit):
    return float(v.replace(",", "")) * SCALE.get(unit, 1)


def main():
    rep, launches_json, prefix = sys.argv[1:4]
    meta = json.load(open(launches_json))
    # a .ncu-rep of a whole step is 60-100 MB (over gpurun's 64 MiB return limit): tools/gpu_r02.sh exports the raw page
    # to CSV on the GPU box and deletes the report, so this tool normally reads the CSV
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(head)}
    cols = [c for c in COLS if c in ix]
    # a python-side launch may issue several kernels (voxelize: memset + scatter) or none; join conv launches in order
    convs = [l for l in meta["launches"] if l["flops"] > 0]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    tf_peak, hbm_peak = float(peaks["bf16_tflops"]), float(peaks["hbm_gbs"])
    out = [["layer", "ID", "Kernel Name", "Grid Size", "Block Size"] + cols, ["", "", "", "", ""] + [units[ix[c]] for c in cols]]
    md = ["# Per-launch roofline: %s, %s, %d units per step (%s)" % (meta["config"], meta["precision"], meta["units"],
                                                                       os.path.basename(rep)), "",
          "Peaks (MEASURED_PEAKS.json): %.1f TFLOP/s burst bf16 (dense, cuBLAS), %.1f GB/s HBM copy.  `pipe TFLOP/s` = the "
          "launch's multiply-accumulates x tensor-core passes per k-step / its duration." % (tf_peak, hbm_peak), "",
          "| # | kernel | launch | us | GFLOP (1 pass) | passes | pipe TFLOP/s | % tensor peak | DRAM MB rd+wr | GB/s | % HBM peak | tensor pipe active % | regs | nearest bound |",
          "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    ci = 0
    tot_t = tot_conv_t = rd_b = wr_b = 0.0
    for i, r in enumerate(data):
        kname = r[ix["Kernel Name"]]
        short = kname.split("(")[0].replace("void ", "").replace("v2x::", "")[:46]
        is_conv = "conv_tc" in kname or "conv_pack3" in kname
        t = num(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
        rd = num(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = num(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        label, gf, passes = "", 0.0, 0
        if is_conv and ci < len(convs):
            label, gf, passes = convs[ci]["label"], convs[ci]["flops"] / 1e9, convs[ci]["passes"]
            ci += 1
            rd_b += rd
            wr_b += wr
            tot_conv_t += t
        tot_t += t
        tfs = (gf * passes * 1e9) / (t * 1e-6) / 1e12 if t else 0.0
        gbs = (rd + wr) / (t * 1e-6) / 1e9 if t else 0.0
        pt, ph = 100 * tfs / tf_peak, 100 * gbs / hbm_peak
        out.append([label or short, r[ix["ID"]], kname[:90], r[ix["Grid Size"]], r[ix["Block Size"]]] + [r[ix[c]] for c in cols])
        md.append("| %d | %s | %s | %.1f | %.1f | %d | %.0f | %.0f%% | %.0f | %.0f | %.0f%% | %s | %s | %s |" % (
            i, short, label, t, gf, passes, tfs, pt, (rd + wr) / 1e6, gbs, ph,
            r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]][:5] if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in ix else "",
            r[ix["launch__registers_per_thread"]] if "launch__registers_per_thread" in ix else "",
            "tensor" if pt >= ph else "HBM"))
    md += ["", "Step under ncu: %.0f us, %.0f us (%.0f%%) in the %d conv launches; conv DRAM traffic %.3f GB read + %.3f GB written."
           % (tot_t, tot_conv_t, 100 * tot_conv_t / max(tot_t, 1e-9), ci, rd_b / 1e9, wr_b / 1e9)]
    with open(prefix + "_ncu_full_summary.csv", "w", newline="") as f:
        csv.writer(f).writerows(out)
    with open(prefix + "_per_layer_roofline.md", "w") as f:
        f.write("\n".join(md) + "\n")
    tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")
    try:
        tj = json.load(open(tpath))
        caps = tj.get("captures") or [dict(tj, config="v2v_det", precision="bf16")]
    except Exception:
        caps = []
    caps = [c for c in caps if not (c.get("config") == meta["config"] and c.get("precision") == meta["precision"]
                                    and int(c.get("scenes_per_step", -1)) == meta["units"])]
    caps.append({"config": meta["config"], "precision": meta["precision"], "scenes_per_step": meta["units"],
                 "conv_dram_bytes_per_step": rd_b + wr_b, "conv_dram_read": rd_b, "conv_dram_write": wr_b,
                 "source": os.path.basename(prefix) + "_ncu_full_summary.csv (ncu --set full, one eager step)"})
    json.dump({"captures": caps}, open(tpath, "w"), indent=1)
    print("\n".join(md[-2:]))


if __name__ == "__main__":
    main()

"""Per-launch roofline table from an ncu full summary (tools/ncu_summary.py output):
   python tools/roofline_report.py profiles/r01_v10_ncu_full_summary.csv profiles/r01_v10_per_layer_roofline.md
Algorithmic FLOPs per launch = SURVEY.md section 8(a) per-layer table (A = 5 agents) x scenes per step; peaks from
MEASURED_PEAKS.json (sustained bf16 TFLOP/s, copy GB/s).  ncu durations are cold-clock / serialised, so the absolute
numbers are a few percent slower than the live CUDA-event timing in bench.py; the per-layer picture is what matters."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# GFLOP per frame (5 agents); gru_m / gru_h split of the 108.72 GFLOP W_ih work as executed here:
# mean half once per frame (36.24), own-state half in each of the 3 rounds (3 x 12.08 + identity columns not counted)
GF = {"pre_1": 2.454, "pre_2": 6.040, "c1_1": 3.020, "c1_2": 6.040, "c3d_1": 0.671, "c2_1": 3.020, "c2_2": 6.040,
      "c3d_2": 0.671, "c3_1": 3.020, "c3_2": 6.040, "c4_1": 3.020, "c4_2": 6.040, "gru_m": 18.12, "gru1": 18.12,
      "gru2": 18.12, "gru3": 18.12, "c5_1": 18.119, "c5_2": 6.040, "c6_1": 18.119, "c6_2": 6.040, "c7_1": 18.119,
      "c7_2": 6.040, "c8_1": 18.119, "c8_2": 6.040, "heads": 13.087}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    scenes = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        peaks = json.load(f)
    tf_peak, hbm_peak = float(peaks["bf16_tflops_sustained"]), float(peaks["hbm_gbs"])
    rows = list(csv.reader(open(src)))
    head = rows[0]
    ix = {h: i for i, h in enumerate(head)}
    out = ["# Per-launch roofline, one 8-scene step (%s)" % os.path.basename(src), "",
           "Peaks: %.1f TFLOP/s sustained bf16, %.1f GB/s HBM copy (MEASURED_PEAKS.json).  GFLOP = algorithmic conv FLOPs "
           "of the launch (SURVEY 8(a) x %d scenes); the 3 GRU rounds + the hoisted mean half execute 72.5 of the 108.7 "
           "GFLOP/frame the W_ih convs cost when recomputed every round." % (tf_peak, hbm_peak, scenes), "",
           "| launch | grid | us | GFLOP | TFLOP/s | % tensor peak | DRAM MB (rd+wr) | GB/s | % HBM peak | tensor pipe active % | nearest bound |",
           "|---|---|---|---|---|---|---|---|---|---|---|"]
    tot_t = tot_gf = 0.0
    for r in rows[2:]:
        name = r[0]
        t = float(r[ix["gpu__time_duration.sum"]])
        mb = float(r[ix["dram__bytes_read.sum"]]) + float(r[ix["dram__bytes_write.sum"]])
        gf = GF.get(name, 0.0) * scenes
        tfs = gf / t * 1e3 if t else 0.0           # GFLOP / us = 1000 TFLOP/s
        gbs = mb / t * 1e3 if t else 0.0           # MB / us = 1000 GB/s
        pt, ph = 100 * tfs / tf_peak, 100 * gbs / hbm_peak
        bound = "tensor" if pt >= ph else "HBM"
        out.append("| %s | %s | %.1f | %.1f | %.0f | %.0f%% | %.0f | %.0f | %.0f%% | %s | %s |" % (
            name, r[ix["Grid Size"]].replace(" ", ""), t, gf, tfs, pt, mb, gbs, ph,
            r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]][:5], bound))
        tot_t += t
        tot_gf += gf
    out += ["", "Step: %.0f us under ncu, %.1f GFLOP algorithmic as executed -> %.0f TFLOP/s over the whole step "
            "(264.51 GFLOP/frame x %d = %.0f GFLOP of SURVEY-algorithmic work -> %.0f TFLOP/s)." % (
                tot_t, tot_gf, tot_gf / tot_t * 1e3, scenes, 264.51 * scenes, 264.51 * scenes / tot_t * 1e3)]
    with open(dst, "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out[-3:]))


if __name__ == "__main__":
    main()

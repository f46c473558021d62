"""Summarise an `ncu --set full` capture of one eager bench step into profiles/:
   python tools/ncu_summary.py gpurun_out/prof_step.ncu-rep profiles/r01_v9   (reads the report here; no GPU needed)
Writes <prefix>_ncu_full_summary.csv (one row per launch: duration, DRAM bytes, tensor-pipe %, ...) and refreshes
profiles/conv_traffic.json (DRAM read+write summed over the conv launches, which bench.py reports as roofline.traffic)."""
import csv
import io
import json
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
NAMES = ["pack_in", "pre_1", "pre_2", "c1_1", "c1_2", "c3d_1", "c2_1", "c2_2", "c3d_2", "c3_1", "c3_2", "c4_1", "c4_2", "warp",
         "gru_m", "gru1", "gru2", "gru3", "c5_1", "c5_2", "c6_1", "c6_2", "c7_1", "c7_2", "c8_1", "c8_2", "heads"]


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * scale.get(unit, 1)


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    scenes = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(head)}
    cols = [c for c in COLS if c in ix]
    out = [["layer", "ID", "Kernel Name", "Grid Size", "Block Size"] + cols,
           ["", "", "", "", ""] + [units[ix[c]] for c in cols]]
    rd = wr = 0.0
    for i, r in enumerate(data):
        name = NAMES[i] if len(data) == len(NAMES) else str(i)
        out.append([name, r[ix["ID"]], r[ix["Kernel Name"]][:80], r[ix["Grid Size"]], r[ix["Block Size"]]] + [r[ix[c]] for c in cols])
        if "conv_tc" in r[ix["Kernel Name"]] or "conv_pack3" in r[ix["Kernel Name"]]:
            rd += to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
            wr += to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    path = prefix + "_ncu_full_summary.csv"
    with open(path, "w", newline="") as f:
        csv.writer(f).writerows(out)
    with open("profiles/conv_traffic.json", "w") as f:
        json.dump({"source": "%s (ncu --set full, one eager step, %d scenes x 5 agents, bf16)" % (path, scenes),
                   "scenes_per_step": scenes, "conv_dram_bytes_per_step": rd + wr, "conv_dram_read": rd,
                   "conv_dram_write": wr}, f)
    print("wrote", path, "conv DRAM bytes/step %.3e over %d launches" % (rd + wr, len(data)))


if __name__ == "__main__":
    main()

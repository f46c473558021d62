"""bench.py's one-line JSON contract (the driver parses it): the reference arm on CPU, the measured arm on the GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "bench.py must print exactly ONE line on stdout, got %d" % len(lines)
    return json.loads(lines[0])


def test_reference_arm_line():
    """`bench.py --impl reference` times the LIVE reference module on the host's cores (from /root/reference here, from the
    copy staged under oracle/_ref on the GPU box; the oracle port only if neither exists) and prints the same schema with
    impl = reference, zero transfer bytes and a cpu_baseline describing the run."""
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"], 600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    from oracle import ref_loader
    if ref_loader.available():
        assert cb["kind"] == "reference" and cb["reference_s_per_unit"] > 0 and cb["port_s_per_unit"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_measured_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--no-cpu-baseline"], 900)
    assert (BASE_KEYS | {"roofline", "clocks", "gpu_launches", "e2e_detections"}) <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["dtype"].startswith("fp16 hi/lo") and d["data"] == "synthetic"
    assert d["config"]["precision"] == "mixed"      # the default = the mode the 1e-3 parity tests assert
    assert d["value"] > 100 and d["gpu_launches"] == d["kernels_per_step"] * 3 > 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["frac_executed"] < r["frac"] < r["frac_tensor_pipe"] < 1   # V2VNet hoists work (executed < algorithmic); mixed runs 1-3 passes
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 8 * 5 * 256 * 256 * 13 * 4 + 8 * 25 * 16 * 8 + 8 * 5 * 8
    assert e["d2h_bytes_per_step"] == 8 * 5 * 256 * 256 * 6 * (2 + 6) * 4 and 0 < e["value"] < d["value"]
    assert d["e2e_detections"]["d2h_bytes_per_step"] < e["d2h_bytes_per_step"] // 100
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


@pytest.mark.parametrize("config,unit", [("faf_lower", "agent-frames/s"), ("w2c_seg", "frames/s"), ("faf_upper_dp", "frames/s"),
                                         ("w2c_det", "frames/s")])
def test_reference_arm_other_configs(config, unit):
    d = _run(["--impl", "reference", "--config", config, "--steps", "1", "--warmup", "1"], 900)
    assert d["impl"] == "reference" and d["unit"] == unit and d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"]


@pytest.mark.gpu
@pytest.mark.parametrize("config", ["faf_lower", "w2c_seg", "faf_upper_dp", "w2c_det"])
def test_measured_arm_other_configs(config):
    d = _run(["--config", config, "--steps", "2", "--warmup", "3", "--no-cpu-baseline"], 900)
    assert (BASE_KEYS | {"roofline", "clocks", "gpu_launches"}) <= set(d)
    assert d["value"] > 0 and 0 < d["e2e"]["value"] < d["value"] and d["e2e"]["h2d_bytes_per_step"] > 0
    assert 0 < d["roofline"]["frac"] < 1 and d["gpu_launches"] == d["kernels_per_step"] * 2


def test_our_arm_executes_end_to_end_under_the_cpu_dry_run():
    """tools/bench_dryrun.py: bench.py's own arm with kernels as no-ops.  Every key of the bench contract must be on the
    one JSON line it prints (values are meaningless here; the B200 lines are under profiles/)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "bench_dryrun.py"), "--steps", "2", "--warmup", "1",
                        "--scenes", "1", "--no-cpu-baseline"], cwd=root, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["warmup"] >= 3 and d["config"]["workload"] and "model" not in d["config"]
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and d["gpu_launches"] == 27 * 2


@pytest.mark.parametrize("extra", [[], ["--config", "w2c_det"]], ids=["v2v_det", "w2c_det"])
def test_two_rank_launch_of_our_arm_under_the_cpu_dry_run(extra):
    """The driver's N > 1 launch (torch.distributed.run, one rank per GPU) with kernels as no-ops and gloo in place of
    NCCL: unit-sharded plan, per-step exchange, graph capture, barrier + max-over-ranks timing, ONE line from rank 0."""
    import json
    import socket
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tools", "bench_dryrun.py"), "--gpus", "2", "--steps", "2",
           "--warmup", "1", "--scenes", "2", "--no-cpu-baseline"] + extra
    r = subprocess.run(cmd, cwd=root, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and "e2e" in d and "roofline" in d

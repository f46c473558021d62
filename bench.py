#!/usr/bin/env python
"""bench.py -- frames/sec of the 5-agent V2VNet detection forward on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our sm_100a path (one rank per GPU)
  python bench.py --impl reference ...                      the reference CPU path (oracle port) on host cores

A *frame* is one scene: all 5 agents' 256x256x13 BEVs -> all agents' loc/cls (BASELINE.json,
SURVEY.md section 8(d)).  A *step* is one forward over ``--scenes`` scenes per GPU (default 8, so the
step's inputs, 8 x 17 MB fp32, exceed the 126 MB L2 and nothing survives between timed steps).

  value  frames/s with inputs resident in HBM, CUDA-graph replay, CUDA events on the launch stream
  e2e    frames/s through the drop-in ``coperception.models.det.V2VNet.forward`` with pinned HOST
         inputs: H2D of bevs/trans/num_agent and D2H of loc+cls (fp32, as the reference's predict_all
         moves them, CoDetModule.py:484-511 -> detection_util.py:256) inside the timed region
  roofline   the conv implicit-GEMM kernel family: algorithmic conv FLOPs / summed live per-launch time
  cpu_baseline  the oracle port (the reference's CPU algorithm) on this host, bounded sample

Multi-GPU (N > 1), weak scaling, 8 scenes x 5 agents = 40 units per GPU, timing = max over ranks of the device
time between barriers:
  --shard unit  (default)  the 40*N agent-major units are split contiguously over the ranks, so a scene's agents live
                on different GPUs; the only data-path collective is one NCCL all-gather of the layer-3 maps per
                step, overlapped with the x_4 encoder branch (BASELINE configs[3])
  --shard scene            every rank keeps whole scenes; no data-path collective at all
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "v2x-sim_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

AGENTS = 5
GNN_ITER = 3
# algorithmic conv/linear FLOPs per frame (A=5), SURVEY.md section 8(d) / BASELINE.md section 2
GFLOP_PER_FRAME = 264.51
METRIC = "frames/sec 5-agent V2VNet fwd"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1450.0))),
                "tflops_burst": float(d.get("bf16_tflops", 0.0)), "hbm_gbs": float(d.get("hbm_gbs", 0.0)),
                "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md: 1.59 PF burst / ~1.4 PF sustained)"}


def conv_traffic(scenes):
    """dram__bytes_read.sum + dram__bytes_write.sum summed over the conv launches of one step, from the committed
    ncu --set full capture (profiles/conv_traffic.json); None when the capture was taken at another batch size."""
    path = os.path.join(ROOT, "profiles", "conv_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["conv_dram_bytes_per_step"]) if int(d["scenes_per_step"]) == scenes else None
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _anchor_table():
    """The detection anchors [256*256*6, 6] = (x, y, w, h, sin, cos) of the reference's default config
    (CP/utils/obj_util.py:611-633 init_anchors_no_check with Config.py:76-84,154-163), built on the host in numpy."""
    import math
    import numpy as np
    sizes = np.asarray([[2.0, 4.0, 0.0], [2.0, 4.0, math.pi / 2.0], [2.0, 4.0, -math.pi / 4.0],
                        [3.0, 12.0, 0.0], [3.0, 12.0, math.pi / 2.0], [3.0, 12.0, -math.pi / 4.0]])
    a = np.zeros((256, 256, 6, 6))
    a[..., 2:4] = sizes[:, :2]
    a[..., 4], a[..., 5] = np.sin(sizes[:, 2]), np.cos(sizes[:, 2])
    centre = np.arange(256) * 0.25 - 32.0 + 0.125
    a[..., 0] = centre[None, :, None]
    a[..., 1] = centre[:, None, None]
    return a.reshape(-1, 6).astype(np.float32)


def cpu_reference(steps, warmup, scenes=1):
    """The reference's CPU algorithm (oracle port, literal restatement incl. the W_hh conv over the zero
    hidden state and per-round warps) on this host's cores.  torch's CPU backend gets slower, not faster,
    when handed every core of a big host for these small convolutions, so the thread count is chosen
    from {8, 16, 32, 64, all} by one probe forward each and the best one is used and reported.
    Returns (frames/s, seconds per frame, threads)."""
    from oracle import restate, synth
    sd = synth.v2vnet_det_state(0)
    bevs, trans, nat = synth.make_scene(scenes, AGENTS, seed=0)

    def fwd():
        t0 = time.perf_counter()
        with torch.no_grad():
            restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=scenes, agent_num=AGENTS, gnn_iter=GNN_ITER)
        return time.perf_counter() - t0

    ncpu = os.cpu_count() or 1
    best_t, best_n = None, None
    for n in sorted({min(c, ncpu) for c in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(n)
        fwd()
        dt = fwd()
        if best_t is None or dt < best_t:
            best_t, best_n = dt, n
        if dt > 20.0:
            break
    torch.set_num_threads(best_n)
    times = [fwd() for _ in range(warmup + steps)][warmup:]
    times.sort()
    med = times[len(times) // 2]
    return scenes / med, med / scenes, best_n


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    fps, spf, threads = cpu_reference(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": spf * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "V2VNet 5-agent detection fwd, 256x256x13 BEV, 1 scene/step, reference CPU path"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": "%d timed forwards of 1 scene (5 agents), median; torch CPU fp32, best of "
                                       "{8,16,32,64,all} threads = %d" % (steps, threads)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def time_launch_list(launches, reps=20):
    """Device time (ms) of one pass over `launches`: the list is captured into a CUDA graph (like the timed step, so
    no host launch gaps leak in) and replayed `reps` times between CUDA events on the replay stream."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for l in launches:
            l()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for l in launches:
            l()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_ours(args, rank, world, local_rank):
    from v2x_b200 import default_det_config, nets
    from v2x_b200 import synthetic as synth   # seeded synthetic weights / inputs (nothing under oracle/ on this arm)
    from coperception.models.det import V2VNet
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    planes = {"bf16": 1, "bf16x3": 2}[args.precision]
    B = args.scenes
    sd = synth.v2vnet_det_state(0)
    bevs, trans, nat = synth.make_scene(B, AGENTS, seed=rank)
    # Planted head biases (synthetic-data helper): with plain random weights about half of all anchors pass
    # the reference's 0.7 score filter, which no trained detector does; shift the foreground bias -- from THIS path's
    # own logits on one scene -- so that ~150 anchors per agent do.  Backbone / fusion weights and every kernel's work
    # are unchanged; only the post-processing leg (e2e_detections) depends on it.
    probe = nets.V2VNetDetPlan(sd, 1, AGENTS, gnn_iter=GNN_ITER, planes=planes, device=dev)
    cls0 = probe.forward(bevs[::B].to(dev), trans[:1].to(dev), nat[:1].to(dev))["cls"].float().cpu()
    sd = synth.plant_detections(sd, cls0, per_agent=150)
    del probe, cls0
    torch.cuda.empty_cache()

    def barrier():
        if world > 1:
            dist.barrier()

    # ---------------- value: device-resident inputs, graph replay ----------------
    unit_sharded = world > 1 and args.shard == "unit"
    if unit_sharded:
        from v2x_b200 import sharding
        gb, gt, gn = synth.make_scene(B * world, AGENTS, seed=0)      # the global batch: B * world scenes
        off, n_loc = sharding.unit_range(B * world * AGENTS, rank, world)
        plan = nets.V2VNetDetShardedPlan(sd, B * world, AGENTS, rank, world, gnn_iter=GNN_ITER, planes=planes, device=dev,
                                         exchange=args.exchange)
        plan.set_inputs(gb[off:off + n_loc].to(dev), gt.to(dev), gn.to(dev))
        del gb
    else:
        plan = nets.V2VNetDetPlan(sd, B, AGENTS, gnn_iter=GNN_ITER, planes=planes, device=dev)
        plan.set_inputs(bevs.to(dev), trans.to(dev), nat.to(dev))
    torch.cuda.synchronize()
    plan.capture()
    for _ in range(args.warmup):
        plan.run()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    frames = B * world * args.steps
    value = frames / (ms_total * 1e-3)

    # ---------------- e2e: drop-in module, pinned host buffers, H2D + D2H in the timed region ----------------
    model = V2VNet(default_det_config(), GNN_ITER, 3, 256, num_agent=AGENTS)
    model.load_state_dict(sd, strict=True)
    model.precision = args.precision
    model = model.to(dev).eval()
    NBUF = 2
    h_bev = [bevs.clone().pin_memory() for _ in range(NBUF)]
    h_trans = [trans.clone().pin_memory() for _ in range(NBUF)]
    h_nat = [nat.clone().pin_memory() for _ in range(NBUF)]
    d_bev = [torch.empty_like(bevs, device=dev) for _ in range(NBUF)]
    d_trans = [torch.empty_like(trans, device=dev) for _ in range(NBUF)]
    d_nat = [torch.empty_like(nat, device=dev) for _ in range(NBUF)]
    n_maps = B * AGENTS
    h_loc = [torch.empty((n_maps, 256, 256, 6, 1, 6), dtype=torch.float32).pin_memory() for _ in range(NBUF)]
    h_cls = [torch.empty((n_maps, 256 * 256 * 6, 2), dtype=torch.float32).pin_memory() for _ in range(NBUF)]
    d_loc = [torch.empty((n_maps, 256, 256, 6, 1, 6), dtype=torch.float32, device=dev) for _ in range(NBUF)]
    d_cls = [torch.empty((n_maps, 256 * 256 * 6, 2), dtype=torch.float32, device=dev) for _ in range(NBUF)]
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    main = torch.cuda.current_stream()
    in_ready = [torch.cuda.Event() for _ in range(NBUF)]
    in_free = [torch.cuda.Event() for _ in range(NBUF)]
    out_ready = [torch.cuda.Event() for _ in range(NBUF)]
    out_free = [torch.cuda.Event() for _ in range(NBUF)]

    def e2e_steps(k):
        """Software-pipelined: H2D of step i+1 and D2H of step i-1 overlap the forward of step i
        (three streams); every step's inputs come from host memory and its outputs land in host memory."""
        for i in range(k):
            b = i % NBUF
            with torch.cuda.stream(s_in):
                s_in.wait_event(in_free[b])
                d_bev[b].copy_(h_bev[b], non_blocking=True)
                d_trans[b].copy_(h_trans[b], non_blocking=True)
                d_nat[b].copy_(h_nat[b], non_blocking=True)
                in_ready[b].record(s_in)
            main.wait_event(in_ready[b])
            with torch.no_grad():
                out = model(d_bev[b], d_trans[b], d_nat[b], batch_size=B)
            in_free[b].record(main)
            main.wait_event(out_free[b])
            d_loc[b].copy_(out["loc"], non_blocking=True)  # plan outputs are reused next step: snapshot them
            d_cls[b].copy_(out["cls"], non_blocking=True)
            out_ready[b].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(out_ready[b])
                h_loc[b].copy_(d_loc[b], non_blocking=True)
                h_cls[b].copy_(d_cls[b], non_blocking=True)
                out_free[b].record(s_out)
        main.wait_stream(s_out)
        main.wait_stream(s_in)

    e2e_steps(max(2, args.warmup))
    torch.cuda.synchronize()
    barrier()
    e0.record()
    e2e_steps(args.steps)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = frames / (e2e_ms * 1e-3)
    h2d = bevs.numel() * 4 + trans.numel() * 8 + nat.numel() * 8
    d2h = (h_loc[0].numel() + h_cls[0].numel()) * 4

    # ---------------- e2e_detections: uint8 BEVs in, kept boxes out (SURVEY 8(f2)+(f3)) ----------------
    # The same forward, fed the dataset's bool occupancy grid (V2XSimDet.py:299, before .astype(np.float32)) and followed
    # by the on-device apply_nms_det, so only the kept boxes cross PCIe -- what test_codet.py consumes per frame.
    from v2x_b200 import postproc
    from v2x_b200.postproc import DetPostprocessor
    del d_loc, d_cls, h_loc, h_cls
    h_u8 = [(bevs > 0).to(torch.uint8).pin_memory() for _ in range(NBUF)]
    d_u8 = [torch.empty_like(h_u8[0], device=dev) for _ in range(NBUF)]
    anchors = torch.from_numpy(_anchor_table()).to(dev)
    posts = [DetPostprocessor(n_maps, 256 * 256 * 6, cap=2048, device=dev) for _ in range(NBUF)]
    h_det = [torch.empty(posts[0].buf.shape, dtype=torch.int32).pin_memory() for _ in range(NBUF)]

    def det_steps(k):
        for i in range(k):
            b = i % NBUF
            with torch.cuda.stream(s_in):
                s_in.wait_event(in_free[b])
                d_u8[b].copy_(h_u8[b], non_blocking=True)
                d_trans[b].copy_(h_trans[b], non_blocking=True)
                d_nat[b].copy_(h_nat[b], non_blocking=True)
                in_ready[b].record(s_in)
            main.wait_event(in_ready[b])
            with torch.no_grad():
                out = model(d_u8[b], d_trans[b], d_nat[b], batch_size=B)
            in_free[b].record(main)
            main.wait_event(out_free[b])
            posts[b].run(out["loc"], out["cls"], anchors)
            out_ready[b].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(out_ready[b])
                posts[b].fetch_async(h_det[b])
                out_free[b].record(s_out)
        main.wait_stream(s_out)
        main.wait_stream(s_in)

    det_steps(max(2, args.warmup))
    torch.cuda.synchronize()
    barrier()
    e0.record()
    det_steps(args.steps)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    det_ms = float(t.item())
    dets = posts[(args.steps - 1) % NBUF].unpack(h_det[(args.steps - 1) % NBUF])
    det_info = {"value": frames / (det_ms * 1e-3), "unit": "frames/s", "ms_per_step": det_ms / args.steps,
                "h2d_bytes_per_step": h_u8[0].numel() + trans.numel() * 8 + nat.numel() * 8,
                "d2h_bytes_per_step": h_det[0].numel() * 4,
                "kept_boxes_per_agent": sum(len(d["selected_idx"]) for d in dets) / max(1, len(dets)),
                "api": "coperception.models.det.V2VNet.forward(uint8 BEV) + v2x_b200.postproc (apply_nms_det on device), "
                       "pinned host in/out, 3-stream pipeline"}

    if rank != 0:
        return
    # ---------------- roofline of the conv kernel family ----------------
    # conv time per step = the conv launches of one step replayed back to back on ONE stream, timed live with CUDA
    # events on that stream (the step itself overlaps the x_4 branch with the warp kernel on a side stream, so
    # "step minus the other kernels" would credit hidden time to the convs).
    peaks = measured_peaks()
    conv_launches = [l for l in plan.launches if getattr(l, "flops", 0.0) > 0]
    other_launches = [l for l in plan.launches if getattr(l, "flops", 0.0) == 0]
    other_ms = time_launch_list(other_launches)
    step_ms = ms_total / args.steps
    conv_ms = time_launch_list(conv_launches)
    alg_flops = GFLOP_PER_FRAME * 1e9 * B
    achieved = alg_flops / (conv_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor",
                "kernel": "v2x::conv_tc_kernel<BN,PLANES,KSTEPS,HALO> (the %d conv launches of a step)" % len(conv_launches),
                "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                "traffic": conv_traffic(B), "traffic_unit": "bytes of DRAM read+write per step over the conv launches (ncu --set full, profiles/conv_traffic.json)",
                "peak_source": peaks["source"],
                "conv_ms_per_step": conv_ms, "other_kernels_ms_per_step": other_ms,
                "conv_share_of_step": conv_ms / step_ms,
                "launches_per_step": {"conv": len(conv_launches), "other": len(other_launches)},
                "algorithmic_gflop_per_frame": GFLOP_PER_FRAME,
                "note": "algorithmic FLOPs (SURVEY 8(d)) / serial conv kernel time; executed FLOPs are ~1% higher "
                        "(13->16 channel pad, block-diagonal head 1x1, tap-pack edge columns); conv + other can exceed "
                        "the step because the x_4 branch overlaps the warp kernel"}

    # ---------------- CPU baseline (oracle port) on this host, bounded sample ----------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:   # rank 0 at N = 1 only
        fps, spf, threads = cpu_reference(3, 1)
        cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "3 timed forwards of 1 scene (5 agents), median %.3f s/frame; torch CPU fp32, best of "
                         "{8,16,32,64,all} threads" % spf}

    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if planes == 1 else "bf16x3", "data": "synthetic",
            "config": {"workload": "V2VNet 5-agent detection fwd, 256x256x13 BEV, gnn_iter=3, layer=3 (BASELINE configs[1])",
                       "scenes_per_gpu_per_step": B, "agents": AGENTS, "precision": args.precision,
                       "l2": "no flush: per-step inputs %.0f MB and activations ~%.1f GB exceed the 126 MB L2"
                             % (B * 17.04, 0.4 * B),
                       "parallelism": ("unit-sharded x%d (40 agent-major units per GPU), one NCCL %s of layer-3 "
                                       "maps per step overlapped with the x_4 branch"
                                       % (world, "all-gather" if args.exchange == "allgather"
                                          else "neighbour exchange (grouped send/recv of the 4 other agents' maps)"))
                       if unit_sharded
                       else "scene-sharded x%d, no data-path collective" % world},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps,
                    "api": "coperception.models.det.V2VNet.forward (pinned host in/out, 3-stream pipeline)"},
            "e2e_detections": det_info,
            "gpu_launches": plan.n_kernels * args.steps,
            "kernels_per_step": plan.n_kernels,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    emit(line)


_REAL_STDOUT = None


def emit(line: dict):
    """The driver reads ONE JSON line from stdout; libraries (NCCL banners, the reference's prints) are kept off it
    by pointing fd 1 at stderr for the life of the process and writing the result to the saved descriptor."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=8, help="scenes (frames) per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", default="unit", choices=["unit", "scene"], help="multi-GPU partition (N > 1)")
    ap.add_argument("--exchange", default="allgather", choices=["neighbours", "allgather"],
                    help="unit-sharded x_3 exchange: one all-gather (default; measured at 2/4/8 GPUs), or NCCL send/recv of "
                         "just the needed neighbour maps (equal at 2/4 GPUs, not yet measured at 8)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.gpus > 1 and args.impl == "ours" and "WORLD_SIZE" not in os.environ:
        # `python bench.py --gpus N` typed directly: relaunch as one rank per GPU (the driver launches torchrun itself)
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        os.dup2(_REAL_STDOUT, 1)
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                                  str(args.gpus), "--master-addr", "127.0.0.1", "--master-port", str(port),
                                  os.path.abspath(__file__)] + sys.argv[1:])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and args.impl == "ours":
        print("bench.py: --gpus %d but WORLD_SIZE=%d; measuring %d rank(s)" % (args.gpus, world, world), file=sys.stderr)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

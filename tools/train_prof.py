"""Where does a training step spend its time?  Runs one V2VNet train step (forward + backward, 4 scenes) under
`ncu --metrics gpu__time_duration.sum` and aggregates the launch list by kernel -> gpurun_out/train_prof.txt.
   python tools/train_prof.py            (on the GPU box)"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "v2x-sim_b200")]


def child(scenes):
    import torch
    from coperception.models.det import V2VNet
    from v2x_b200 import default_det_config
    from v2x_b200 import synthetic as synth
    sd = synth.v2vnet_det_state(0)
    bevs, trans, nat = synth.make_scene(scenes, 5, 0)
    model = V2VNet(default_det_config(), 3, 3, 256, num_agent=5)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    b, t, n = bevs.cuda(), trans.cuda(), nat.cuda()
    for it in range(3):
        if it == 2:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
        for p in model.parameters():
            p.grad = None
        out = model(b, t, n, batch_size=scenes)
        loss = out["cls"].square().mean() + out["loc"].square().mean()
        if it == 2:
            torch.cuda.synchronize()
            print("FWD_DONE", flush=True)
        loss.backward()
        torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


def main():
    scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    if "--child" in sys.argv:
        return child(scenes)
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    log = os.path.join(out_dir, "train_prof_launches.csv")
    subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--profile-from-start", "off",
                    "--csv", "--log-file", log, sys.executable, os.path.abspath(__file__), str(scenes), "--child"],
                   check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for r in csv.reader(open(log)):
        if len(r) > 14 and r[0].isdigit():
            name = r[4].split("(")[0].replace("void ", "").replace("v2x::", "")
            name = name.split("<")[0] if not name.startswith(("conv_tc", "conv_pack3", "conv_wgrad", "channel_reduce", "resample2")) else name
            us = float(r[14]) / 1e3
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += us
            total += us
            n += 1
    lines = ["one V2VNet training step (forward + backward), %d scenes = %d maps: %d kernel launches, %.1f ms of kernel time "
             "(ncu gpu__time_duration, serialised)" % (scenes, scenes * 5, n, total / 1e3), "",
             "%-64s %6s %10s %6s" % ("kernel", "count", "us", "share")]
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-64s %6d %10.0f %5.1f%%" % (name[:64], cnt, us, 100 * us / max(total, 1e-9)))
    open(os.path.join(out_dir, "train_prof.txt"), "w").write("\n".join(lines) + "\n")
    os.remove(log)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()

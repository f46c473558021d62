"""Training-step timing on the sm_100a path: train-mode forward + backward of det V2VNet / FaFNet through
the drop-in modules (torch.autograd boundary), CUDA events, with the tensor-core and the CUDA-core weight-gradient kernels.
   python tools/train_bench.py [scenes]        -> gpurun_out/train_bench.json
The eager tape re-packs every filter each step (weights change under the optimizer) and launches ~600 kernels; nothing is
captured in a CUDA graph yet, so small batches are launch-bound -- the number to read is the large-batch one."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "v2x-sim_b200")]


def run(scenes, steps=5):
    import torch
    from coperception.models.det import FaFNet, V2VNet
    from v2x_b200 import default_det_config
    from v2x_b200 import synthetic as synth
    res = {}
    for name in ("v2vnet", "fafnet"):
        torch.manual_seed(0)
        if name == "v2vnet":
            sd = synth.v2vnet_det_state(0)
            bevs, trans, nat = synth.make_scene(scenes, 5, 0)
            model = V2VNet(default_det_config(), 3, 3, 256, num_agent=5)
            call = lambda: model(bevs_d, trans.cuda(), nat.cuda(), batch_size=scenes)   # noqa: E731
        else:
            sd = synth.fafnet_state(0)
            bevs = synth.make_bevs(scenes * 5, 0)
            model = FaFNet(default_det_config(), kd_flag=0, num_agent=5)
            call = lambda: model(bevs_d, batch_size=scenes)   # noqa: E731
        model.load_state_dict(sd, strict=True)
        model = model.cuda().train()
        bevs_d = bevs.cuda()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        fwd = bwd = 0.0
        for it in range(steps + 2):
            opt.zero_grad(set_to_none=True)
            ev[0].record()
            out = call()
            loss = out["cls"].square().mean() + out["loc"].square().mean()
            ev[1].record()
            loss.backward()
            ev[2].record()
            opt.step()
            torch.cuda.synchronize()
            if it >= 2:
                fwd += ev[0].elapsed_time(ev[1]) / steps
                bwd += ev[1].elapsed_time(ev[2]) / steps
        maps = scenes * 5
        res[name] = {"maps": maps, "forward_ms": fwd, "backward_ms": bwd, "maps_per_s": maps / ((fwd + bwd) * 1e-3),
                     "loss": float(loss.item()), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
        del model, opt, out, loss
        torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "--child":
        print(json.dumps(run(int(sys.argv[1]))))
        sys.exit(0)
    scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    out = {"scenes": scenes}
    for mode in ("tc", "cuda"):
        env = dict(os.environ, V2X_WGRAD=mode)
        txt = subprocess.run([sys.executable, os.path.abspath(__file__), str(scenes), "--child"], env=env, capture_output=True, text=True)
        try:
            out["wgrad_" + mode] = json.loads(txt.stdout.strip().splitlines()[-1])
        except Exception:
            out["wgrad_" + mode] = {"error": txt.stderr[-800:]}
        print(mode, out["wgrad_" + mode], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "train_bench.json"), "w"), indent=1)

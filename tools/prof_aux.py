"""Profiler target for the kernels OUTSIDE the conv family that the bench configs do not reach: the DiscoNet / AgentWise
fuse kernels (pair_score, agent_softmax, warp_weighted), Max fusion (warp_reduce), input densification (voxel_scatter,
pack_input_u8), detection post-processing (nms_collect, nms_map) and one V2VNet training step (BN statistics / apply /
backward, resample2, GRU gates fwd / bwd, warp_mean_bwd, wgrad on CUDA cores and on tcgen05).
   ncu --metrics <few> --profile-from-start off --csv --log-file gpurun_out/aux.csv python tools/prof_aux.py
Everything is built and warmed up first; one pass of each phase sits between cudaProfilerStart / Stop."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "v2x-sim_b200")]
import torch  # noqa: E402

from v2x_b200 import default_det_config, nets  # noqa: E402
from v2x_b200 import synthetic as synth  # noqa: E402


def main():
    dev = torch.device("cuda")
    B, A = 8, 5
    bevs, trans, nat = synth.make_scene(B, A, 0)
    bevs, trans, nat = bevs.to(dev), trans.to(dev), nat.to(dev)
    phases = []
    # --- fusion family forwards (eager launch lists)
    from coperception.models import det as det_models
    torch.manual_seed(0)
    for kind, cls in (("disco", det_models.DiscoNet), ("agent", det_models.AgentWiseWeightedFusion), ("max", det_models.MaxFusion)):
        fm = cls(default_det_config(), layer=3, kd_flag=0, num_agent=A).to(dev).eval()   # default-initialised weights
        phases.append(("fusion_" + kind, lambda fm=fm: fm(bevs, trans, nat, batch_size=B)))
    # --- uint8 input + on-device post-processing (what e2e_detections runs)
    from v2x_b200.postproc import DetPostprocessor
    sd = synth.v2vnet_det_state(0)
    v2v = nets.V2VNetDetPlan(sd, B, A, gnn_iter=3, planes="mixed", device=dev, input_mode="u8")
    u8 = (bevs > 0).to(torch.uint8)
    out = v2v.forward(u8, trans, nat)
    sd2 = synth.plant_detections(sd, out["cls"].float().cpu()[::B], per_agent=150)
    v2v = nets.V2VNetDetPlan(sd2, B, A, gnn_iter=3, planes="mixed", device=dev, input_mode="u8")
    import bench
    anchors = torch.from_numpy(bench._anchor_table()).to(dev)
    post = DetPostprocessor(B * A, 256 * 256 * 6, cap=2048, device=dev)

    def det_phase():
        o = v2v.forward(u8, trans, nat)
        post.run(o["loc"], o["cls"], anchors)
    phases.append(("u8_forward_plus_nms", det_phase))
    # --- sparse voxel rows -> dense BEV on device
    rows = torch.nonzero(bevs[:, 0] > 0)                      # (map, h, w, z) of the rotated grid
    vox = torch.stack([rows[:, 0], 255 - rows[:, 2], rows[:, 1], rows[:, 3]], 1).to(torch.int32).contiguous()
    from coperception.models.det import V2VNet
    model = V2VNet(default_det_config(), 3, 3, 256, num_agent=A)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    phases.append(("voxel_forward", lambda: model.forward_voxels(vox, trans, nat, batch_size=B)))
    # --- one training step (4 scenes): forward + backward through torch.autograd
    tb, tt, tn = synth.make_scene(4, A, 1)
    tmodel = V2VNet(default_det_config(), 3, 3, 256, num_agent=A)
    tmodel.load_state_dict(sd, strict=True)
    tmodel = tmodel.to(dev).train()
    tb, tt, tn = tb.to(dev), tt.to(dev), tn.to(dev)

    def train_phase():
        for p in tmodel.parameters():
            p.grad = None
        o = tmodel(tb, tt, tn, batch_size=4)
        (o["cls"].square().mean() + o["loc"].square().mean()).backward()
    phases.append(("train_step_v2vnet", train_phase))

    with torch.no_grad():
        for name, fn in phases[:-1]:
            fn()
    phases[-1][1]()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for name, fn in phases:
        if name.startswith("train"):
            fn()
        else:
            with torch.no_grad():
                fn()
        torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("profiled phases:", [n for n, _ in phases])


if __name__ == "__main__":
    main()

"""GPU: the drop-in nn.Module surface (coperception.models.det.V2VNet / FaFNet) end to end:
strict state_dict load, forward signature, output contract, parity with the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_v2vnet_module_matches_oracle():
    from coperception.models.det import V2VNet
    from oracle import restate, synth
    from v2x_b200 import default_det_config
    sd = synth.v2vnet_det_state(2)
    model = V2VNet(default_det_config(), 3, 3, 256, num_agent=5)
    model.load_state_dict(sd, strict=True)
    model.precision = "bf16x3"
    model = torch.nn.DataParallel(model.cuda().eval(), device_ids=[0])  # as train/test_codet.py wrap it
    bevs, trans, nat = synth.make_scene(1, 5, seed=2, present=[4])
    with torch.no_grad():
        out = model(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)
        ref = restate.v2vnet_det_forward(bevs, trans, nat, sd, batch_size=1)
    assert out["loc"].shape == (5, 256, 256, 6, 1, 6) and out["cls"].shape == (5, 393216, 2)
    for k in ("loc", "cls"):
        err = ((out[k].cpu() - ref[k]).abs().max() / ref[k].abs().max()).item()
        print("dropin v2vnet", k, err)
        assert err < 1e-3
    # a second call reuses the captured CUDA graph and must give the same answer
    with torch.no_grad():
        out2 = model(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)
    assert torch.equal(out["cls"], out2["cls"])


def test_fafnet_module_kd_outputs():
    from coperception.models.det import FaFNet
    from oracle import restate, synth
    from v2x_b200 import default_det_config
    sd = synth.fafnet_state(3)
    model = FaFNet(default_det_config(), kd_flag=1, num_agent=5)
    model.load_state_dict(sd, strict=True)
    model.precision = "bf16x3"
    model = model.cuda().eval()
    bevs = synth.make_bevs(1, 3)
    with torch.no_grad():
        out = model(bevs.cuda(), torch.zeros(1), torch.zeros(1), batch_size=1)  # fusion-style call (Q12)
        ref = restate.fafnet_forward(bevs, sd, stages=True)
    result, x8, x7, x6, x5, x3 = out
    for got, want in ((result["loc"], ref["loc"]), (x8, ref["dec"][0]), (x7, ref["dec"][1]), (x6, ref["dec"][2]),
                      (x5, ref["dec"][3]), (x3, ref["enc"][3])):
        assert got.shape == want.shape
        assert ((got.cpu() - want).abs().max() / want.abs().max()).item() < 1e-3


def test_train_mode_is_refused_where_not_built():
    """Every det model class and every seg class but DiscoNet trains on the sm_100a path (tests/test_gpu_train.py); what is
    not built refuses loudly instead of silently doing something else: training with fewer than 32 compressed channels, seg DiscoNet, the
    seg models' kd_flag outputs, and a When2com in .train() asked for the gated inference pass (training=False)."""
    from coperception.models.det import MeanFusion, When2com
    from coperception.models.seg import DiscoNet as SegDiscoNet
    from v2x_b200 import default_det_config
    args = (torch.zeros((5, 1, 256, 256, 13), device="cuda"), torch.zeros((1, 5, 5, 4, 4), device="cuda"),
            torch.full((1, 5), 5, device="cuda"))
    with pytest.raises(NotImplementedError):
        # (compress_level 1..3 trains since the compresser pair went on the tape; 4 leaves 16 channels, below the 32 it takes)
        MeanFusion(default_det_config(), layer=3, kd_flag=0, num_agent=5, compress_level=4).cuda().train()(*args, batch_size=1)
    with pytest.raises(NotImplementedError):
        When2com(default_det_config(), layer=3, warp_flag=1, num_agent=5).cuda().train()(*args, training=False, batch_size=1)
    for kd in (True, False):     # seg DiscoNet: the train step exists (Tape.disco_fuse) but measured a cosine of 0.9985 on the
        with pytest.raises(NotImplementedError):   # first conv's gradient, below the 0.999 bar, so it stays switched off
            SegDiscoNet(13, 8, 5, kd_flag=kd).cuda().train()(torch.zeros((5, 13, 256, 256), device="cuda"), args[1], args[2])


def _v2v_model(seed=2):
    from coperception.models.det import V2VNet
    from oracle import synth
    from v2x_b200 import default_det_config
    sd = synth.v2vnet_det_state(seed)
    model = V2VNet(default_det_config(), 3, 3, 256, num_agent=5)
    model.load_state_dict(sd, strict=True)
    return model.cuda().eval(), sd


def test_default_precision_is_the_parity_mode():
    """A user who drops the module in gets the mode the 1e-3 parity tests assert (tests/test_gpu_nets.py)."""
    model, _ = _v2v_model()
    assert model.precision == "mixed"


def test_consecutive_forwards_return_independent_tensors():
    """The reference returns fresh tensors; the plan's static output buffers must not leak to the caller (a loop that
    collects outputs would otherwise end up with every entry equal to the last step).  ``alias_outputs`` opts out."""
    from oracle import synth
    model, _ = _v2v_model()
    b1, t1, n1 = synth.make_scene(1, 5, seed=2)
    b2, t2, n2 = synth.make_scene(1, 5, seed=9)
    with torch.no_grad():
        o1 = model(b1.cuda(), t1.cuda(), n1.cuda(), batch_size=1)
        keep = o1["cls"].clone()
        o2 = model(b2.cuda(), t2.cuda(), n2.cuda(), batch_size=1)
    torch.cuda.synchronize()
    assert o1["cls"].data_ptr() != o2["cls"].data_ptr()
    assert torch.equal(o1["cls"], keep) and not torch.equal(o1["cls"], o2["cls"])
    model.alias_outputs = True
    with torch.no_grad():
        a1 = model(b1.cuda(), t1.cuda(), n1.cuda(), batch_size=1)
        a2 = model(b2.cuda(), t2.cuda(), n2.cuda(), batch_size=1)
    assert a1["cls"].data_ptr() == a2["cls"].data_ptr()


def test_inplace_parameter_edit_rebuilds_the_packed_operands():
    """Packed weights are a cache keyed by a fingerprint of the parameters' (pointer, version): p.data.copy_ / EMA /
    torch.nn.init after the first forward must change the next forward."""
    from oracle import restate, synth
    model, sd = _v2v_model()
    bevs, trans, nat = synth.make_scene(1, 5, seed=2)
    with torch.no_grad():
        before = model(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)["cls"].clone()
        model.classification.conv2.bias.add_(0.5)          # tracked in-place edit: version counter bumps
        after = model(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)["cls"].clone()
        # an edit through .data is invisible to the version counter: the opt-in value checksum catches it
        model.verify_weights = True
        model(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)
        model.classification.conv2.bias.data.add_(0.25)
        after_data = model(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)["cls"]
        assert (after_data - after).abs().max().item() > 0.2
        sd2 = {k: v.clone() for k, v in sd.items()}
        sd2["classification.conv2.bias"] = sd["classification.conv2.bias"] + 0.5
        ref = restate.v2vnet_det_forward(bevs, trans, nat, sd2, batch_size=1)
    assert (after - before).abs().max().item() > 0.4
    assert ((after.cpu() - ref["cls"]).abs().max() / ref["cls"].abs().max()).item() < 1e-3


def test_fafnet_kd_returns_the_decompressed_x3():
    """kd_flag == 1 with compress_level > 0: x_3 is encoded_layers[3], i.e. AFTER com_compresser / com_decompresser
    (Backbone.py:138-141) -- ADVICE r1."""
    from coperception.models.det import FaFNet
    from oracle import restate, synth
    from v2x_b200 import default_det_config
    sd = synth.fafnet_state(5, compress_level=2)
    model = FaFNet(default_det_config(), kd_flag=1, num_agent=5, compress_level=2)
    model.load_state_dict(sd, strict=True)
    model.precision = "fp16x3"
    model = model.cuda().eval()
    bevs = synth.make_bevs(1, 5)
    with torch.no_grad():
        out = model(bevs.cuda(), batch_size=1)
        ref = restate.fafnet_forward(bevs, sd, compress_level=2, stages=True)
    x3 = out[5]
    assert ((x3.cpu() - ref["enc"][3]).abs().max() / ref["enc"][3].abs().max()).item() < 1e-3


def test_forward_under_enable_grad_warns_once():
    import warnings
    from oracle import synth
    model, _ = _v2v_model()
    bevs, trans, nat = synth.make_scene(1, 5, seed=2)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        model(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)
        model(bevs.cuda(), trans.cuda(), nat.cuda(), batch_size=1)
    assert sum("no autograd graph" in str(x.message) for x in w) == 1


@pytest.mark.gpu
def test_peer_region_single_rank_protocol():
    """The device-side exchange (csrc/peer_kernels.cu, sharding.PeerRegion) with world = 1: the step counter advances, the
    push lands every plane of the source in the payload slot it names, flags carry the step number, no wait times out.
    (Cross-process mapping of the regions is what tests/multigpu_check.py with V2X_EXCHANGE=push checks on 2+ GPUs.)"""
    import torch
    from v2x_b200 import ops, sharding
    ops.require_gpu()
    dev = torch.device("cuda")
    planes, n, elems = 2, 3, 32 * 32 * 64
    region = sharding.PeerRegion(planes * 5 * elems * 2, 0, 1, device=dev)
    payload = region.payload((planes, 5, elems), torch.float16)
    assert payload.data_ptr() == region.payload_ptr(0) and float(payload.abs().max()) == 0.0
    for step in range(1, 4):
        src = torch.randn((planes, n, elems), device=dev).half()
        region.begin()
        region.push(src, 2 * elems, 5 * elems, [planes])     # units 2..4 of the payload
        region.wait()
        region.done()
        torch.cuda.synchronize()
        region.check()
        assert int(region.step.item()) == step
        assert torch.equal(payload[:, 2:5], src) and float(payload[:, :2].abs().max()) == 0.0
        flags = region._bytes[:sharding.PEER_FLAG_BYTES].view(torch.int64)
        assert int(flags[0]) == step and int(flags[8]) == step      # ready[0], consumed[0]
    region.close()


@pytest.mark.gpu
def test_sharded_plan_push_exchange_world1_matches_unsharded_plan():
    """V2VNetDetShardedPlan with the device-side push exchange (one CUDA graph: x_4 branch forked beside begin / push /
    wait / fuse) on a single rank reproduces the unsharded plan bit for bit, eagerly and as a replayed graph."""
    import torch
    from v2x_b200 import nets, synthetic
    sd = synthetic.v2vnet_det_state(3)
    bevs, trans, nat = synthetic.make_scene(2, 5, seed=3, present=[5, 3])
    args = (bevs.cuda(), trans.cuda(), nat.cuda())
    full = nets.V2VNetDetPlan(sd, 2, 5, planes="mixed")
    want = {k: v.clone() for k, v in full.forward(*args).items()}
    plan = nets.V2VNetDetShardedPlan(sd, 2, 5, 0, 1, planes="mixed", exchange="push")
    got = {k: v.clone() for k, v in plan.forward(*args).items()}
    plan.capture()
    for _ in range(3):
        rep = plan.forward(*args)
    torch.cuda.synchronize()
    plan.peer.check()
    for k in want:
        assert torch.equal(got[k], want[k]), k
        assert torch.equal(rep[k], want[k]), k
    assert int(plan.peer.step.item()) == 1 + 2 + 3      # eager forward, two warm-up steps of capture(), three replays
    plan.close()

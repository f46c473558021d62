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


def test_train_mode_is_refused():
    from coperception.models.det import V2VNet
    from v2x_b200 import default_det_config
    model = V2VNet(default_det_config(), 3, 3, 256).cuda().train()
    with pytest.raises(NotImplementedError):
        model(torch.zeros((5, 1, 256, 256, 13), device="cuda"), torch.zeros((1, 5, 5, 4, 4), device="cuda"),
              torch.full((1, 5), 5, device="cuda"), batch_size=1)

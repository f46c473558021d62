"""GPU: constructor options of the drop-in modules beside the reference scripts' defaults.

This file sorts LAST on purpose.  It was written after the round's GPU budget was spent: everything here has been checked
as far as a CPU can check it -- the oracle is pinned to live-reference fixtures for every case
(tests/test_oracle_golden.py), the host logic is covered in tests/test_host_logic.py, the library builds -- but its first
run on a B200 is the driver's round-end run.  Each case only recombines kernels the files before it already exercise
(the same launches with a different operand, layer or channel count); the parity bounds are the ones of those files.
"""
import pytest

pytestmark = pytest.mark.gpu

MODES = ["mixed", "fp16x3", "bf16"]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("tag", ["when2com_det_noquery_activated_seed41",
                                 "when2com_det_layer2_sparse_activated_seed42_present4",
                                 "when2com_det_layer2_noquery_nowarp_softmax_B2_seed43"])
def test_when2com_det_options(tag, mode, golden_dir, parity_log):
    """has_query=False (ones as every agent's query, When2com.py:241-245), sparse=True (a no-op in the reference,
    :374-412) and layer=2 (:167-190), against the oracle and the live-reference fixtures."""
    from test_gpu_nets import check_when2com_det
    check_when2com_det(tag, mode, golden_dir, parity_log)


@pytest.mark.parametrize("mode", MODES)
def test_seg_when2com_options(mode, golden_dir, parity_log):
    """seg When2Com_UNet(has_query=False, sparse=True) (When2Com_UNet.py:219-225, 408-446)."""
    from test_gpu_seg import check_seg_model
    check_seg_model("seg_when2com_noquery_sparse_activated_seed44", "when2com", mode, golden_dir, parity_log)


def test_when2com_refuses_what_the_reference_cannot_run():
    """MO_flag=False and layer=2 + argmax_test raise in the reference itself (profiles/r02_reference_option_probe.txt)."""
    import torch
    from coperception.models.det import When2com
    from oracle import synth
    from v2x_b200 import default_det_config
    bevs, trans, nat = synth.make_scene(1, 5, 0)
    m = When2com(default_det_config(), layer=2, num_agent=5).cuda().eval()
    with torch.no_grad():
        with pytest.raises(NotImplementedError):
            m(bevs.cuda(), trans.cuda(), nat.cuda(), training=False, MO_flag=False, batch_size=1)
        with pytest.raises(NotImplementedError):
            m(bevs.cuda(), trans.cuda(), nat.cuda(), training=False, inference="argmax_test", batch_size=1)

"""coperception.models.det.When2com on the sm_100a path (reference: CP/models/det/When2com.py:15-332)."""
import torch
import torch.nn as nn

from ._base import B200DetModel
from ._schema import BackboneParams


class _CBR(nn.Module):
    """Parameter container of Conv2DBatchNormRelu (Backbone.py:309-342): ``cbr_unit.0`` conv, ``cbr_unit.1`` BN."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.cbr_unit = nn.Sequential(nn.Conv2d(cin, cout, 3, stride, 1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class PolicyNet4(nn.Module):
    """Parameters of PolicyNet4 (When2com.py:335-359): a full Backbone + conv1..conv5."""

    def __init__(self, in_channels=13):
        super().__init__()
        self.lidar_encoder = BackboneParams(in_channels)
        self.conv1 = _CBR(512, 512, 1)
        self.conv2 = _CBR(512, 256, 1)
        self.conv3 = _CBR(256, 256, 2)
        self.conv4 = _CBR(256, 256, 1)
        self.conv5 = _CBR(256, 256, 2)


class KmGenerator(nn.Module):
    """Parameters of KmGenerator (When2com.py:415-430)."""

    def __init__(self, out_size=128, input_feat_sz=32.0):
        super().__init__()
        feat_map_sz = input_feat_sz // 4
        self.n_feat = int(256 * feat_map_sz * feat_map_sz)
        self.fc = nn.Sequential(nn.Linear(self.n_feat, 256), nn.ReLU(inplace=True), nn.Linear(256, 128),
                                nn.ReLU(inplace=True), nn.Linear(128, out_size))


class MIMOGeneralDotProductAttention(nn.Module):
    def __init__(self, query_size, key_size, warp_flag):
        super().__init__()
        self.linear = nn.Linear(query_size, key_size)
        self.warp_flag = warp_flag


class When2com(B200DetModel):
    """When2com / who2com (https://github.com/GT-RIPL/MultiAgentPerception): key/query handshake attention over the
    agents' layer-3 maps.  Constructor / forward signatures, parameter names and the ``*_paras`` lists match the
    reference (When2com.py:22-92, :150-161)."""

    def __init__(self, config, n_classes=21, in_channels=13, feat_channel=512, feat_squeezer=-1, attention="additive",
                 has_query=True, sparse=False, layer=3, warp_flag=1, image_size=512, shared_img_encoder="unified",
                 key_size=1024, query_size=32, num_agent=5, compress_level=0, only_v2i=False):
        super().__init__(config, layer, in_channels, num_agent=num_agent, only_v2i=only_v2i)
        if layer not in (2, 3, 4):
            raise NotImplementedError("When2com communicates at layer 2, 3 or 4 (When2com.py:167-190)")
        self.compress_level = compress_level
        self.u_encoder = BackboneParams(in_channels, compress_level)
        self.decoder = BackboneParams(in_channels)
        # ``sparse`` is accepted and, exactly as in the reference, changes nothing: it is handed to
        # MIMOGeneralDotProductAttention.forward, which never reads it (When2com.py:255-257, 374-412 -- always nn.Softmax;
        # the live reference gives bit-identical outputs for sparse=True / False, profiles/r02_reference_option_probe.txt)
        self.sparse, self.key_size, self.query_size = sparse, key_size, query_size
        self.shared_img_encoder, self.has_query, self.warp_flag = shared_img_encoder, has_query, warp_flag
        self.key_net = KmGenerator(out_size=key_size, input_feat_sz=image_size / 32)
        self.attention_net = MIMOGeneralDotProductAttention(query_size, key_size, warp_flag)
        self.query_key_net = PolicyNet4(in_channels=in_channels)
        if has_query:      # without it every agent's query is a vector of ones (When2com.py:66-69, 241-245)
            self.query_net = KmGenerator(out_size=query_size, input_feat_sz=image_size / 32)
        # parameter groups the reference exposes (When2com.py:72-88)
        self.attention_paras = list(self.attention_net.parameters())
        self.img_net_paras = list(self.u_encoder.parameters()) + list(self.decoder.parameters())
        self.policy_net_paras = (list(self.query_key_net.parameters()) + list(self.key_net.parameters())
                                 + self.attention_paras)
        if has_query:
            self.policy_net_paras = self.policy_net_paras + list(self.query_net.parameters())
        self.all_paras = self.img_net_paras + self.policy_net_paras

    def forward(self, bevs, trans_matrices, num_agent_tensor, maps=None, vis=None, training=True, MO_flag=True,
                inference="activated", batch_size=1):
        from v2x_b200 import nets
        if not MO_flag:
            # the reference itself cannot run this: with a single query prob_action is [B, A, 1] and the reshape of the
            # A x A ``small_bis`` at When2com.py:273-274 raises (profiles/r02_reference_option_probe.txt)
            raise NotImplementedError("MO_flag=False raises in the reference too (When2com.py:274 reshapes an A x A "
                                      "identity to [1, A, 1]); it is not built on the sm_100a path")
        if self.training:       # what the training tape does not take is refused before the device is touched
            if not training:
                raise NotImplementedError("model.train() with training=False (the gated second pass) is not built")
            if self.compress_level > 3:
                raise NotImplementedError("training with compress_level > 3 (fewer than 32 compressed channels) is not "
                                          "built on the sm_100a path")
            if not self.has_query or self.layer != 3:
                raise NotImplementedError("training on the sm_100a path: has_query=True, layer 3 (the reference "
                                          "scripts' defaults)")
        dev = bevs.device
        if dev.type != "cuda":
            raise RuntimeError("v2x_b200 When2com needs CUDA tensors (no CPU fallback); got %s" % dev)
        if self.training:
            # model.train(): the train-mode forward with a backward pass behind torch.autograd, as FaFModule.step drives
            # it (CoDetModule.py:232-247 calls the model with its default training=True)
            from v2x_b200.train import When2comTrainStep
            loc, cls = When2comTrainStep.apply(self, bevs, trans_matrices, num_agent_tensor, int(batch_size),
                                               *self.parameters())
            return {"loc": loc, "cls": cls}
        self._check_eval()
        if inference not in ("softmax", "activated", "argmax_test"):
            raise ValueError("Incorrect inference mode")
        if self.layer != 3 and inference == "argmax_test" and not training:
            raise NotImplementedError("argmax_test only exists for layer 3 in the reference (When2com.py:289-291 hands "
                                      "the fused map to the decoder's layer-3 slot whatever self.layer is)")
        key = ("w2c", int(batch_size), dev.index, self.precision, bool(training), inference)
        plan = self._get_plan(key, lambda: nets.When2comDetPlan(
            self._state(), int(batch_size), self.agent_num, planes=self._planes(), device=dev, warp_flag=self.warp_flag,
            inference=inference, training_pass_only=bool(training), only_v2i=self.only_v2i, has_query=self.has_query,
            layer=self.layer))
        return plan.forward(bevs.to(torch.float32), trans_matrices.to(torch.float64), num_agent_tensor.to(torch.int64))

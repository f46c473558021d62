"""coperception.models.seg.When2Com_UNet on the sm_100a path (reference: CP/models/seg/When2Com_UNet.py:10-307)."""
import torch
import torch.nn as nn

from ..det.When2com import KmGenerator, MIMOGeneralDotProductAttention, _CBR
from .SegModelBase import DoubleConv, Down, SegModelBase


class PolicyNet4(nn.Module):
    """Parameters of the seg PolicyNet4 (When2Com_UNet.py:310-339): own inc / down1..3 + conv1..conv5."""

    def __init__(self, in_channels=13):
        super().__init__()
        self.inc = DoubleConv(in_channels, 64)
        self.down1, self.down2, self.down3 = Down(64, 128), Down(128, 256), Down(256, 512)
        self.conv1, self.conv2, self.conv3 = _CBR(512, 512, 1), _CBR(512, 256, 1), _CBR(256, 256, 2)
        self.conv4, self.conv5 = _CBR(256, 256, 1), _CBR(256, 256, 2)


class When2Com_UNet(SegModelBase):
    def __init__(self, config, n_classes=21, in_channels=13, has_query=True, sparse=False, layer=3, warp_flag=1,
                 image_size=512, shared_img_encoder="unified", key_size=1024, query_size=32, num_agent=5,
                 compress_level=0, only_v2i=False):
        super().__init__(in_channels, n_classes, num_agent=num_agent, compress_level=compress_level, only_v2i=only_v2i)
        # ``sparse`` is accepted and, as in the reference, changes nothing: the attention module never reads it
        # (When2Com_UNet.py:235-237, 408-446 -- always nn.Softmax)
        self.sparse, self.key_size, self.query_size = sparse, key_size, query_size
        self.shared_img_encoder, self.has_query, self.warp_flag, self.layer = shared_img_encoder, has_query, warp_flag, layer
        self.key_net = KmGenerator(out_size=key_size, input_feat_sz=image_size / 32)
        self.attention_net = MIMOGeneralDotProductAttention(query_size, key_size, warp_flag)
        self.query_key_net = PolicyNet4(in_channels=in_channels)
        if has_query:      # without it every agent's query is a vector of ones (When2Com_UNet.py:53-56, 219-225)
            self.query_net = KmGenerator(out_size=query_size, input_feat_sz=image_size / 32)
        self.attention_paras = list(self.attention_net.parameters())
        self.policy_net_paras = (list(self.query_key_net.parameters()) + list(self.key_net.parameters())
                                 + self.attention_paras)
        if has_query:
            self.policy_net_paras = self.policy_net_paras + list(self.query_net.parameters())

    def forward(self, bevs, trans_matrices, num_agent_tensor, maps=None, vis=None, training=True, MO_flag=True,
                inference="activated", batch_size=1):
        from v2x_b200 import nets_seg
        if not MO_flag:
            # the reference cannot run it either: the A x A ``small_bis`` is reshaped to [1, A, 1] (When2Com_UNet.py:243-244)
            raise NotImplementedError("MO_flag=False raises in the reference too (When2Com_UNet.py:244); it is not built "
                                      "on the sm_100a path")
        if self.training:
            # model.train(): train-mode forward with a backward pass behind torch.autograd, as SegModule.step drives it
            # (SegModule.py:66-89 calls the model with training=True)
            if not training:
                raise NotImplementedError("model.train() with training=False (the gated pass) is not built")
            if not self.has_query:
                raise NotImplementedError("training on the sm_100a path: has_query=True (the reference scripts' default)")
            batch = int(bevs.shape[0]) // self.num_agent
            return self._train_forward(bevs, (trans_matrices, num_agent_tensor, batch, self.num_agent, bool(self.only_v2i),
                                              "when2com", int(self.warp_flag)))
        self._check(bevs)
        if inference not in ("softmax", "activated", "argmax_test"):
            raise ValueError("Incorrect inference mode")
        batch = int(bevs.shape[0]) // self.num_agent   # the reference recomputes it from the input (When2Com_UNet.py:166)
        key = ("w2c", batch, bevs.device.index, self.precision, bool(training), inference)
        plan = self._get_plan(key, lambda: nets_seg.SegWhen2comPlan(
            self._state(), batch, self.num_agent, planes=self._planes(), device=bevs.device, warp_flag=self.warp_flag,
            inference=inference, training=bool(training), only_v2i=self.only_v2i, has_query=self.has_query))
        return plan.forward(bevs.to(torch.float32).contiguous(), trans_matrices.to(torch.float64),
                            num_agent_tensor.to(torch.int64))

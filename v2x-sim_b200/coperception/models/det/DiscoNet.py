"""coperception.models.det.DiscoNet on the sm_100a path (reference: CP/models/det/DiscoNet.py:7-155)."""
from ._fusion import FusionBase, PairWeightNet


class DiscoNet(FusionBase):
    """Pixel-wise learned weights: softmax over the list members at every pixel, weighted sum (DiscoNet.py:80-107)."""
    KIND = "disco"

    def __init__(self, config, layer=3, in_channels=13, kd_flag=True, num_agent=5, compress_level=0, only_v2i=False):
        super().__init__(config, layer, in_channels, kd_flag, num_agent, compress_level, only_v2i)
        if layer == 3:        # DiscoNet.py:23-26: the weight net only exists for layers 3 and 2
            self.pixel_weighted_fusion = PairWeightNet(256)
        elif layer == 2:
            self.pixel_weighted_fusion = PairWeightNet(128)

    def forward(self, bevs, trans_matrices, num_agent_tensor, batch_size=1):
        """kd_flag == 1: (result, x_8, x_7, x_6, x_5, fused); else (result, save_agent_weight_list) where the list holds,
        per present (scene, agent), the per-pixel weight maps [32,32] of its list members in the H-flipped domain the
        reference computes them in (DiscoNet.py:55,97-113,125-129)."""
        import torch
        plan, result = self._run(bevs, trans_matrices, num_agent_tensor, batch_size)
        if plan is None:        # model.train(): the train step (v2x_b200/train.py); the visualisation list is not rebuilt
            return (result, *self._train_kd) if self.kd_flag == 1 else (result, [])
        if self.kd_flag == 1:
            return (result, *plan.kd_layers())
        hw = int(round(plan.fuse.scores.shape[-1] ** 0.5))
        scores = plan.fuse.scores.view(batch_size, self.agent_num, self.agent_num, hw, hw)
        na = num_agent_tensor[:, 0].tolist()
        weights = []
        for b in range(batch_size):
            n = int(na[b])
            for i in range(n):
                ks = [i] + [k for k in range(n) if k != i and not (self.only_v2i and i != 0 and k != 0)]
                w = torch.softmax(scores[b, i, ks], dim=0)
                weights.append([torch.flip(w[j], (0,)) for j in range(len(ks))])
        return result, weights

"""Precision modes of the sm_100a path: storage format of the act tensors + tensor-core passes per conv launch.

  "bf16"   : one bf16 plane everywhere, 1 MMA per k-step -- the throughput mode BASELINE.json's config names
             (~2e-2 of the fp32 reference after ~25 layers; NOT inside the north-star 1e-3 tolerance).
  "fp16x3" : fp16 hi/lo planes (x = hi + lo, ~22 mantissa bits) for activations AND weights, 3 MMAs per k-step
             (hi*hi + hi*w_lo + a_lo*w_hi, fp32 accumulate): ~1e-5 of the fp32 reference.
  "mixed"  : (default) fp16 hi/lo activations everywhere; per launch the FEWEST tensor-core passes that keep the
             whole forward inside 1e-3 of the fp32 reference WITH MARGIN.  Measured layer by layer on B200
             (tests/precision_sweep.py -> profiles/r02_precision_sweep_*.txt; CPU emulation oracle/precision_study.py):
             running ONE conv with single-rounded (11-bit) weights costs 3-7e-4 of max-norm error at the V2VNet
             outputs and 4-10e-4 at FaFNet's, so at most one or two layers can afford it; the ConvGRU is the exception
             -- its gates squash the error to 2e-4 at the outputs even with BOTH operands single-rounded (its own
             output h3 is then 8e-4 off).  Policy: the GRU launches (32% of the step's FLOPs) run 1 pass (hi*hi),
             conv5_1 -- the heaviest tensor-bound decoder layer -- runs 2 passes (a_hi*w + a_lo*w, weights one fp16
             plane), everything else runs 3.  Measured: V2VNet 5.0e-4, FaFNet 5.3e-4, when2com < 9e-4 (gates exact).
             Adding conv8_1 / conv7_1 bought 2% / 2.5% more speed but took FaFNet to 7.7e-4 / 1.14e-3.

"bf16x3" is accepted as an alias of "fp16x3" (round 1's split mode used bf16 planes; fp16 planes cost the same and are
64x more accurate).  A plan's ``planes`` argument may be 1 (= "bf16"), 2 (= "fp16x3"), a mode name or a Precision.
"""
from __future__ import annotations

import os
from typing import Dict, Union

# layer name (as in the reference state_dict, without prefix; "gru" = the ConvGRU W_ih launches) -> MMA passes
MIXED_POLICY: Dict[str, int] = {"gru": 1, "conv5_1": 2}


def _env_policy():
    """V2X_MIXED_POLICY="gru=1,conv5_1=2" replaces the table (precision experiments, tests/precision_sweep.py)."""
    txt = os.environ.get("V2X_MIXED_POLICY")
    if txt is None:
        return MIXED_POLICY
    return {k.strip(): int(v) for k, v in (kv.split("=") for kv in txt.split(",") if kv.strip())}


class Precision:
    def __init__(self, name: str):
        name = {"bf16x3": "fp16x3"}.get(name, name)
        if name not in ("bf16", "fp16x3", "mixed"):
            raise ValueError("precision must be one of 'bf16', 'fp16x3', 'mixed' (got %r)" % (name,))
        self.name = name
        self.planes = 1 if name == "bf16" else 2

    def mmas(self, layer: str) -> int:
        if self.planes == 1:
            return 1
        if self.name == "mixed":
            return _env_policy().get(layer, 3)
        return 3

    def __repr__(self):
        return "Precision(%s)" % self.name

    # plans and weights are keyed / compared by mode
    def __eq__(self, other):
        return isinstance(other, Precision) and other.name == self.name

    def __hash__(self):
        return hash(self.name)


DEFAULT = os.environ.get("V2X_PRECISION", "mixed")


def resolve(p: Union[int, str, "Precision", None]) -> Precision:
    if isinstance(p, Precision):
        return p
    if p is None:
        return Precision(DEFAULT)
    if isinstance(p, int):
        return Precision({1: "bf16", 2: "fp16x3"}[p])
    return Precision(p)

"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the collaborative-perception hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the CPU baseline.  The product
path (``v2x-sim_b200/``) never imports this package and fails loudly when its CUDA
library is missing.

Contents
  * ``restate.py``   fp32 torch-CPU restatement of the reference modules on the path
                     (each function cites the reference file:line it follows).
  * ``synth.py``     seeded synthetic inputs / weights.
  * ``postproc.py``  numpy restatement of score -> decode -> NMS -> AP (mAP parity).
  * ``ref_loader.py``imports the LIVE reference modules from /root/reference (only
                     available in the build container; used by ``gen_golden.py``).
  * ``gen_golden.py``writes ``tests/golden/*.npz`` from the live reference.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the restatement is pinned against outputs of the live
reference modules themselves, committed as fixtures under ``tests/golden/`` together
with the generating script (``oracle/gen_golden.py``).
"""

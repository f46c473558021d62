"""TEST INFRASTRUCTURE ONLY -- stage the LIVE reference for the GPU box.

The reference is Python; `/root/reference` exists only in the build container.  This recipe copies the few source
directories the hot path needs (models/, configs/, utils/convolutional_rnn/ -- ~300 KB, no data, no binaries) from
`/root/reference/coperception/coperception` into the git-ignored `oracle/_ref/coperception/`, which travels to the GPU
box with the gpurun snapshot (it is NOT in .gpurunignore) exactly like the built libv2x_b200.so.  Nothing of it enters
the git history.  `oracle/ref_loader.py` imports the reference from there when `/root/reference` is absent, so
`bench.py --impl reference` and the `cpu_baseline` leg time the UNMODIFIED reference modules (kind = "reference").

  python -m oracle.make_ref          # run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import os
import shutil
import sys

SRC = "/root/reference/coperception/coperception"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "coperception")
PARTS = ("models", "configs", os.path.join("utils", "convolutional_rnn"))


def main(quiet=False):
    if not os.path.isdir(os.path.join(SRC, "models", "det")):
        raise RuntimeError("reference tree not found at %s (the GPU box uses the staged copy)" % SRC)
    for part in PARTS:
        dst = os.path.join(DST, part)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, part), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.so"))
    with open(os.path.join(os.path.dirname(DST), "README"), "w") as f:
        f.write("Staged copy of the reference's model sources (oracle/make_ref.py); git-ignored, never committed.\n")
    if not quiet:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print("staged %d reference files under %s" % (n, DST))


if __name__ == "__main__":
    main()
    sys.exit(0)

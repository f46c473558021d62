"""Drop-in ``coperception`` package: the detection models of the hot path on the sm_100a kernels.

Same import paths, class names, constructor / forward signatures and state_dict keys as the
reference (CP/models/det/__init__.py:1-10), so tools/det/{train,test}_codet.py construct and
call these classes unchanged.  Sub-packages this repo does not replace (configs, datasets, utils:
the callers and CPU post-processing, SURVEY.md section 8 "out of scope") resolve to an installed
reference ``coperception`` when one is importable: its directories are appended to ``__path__``.
"""
import importlib.machinery as _m
import os as _os
import sys as _sys

__path__ = [_os.path.dirname(_os.path.abspath(__file__))]


def _extend_with_reference(pkg_path, sub=()):
    """Append the reference package's directory (if installed elsewhere on sys.path) to pkg_path."""
    here = _os.path.dirname(_os.path.abspath(__file__))
    roots = [_os.environ["V2X_REFERENCE_ROOT"]] if "V2X_REFERENCE_ROOT" in _os.environ else []
    roots += [p for p in _sys.path if p]
    for root in roots:
        cand = _os.path.join(root, "coperception", *sub)
        if _os.path.isdir(cand) and not _os.path.abspath(cand).startswith(here) and cand not in pkg_path:
            if _os.path.exists(_os.path.join(root, "coperception", "__init__.py")):
                pkg_path.append(cand)
                return cand
    return None


_extend_with_reference(__path__)

"""Execute bench.py (our arm) or __graft_entry__.smoke() on a CPU host with every kernel a no-op (tests/dryrun_plugin.py):
the python of the whole measured path -- plan build, graph capture, the pipelined e2e loop, the detections leg, the
roofline accounting, the JSON line -- runs; timings and values are meaningless.

Usage: python tools/bench_dryrun.py [bench.py arguments ...]   |   python tools/bench_dryrun.py --smoke"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def main():
    import dryrun_plugin

    class _Cfg:
        pass
    dryrun_plugin.pytest_configure(_Cfg())
    os.chdir(ROOT)
    if "--smoke" in sys.argv:
        import __graft_entry__ as g
        try:
            g.smoke()
        except AssertionError as e:
            print("smoke() reached its numeric assertion: %s" % (str(e)[:80],))
        return 0
    sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
    try:
        runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
    except SystemExit as e:
        return int(e.code or 0)
    return 0


if __name__ == "__main__":
    sys.exit(main())

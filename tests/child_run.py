"""Run a selection of one test file in a child pytest process with a time limit (tests/test_gpu_zz_options.py).

A case that has never met the hardware may fault in ways that outlive the test: a sticky CUDA error poisons every later
test of the session, a kernel that never returns blocks it for good.  In a child process both end with the child: on the
time limit exactly the process group started here is killed, which tears its CUDA context down and frees the GPU."""
import os
import subprocess
import sys


def run_in_child(test_file, select, log_path, time_limit_s, env=None, marker=None):
    """Returns (rc, tail of the log); rc is the child's exit status or the string "time limit of N s"."""
    cmd = [sys.executable, "-m", "pytest", os.path.abspath(test_file), "-q", "-k", select, "-p", "no:cacheprovider", "-s"]
    if marker:
        cmd += ["-m", marker]
    os.makedirs(os.path.dirname(os.path.abspath(log_path)), exist_ok=True)
    with open(log_path, "w") as f:
        proc = subprocess.Popen(cmd, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                env=dict(os.environ, **(env or {})), stdout=f, stderr=subprocess.STDOUT, start_new_session=True)
        try:
            rc = proc.wait(timeout=time_limit_s)
        except subprocess.TimeoutExpired:
            os.killpg(proc.pid, 9)          # exactly the process group started above
            proc.wait()
            rc = "time limit of %d s" % time_limit_s
    with open(log_path) as f:
        return rc, f.read()[-3000:]


def ran_and_passed(rc, tail):
    """True iff the child exited 0 AND actually ran something (a selection that only skips is not a pass)."""
    last = [l for l in tail.splitlines() if l.strip()][-1] if tail.strip() else ""
    return rc == 0 and " passed" in last and "skipped" not in last

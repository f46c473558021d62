"""Static checks (AST) over the python of the product, the oracle, the tools and bench.py.  Most of this code only runs
on a GPU box -- some of it only under a multi-GPU launch -- so the CPU suite cannot execute it; these checks catch the
class of mistake a missing execution hides: a name bound in no enclosing scope, a library entry point that the header /
the ctypes table does not declare.  (tests/test_plan_dryrun.py holds the companion check on ``self.x`` reads.)"""
import ast
import builtins
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILTINS = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__class__"}


def _python_files():
    for top in ("v2x-sim_b200", "oracle", "tools", "tests", "bench.py", "__graft_entry__.py"):
        path = os.path.join(ROOT, top)
        if os.path.isfile(path):
            yield path
            continue
        for dp, _, fns in os.walk(path):
            if "_ref" in dp or "__pycache__" in dp:
                continue
            for fn in fns:
                if fn.endswith(".py"):
                    yield os.path.join(dp, fn)


def _bound_in_scope(scope):
    """Names bound directly in ``scope`` (nested function / class bodies contribute only their own name)."""
    out = set()

    def visit(node):
        for c in ast.iter_child_nodes(node):
            if isinstance(c, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                out.add(c.name)
                continue
            if isinstance(c, ast.Lambda):
                continue
            if isinstance(c, ast.Name) and isinstance(c.ctx, (ast.Store, ast.Del)):
                out.add(c.id)
            elif isinstance(c, (ast.Import, ast.ImportFrom)):
                out.update((a.asname or a.name).split(".")[0] for a in c.names)
            elif isinstance(c, ast.ExceptHandler) and c.name:
                out.add(c.name)
            elif isinstance(c, (ast.Global, ast.Nonlocal)):
                out.update(c.names)
            visit(c)
    visit(scope)
    if isinstance(scope, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
        a = scope.args
        out.update(x.arg for x in a.args + a.kwonlyargs + a.posonlyargs)
        out.update(x.arg for x in (a.vararg, a.kwarg) if x is not None)
    return out


def _undefined_names(path):
    tree = ast.parse(open(path).read())
    found = []

    def walk(scope, env):
        env = env | _bound_in_scope(scope)

        def visit(node, names):
            for c in ast.iter_child_nodes(node):
                if isinstance(c, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
                    walk(c, env)          # (a class body's own names are not visible inside its methods)
                elif isinstance(c, ast.ClassDef):
                    visit(c, names | _bound_in_scope(c))
                else:
                    if isinstance(c, ast.Name) and isinstance(c.ctx, ast.Load) and c.id not in names and c.id not in BUILTINS:
                        found.append("%s:%d %s" % (os.path.relpath(path, ROOT), c.lineno, c.id))
                    visit(c, names)
        visit(scope, env)
    walk(tree, set())
    return found


def test_no_name_is_read_that_no_enclosing_scope_binds():
    problems = [p for f in _python_files() for p in _undefined_names(f)]
    assert not problems, problems


def test_checker_sees_a_planted_mistake(tmp_path):
    f = tmp_path / "planted.py"
    f.write_text("import os\n\ndef f(a):\n    b = a + 1\n    return b + c_undefined + os.sep\n\n"
                 "class K:\n    x = 1\n    def m(self):\n        return x\n")
    assert [p.split()[-1] for p in _undefined_names(str(f))] == ["c_undefined", "x"]


def test_every_library_entry_point_named_in_python_is_declared():
    """``v2x_*`` identifiers in any python file are entry points of include/v2x_b200.h and rows of _lib.SYMBOLS (or the
    params struct / package name); header and table declare the same set."""
    import sys
    src = os.path.join(ROOT, "v2x-sim_b200")
    if src not in sys.path:
        sys.path.insert(0, src)
    from v2x_b200 import _lib
    table = {n for n, _, _ in _lib.SYMBOLS}
    header = set(re.findall(r"\b(v2x_[a-z0-9_]+)\s*\(", open(os.path.join(ROOT, "include", "v2x_b200.h")).read()))
    assert header == table, (sorted(header - table), sorted(table - header))
    allowed_prefixes = ("v2x_b200", "v2x_conv_params", "v2x_sim")
    stray = set()
    for f in _python_files():
        if os.path.basename(f) == "test_static_checks.py":
            continue
        for m in re.finditer(r"\b(v2x_[a-z0-9_]+)\b", open(f).read()):
            if m.group(1) not in table and not m.group(1).startswith(allowed_prefixes):
                stray.add((m.group(1), os.path.relpath(f, ROOT)))
    assert not stray, sorted(stray)


def test_gpu_suite_refusal_cases_pass_under_the_cpu_dry_run():
    """tools/gpu_suite_dryrun.py runs GPU test files on the CPU with every kernel a no-op.  The cases that assert a REFUSAL
    launch no kernel, so they must pass outright there -- a feature that starts to run what a GPU test expects to be
    refused (compressed training did) shows up here instead of stopping the round-end GPU run."""
    import subprocess
    import sys
    sel = ["tests/test_gpu_dropin.py::test_train_mode_is_refused_where_not_built",
           "tests/test_gpu_zz_options.py::test_when2com_refuses_what_the_reference_cannot_run",
           "tests/test_gpu_zz_options.py::test_training_refuses_fewer_than_32_compressed_channels"]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_suite_dryrun.py")] + sel, cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "3 passed" in r.stdout and "NON-NUMERIC failures: 0" in r.stdout, r.stdout[-2000:]

"""CPU: hygiene of bench.py that does not need a GPU — the reference arm must not pull the product into its process
(the driver records which .so files each arm maps), every workload is well formed, and `--impl`/`--config` parse."""
import ast
import importlib.util
import inspect
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _imports(node):
    out = set()
    for n in ast.walk(node):
        if isinstance(n, ast.Import):
            out |= {a.name for a in n.names}
        elif isinstance(n, ast.ImportFrom) and n.module:
            out |= {f"{n.module}.{a.name}" for a in n.names} | {n.module}
    return out


def test_reference_arm_does_not_import_the_product():
    tree = ast.parse((ROOT / "bench.py").read_text())
    classes = {n.name: n for n in tree.body if isinstance(n, ast.ClassDef)}
    ref = _imports(classes["ReferenceStep"])
    assert not any(m.startswith("materialrefgs_b200") for m in ref), ref
    base = {m for m in _imports(classes["StepBase"]) if m.startswith("materialrefgs_b200")}
    # the shared base may use the numpy/torch scene generators and, for MULTI-rank runs (our arm only), the view assignment
    assert base <= {"materialrefgs_b200", "materialrefgs_b200.synthetic", "materialrefgs_b200.parallel",
                    "materialrefgs_b200.parallel.assign_views_balanced"}, base
    top = {m for m in _imports(ast.Module(body=[n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))],
                                          type_ignores=[])) if m.startswith("materialrefgs_b200")}
    assert not top, top      # nothing of the product is imported at module level
    # neither materialrefgs_b200/__init__.py nor synthetic.py loads libmrgs.so
    for f in ("__init__.py", "synthetic.py"):
        src = (ROOT / "materialrefgs_b200" / f).read_text()
        assert "_lib" not in src and "ctypes" not in src, f


def test_workloads_are_well_formed():
    b = _bench()
    assert set(b.WORKLOADS) == {"C2", "C3", "C4", "C5", "C5-eval"}
    for name, wl in b.WORKLOADS.items():
        assert wl["mode"] in ("train", "eval") and wl["scaling"] in ("weak", "strong"), name
        assert (wl["batch_views"] is None) == (wl["scaling"] == "weak"), name
        assert wl["cube_res"] % (wl["min_res"]) == 0 and wl["metric"] and wl["text"], name
    c3 = b.WORKLOADS["C3"]
    assert (c3["P"], c3["W"], c3["H"], c3["S"], c3["cube_res"], c3["views_per_rank"]) == (1_000_000, 800, 800, 8, 512, 4)
    src = inspect.getsource(b.main)
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--config"):
        assert flag in src


def test_algorithmic_bytes_follow_the_survey_closed_forms():
    b = _bench()
    ab = b.algorithmic_bytes(P=1000, Pv=900, R=9000, N=640, S=8)
    assert ab["render_fwd"] == 9000 * (68 + 4 * 11) + 640 * 4 * 23
    assert ab["render_bwd"] == ab["render_fwd"] + 1000 * 4 * 26
    assert ab["sort"] == 24 * 9000

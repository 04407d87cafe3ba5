"""CPU: oracle/features_oracle.py against vectors produced by the reference's own Python functions
(tests/golden/make_golden_features.py): outputs and all nine parameter gradients."""
import glob
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import features_oracle as fo

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = sorted(glob.glob(str(ROOT / "tests" / "golden" / "features_*.npz")))
OUTS = ("scales", "rotations", "opacities", "features")


def test_golden_files_present():
    assert len(GOLDEN) >= 2


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_oracle_matches_reference_functions(path):
    z = np.load(path)
    p = {k: torch.from_numpy(z["in_" + k]).requires_grad_(True) for k, _ in fo.RAW_FIELDS}
    outs = fo.prepare_features(*[p[k] for k, _ in fo.RAW_FIELDS], torch.from_numpy(z["campos"]))
    for n, o in zip(OUTS, outs):
        assert np.abs(o.detach().numpy() - z["out_" + n]).max() <= 1e-6, n
    sum((o * torch.from_numpy(z["up_" + n])).sum() for n, o in zip(OUTS, outs)).backward()
    for k, _ in fo.RAW_FIELDS:
        ref = z["grad_" + k]
        err = np.abs(p[k].grad.numpy() - ref).max() / max(np.abs(ref).max(), 1e-30)
        assert err <= 1e-5, (k, err)


def test_product_eval_sh_matches_reference_function():
    """materialrefgs_b200/render.py eval_sh (host-side torch, used for the indirect light of duck-typed models) against
    the reference's own utils/sh_utils.py eval_sh at every degree (tests/golden/make_golden_sh.py)."""
    from materialrefgs_b200 import render
    z = np.load(Path(__file__).resolve().parent / "golden" / "eval_sh.npz")
    sh, dirs = torch.from_numpy(z["sh"]), torch.from_numpy(z["dirs"])
    for deg in range(4):
        got = render.eval_sh(deg, sh, dirs).numpy()
        assert np.abs(got - z[f"deg{deg}"]).max() <= 2e-6, deg

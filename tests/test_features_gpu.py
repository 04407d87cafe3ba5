"""GPU: fused per-surfel feature preparation (mrgs_surfel_features_*) through the C ABI against
(a) vectors produced by the reference's own Python functions and (b) the torch oracle at a ragged size."""
import glob
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import features_oracle as fo

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
GOLDEN = sorted(glob.glob(str(ROOT / "tests" / "golden" / "features_*.npz")))
OUTS = ("scales", "rotations", "opacities", "features")
OUT_ATOL = 2e-6       # fp32 exp / sigmoid / 16-term SH sums on two different devices
GRAD_RTOL = 1e-4      # relative to the tensor's max-norm


def _run(p, campos, ups):
    from materialrefgs_b200.features import surfel_features
    outs = surfel_features(campos, *[p[k] for k, _ in fo.RAW_FIELDS])
    sum((o * u).sum() for o, u in zip(outs, ups)).backward()
    return outs


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_against_reference_function_vectors(path):
    dev = torch.device("cuda:0")
    z = np.load(path)
    p = {k: torch.from_numpy(z["in_" + k]).to(dev).requires_grad_(True) for k, _ in fo.RAW_FIELDS}
    ups = [torch.from_numpy(z["up_" + n]).to(dev) for n in OUTS]
    outs = _run(p, torch.from_numpy(z["campos"]).to(dev), ups)
    for n, o in zip(OUTS, outs):
        assert np.abs(o.detach().cpu().numpy() - z["out_" + n]).max() <= OUT_ATOL, n
    for k, _ in fo.RAW_FIELDS:
        ref = z["grad_" + k]
        err = np.abs(p[k].grad.cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1e-30)
        assert err <= GRAD_RTOL, (k, err)


def test_against_oracle_ragged_size_and_model_shapes():
    """P not a multiple of the 128-surfel tile; indirect_dc / indirect_rest in the model's [P,1,3] / [P,15,3] shapes."""
    dev = torch.device("cuda:0")
    P = 100_003
    raw, campos = fo.synthetic_params(P, seed=7)
    raw["indirect_dc"] = raw["indirect_dc"].view(P, 1, 3)
    raw["indirect_rest"] = raw["indirect_rest"].view(P, 15, 3)
    g = torch.Generator().manual_seed(8)
    ups = [torch.randn(P, w, generator=g) for w in (2, 4, 1, 8)]
    pc = {k: v.clone().requires_grad_(True) for k, v in raw.items()}
    outs_c = fo.prepare_features(*[pc[k] for k, _ in fo.RAW_FIELDS], campos)
    sum((o * u).sum() for o, u in zip(outs_c, ups)).backward()
    pg = {k: v.clone().to(dev).requires_grad_(True) for k, v in raw.items()}
    outs_g = _run(pg, campos.to(dev), [u.to(dev) for u in ups])
    for n, a, b in zip(OUTS, outs_g, outs_c):
        assert (a.detach().cpu() - b.detach()).abs().max().item() <= OUT_ATOL, n
    for k, _ in fo.RAW_FIELDS:
        ref = pc[k].grad
        assert pg[k].grad.shape == ref.shape
        err = ((pg[k].grad.cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()
        assert err <= GRAD_RTOL, (k, err)


def test_rejects_cpu_and_bad_shapes():
    from materialrefgs_b200.features import surfel_features
    raw, campos = fo.synthetic_params(10, seed=1)
    with pytest.raises(RuntimeError):
        surfel_features(campos, *[raw[k] for k, _ in fo.RAW_FIELDS])
    dev = torch.device("cuda:0")
    rg = {k: v.to(dev) for k, v in raw.items()}
    rg["rotation"] = rg["rotation"][:, :3]
    with pytest.raises(RuntimeError):
        surfel_features(campos.to(dev), *[rg[k] for k, _ in fo.RAW_FIELDS])
    empty = {k: v[:0].to(dev) for k, v in raw.items()}
    outs = surfel_features(campos.to(dev), *[empty[k] for k, _ in fo.RAW_FIELDS])
    assert [tuple(o.shape) for o in outs] == [(0, 2), (0, 4), (0, 1), (0, 8)]

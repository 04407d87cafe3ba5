"""GPU: materialrefgs_b200/surfel_model.py with the surfels resident on the device — densification
statistics through mrgs_densify_stats, clone / split / prune, Adam-state surgery and reset_opacity0 against the vectors
of the reference's own GaussianModel methods (tests/golden/densify_*.npz, produced on the CPU by
tests/golden/make_golden_densify.py). Selections, counts and integer statistics must be EQUAL; floating-point fields may
differ by the device's exp/log rounding (1 ulp), so they are compared at 1e-6 relative."""
import glob
from pathlib import Path

import numpy as np
import pytest
import torch

from materialrefgs_b200 import surfel_model as sm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = Path(__file__).resolve().parent.parent
GOLDEN = sorted(glob.glob(str(ROOT / "tests" / "golden" / "densify_*.npz")))


def _store(z):
    fields = {n: torch.from_numpy(z["before_p_" + f.group]).to(DEV) for n, f in sm.FIELDS.items()}
    st = sm.SurfelStore(fields, percent_dense=float(z["percent_dense"]))
    for n, f in sm.FIELDS.items():
        if "before_m_" + f.group in z.files:
            st.optimizer.state[st[n]] = {"step": torch.tensor(2.0),
                                        "exp_avg": torch.from_numpy(z["before_m_" + f.group]).to(DEV),
                                        "exp_avg_sq": torch.from_numpy(z["before_v_" + f.group]).to(DEV)}
    return st


def _close(a, b, what):
    a = a.detach().cpu().numpy()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.allclose(a, b, rtol=1e-6, atol=1e-7), (what, float(np.abs(a - b).max()))


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_densification_on_device_matches_reference_methods(path):
    from materialrefgs_b200 import parallel
    z = np.load(path)
    st = _store(z)
    assert st.device.type == "cuda"
    P = int(z["P"])
    # the statistics of every view through the arena kernel (mrgs_densify_stats), as a view-sharded step produces them
    arena = parallel.GradArena.create(P, DEV)
    for g, filt, radii in zip(z["view_grads"], z["view_filters"], z["view_radii"]):
        r = torch.from_numpy(radii).to(DEV).to(torch.int32)
        r = torch.where(torch.from_numpy(filt).to(DEV), r, torch.zeros_like(r))
        arena.accumulate_view({}, torch.from_numpy(g).to(DEV).contiguous(), r)
    st.load_reduced_stats(arena.stats, arena.max_radii)
    _close(st.xyz_gradient_accum, z["stats_xyz_gradient_accum"], "xyz_gradient_accum")
    assert np.array_equal(st.denom.cpu().numpy(), z["stats_denom"])
    assert np.array_equal(st.max_radii2D.cpu().numpy(), z["stats_max_radii2D"])
    # the split offsets come from a HOST generator seeded like the golden run: same stream on every rank and device
    gen = torch.Generator()
    gen.manual_seed(int(z["seed"]) + 300)
    mss = int(z["max_screen_size"])
    st.densify_and_prune(float(z["max_grad"]), float(z["min_opacity"]), float(z["extent"]), None if mss < 0 else mss,
                         generator=gen)
    assert st.num_points == z["after_p_xyz"].shape[0] != P
    for n, f in sm.FIELDS.items():
        assert st[n].is_cuda
        _close(st[n], z["after_p_" + f.group], n)
        state = st.optimizer.state.get(st[n], None)
        if "after_m_" + f.group in z.files:
            assert state["exp_avg"].is_cuda
            _close(state["exp_avg"], z["after_m_" + f.group], n + ".exp_avg")
            _close(state["exp_avg_sq"], z["after_v_" + f.group], n + ".exp_avg_sq")
        else:
            assert state is None or "exp_avg" not in state
    for k in ("xyz_gradient_accum", "denom", "max_radii2D"):
        assert np.array_equal(getattr(st, k).cpu().numpy(), z["after_" + k]), k
    st.reset_opacity0()
    _close(st["opacity"], z["reset_p_opacity"], "reset opacity")
    _close(st.optimizer.state[st["opacity"]]["exp_avg"], z["reset_m_opacity"], "reset moment")
    # the rebuilt parameters feed the rasterizer and the optimizer steps on the device
    for p in st.params.values():
        p.grad = torch.ones_like(p)
    st.optimizer.step()
    torch.cuda.synchronize()

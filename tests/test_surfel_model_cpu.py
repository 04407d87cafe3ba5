"""CPU: materialrefgs_b200/surfel_model.py — the .ply format and the densification bookkeeping against
vectors produced by the reference's own GaussianModel methods (tests/golden/make_golden_densify.py)."""
import glob
from pathlib import Path

import numpy as np
import pytest
import torch

from materialrefgs_b200 import surfel_model as sm

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = sorted(glob.glob(str(ROOT / "tests" / "golden" / "densify_*.npz")))


def random_fields(P, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {n: torch.randn((P, *f.shape), generator=g) for n, f in sm.FIELDS.items()}


def test_attribute_order_is_the_reference_one():
    names = sm.construct_list_of_attributes()
    # scene/gaussian_model.py:462-487 with f_dc 1x3, f_rest 15x3, ind_dc 1x3, ind_rest 15x3, ind_asg 32x5
    assert names[:9] == ["x", "y", "z", "nx", "ny", "nz", "nx2", "ny2", "nz2"]
    assert names[9:12] == ["f_dc_0", "f_dc_1", "f_dc_2"] and names[12] == "f_rest_0" and names[56] == "f_rest_44"
    assert names[57:60] == ["ind_dc_0", "ind_dc_1", "ind_dc_2"] and names[60] == "ind_rest_0" and names[104] == "ind_rest_44"
    assert names[105] == "ind_asg_0" and names[264] == "ind_asg_159"
    assert names[265:269] == ["opacity", "refl_strength", "metalness", "roughness"]
    assert names[269:275] == ["ori_color_0", "ori_color_1", "ori_color_2", "diffuse_color_0", "diffuse_color_1", "diffuse_color_2"]
    assert names[275:] == ["scale_0", "scale_1", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert len(names) == 281


def test_ply_round_trip_and_header(tmp_path):
    f = random_fields(37, seed=3)
    path = tmp_path / "point_cloud" / "iteration_7" / "point_cloud.ply"
    sm.save_ply(path, f)
    raw = path.read_bytes()
    head, _, body = raw.partition(b"end_header\n")
    lines = head.decode().splitlines()
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 37"]
    assert lines[3:] == [f"property float {n}" for n in sm.construct_list_of_attributes()]
    assert len(body) == 37 * 281 * 4
    table = np.frombuffer(body, dtype="<f4").reshape(37, 281)
    # channel-major SH columns: f_rest_k = features_rest[:, k % 15, k // 15]  (transpose(1, 2).flatten, :495)
    assert np.array_equal(table[:, 12 + 17], f["features_rest"][:, 2, 1].numpy())
    assert np.array_equal(table[:, 105 + 33], f["indirect_asg"][:, 1, 1].numpy())   # [P,32,5] -> [P,5,32]
    back = sm.load_ply(path)
    for n in sm.FIELDS:
        assert back[n].shape == tuple(f[n].shape) and np.array_equal(back[n], f[n].numpy()), n
    # attribute lookup is by name: a file with shuffled columns loads the same
    order = np.random.default_rng(0).permutation(281)
    names = sm.construct_list_of_attributes()
    shuffled = tmp_path / "shuffled.ply"
    hdr = ["ply", "format binary_little_endian 1.0", "comment written by a test", "element vertex 37"]
    hdr += [f"property float {names[i]}" for i in order] + ["end_header"]
    shuffled.write_bytes(("\n".join(hdr) + "\n").encode() + np.ascontiguousarray(table[:, order]).tobytes())
    back2 = sm.load_ply(shuffled)
    for n in sm.FIELDS:
        assert np.array_equal(back2[n], f[n].numpy()), n
    with pytest.raises(ValueError):
        sm.load_ply(path, max_sh_degree=2)
    store = sm.SurfelStore.from_ply(path)
    assert store.num_points == 37 and torch.equal(store["rotation"].detach(), f["rotation"])


def store_from_golden(z, prefix="before_"):
    fields = {n: torch.from_numpy(z[prefix + "p_" + f.group]) for n, f in sm.FIELDS.items()}
    st = sm.SurfelStore(fields, percent_dense=float(z["percent_dense"]))
    for n, f in sm.FIELDS.items():
        if prefix + "m_" + f.group in z.files:
            st.optimizer.state[st[n]] = {"step": torch.tensor(2.0), "exp_avg": torch.from_numpy(z[prefix + "m_" + f.group]).clone(),
                                        "exp_avg_sq": torch.from_numpy(z[prefix + "v_" + f.group]).clone()}
    return st


def test_golden_files_present():
    assert len(GOLDEN) >= 2


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_densification_matches_reference_methods(path):
    z = np.load(path)
    st = store_from_golden(z)
    for g, filt, radii in zip(z["view_grads"], z["view_filters"], z["view_radii"]):
        st.add_densification_stats(torch.from_numpy(g), torch.from_numpy(filt), torch.from_numpy(radii))
    assert np.array_equal(st.xyz_gradient_accum.numpy(), z["stats_xyz_gradient_accum"])
    assert np.array_equal(st.denom.numpy(), z["stats_denom"])
    assert np.array_equal(st.max_radii2D.numpy(), z["stats_max_radii2D"])
    torch.manual_seed(int(z["seed"]) + 300)
    mss = int(z["max_screen_size"])
    st.densify_and_prune(float(z["max_grad"]), float(z["min_opacity"]), float(z["extent"]), None if mss < 0 else mss)
    assert st.num_points == z["after_p_xyz"].shape[0] != int(z["P"])
    for n, f in sm.FIELDS.items():
        assert np.array_equal(st[n].detach().numpy(), z["after_p_" + f.group]), n
        state = st.optimizer.state.get(st[n], None)
        if "after_m_" + f.group in z.files:
            assert np.array_equal(state["exp_avg"].numpy(), z["after_m_" + f.group]), n
            assert np.array_equal(state["exp_avg_sq"].numpy(), z["after_v_" + f.group]), n
        else:
            assert state is None or "exp_avg" not in state
    for k in ("xyz_gradient_accum", "denom", "max_radii2D"):
        assert np.array_equal(getattr(st, k).numpy(), z["after_" + k]), k
    st.reset_opacity0()
    assert np.array_equal(st["opacity"].detach().numpy(), z["reset_p_opacity"])
    assert np.array_equal(st.optimizer.state[st["opacity"]]["exp_avg"].numpy(), z["reset_m_opacity"])
    # the optimizer still steps on the rebuilt parameters
    for p in st.params.values():
        p.grad = torch.ones_like(p)
    st.optimizer.step()


def test_reduced_stats_equal_per_view_accumulation():
    P = 50
    g = torch.Generator().manual_seed(1)
    a, b = sm.SurfelStore(random_fields(P)), sm.SurfelStore(random_fields(P))
    stats, mx = torch.zeros(P, 2), torch.zeros(P, dtype=torch.int32)
    for _ in range(3):
        grad = torch.randn(P, 3, generator=g)
        radii = (torch.randint(0, 30, (P,), generator=g) * (torch.rand(P, generator=g) > 0.4)).int()
        filt = radii > 0
        a.add_densification_stats(grad, filt, radii)
        stats[filt, 0] += grad[filt].norm(dim=-1)
        stats[filt, 1] += 1
        mx = torch.maximum(mx, radii)
    b.load_reduced_stats(stats, mx)
    assert torch.allclose(a.xyz_gradient_accum, b.xyz_gradient_accum) and torch.equal(a.denom, b.denom)
    assert torch.equal(a.max_radii2D, b.max_radii2D)


def test_restore_reads_a_checkpoint_written_by_the_reference():
    """tests/golden/chkpnt_reference_small.pth = torch.save((GaussianModel.capture(), iteration)) by the reference's own
    code (make_golden_densify.py): parameters, statistics, Adam moments and learning rates come back; capture() gives
    the same tuple layout."""
    model_args, iteration = torch.load(ROOT / "tests" / "golden" / "chkpnt_reference_small.pth", weights_only=False)
    assert iteration == 1234 and len(model_args) == 22
    st = sm.SurfelStore.restore(model_args)
    assert st.active_sh_degree == 2 and st.spatial_lr_scale == 3.5 and st.num_points == 12
    names = ("xyz", "refl_strength", "metalness", "roughness", "ori_color", "diffuse_color", "features_dc", "features_rest",
             "indirect_dc", "indirect_rest", "indirect_asg", "scaling", "rotation", "opacity", "normal1", "normal2")
    for i, n in enumerate(names):
        if n != "indirect_asg":                      # re-created as zeros by the reference's restore (:169) as well
            assert torch.equal(st[n].detach(), model_args[1 + i].detach()), n
    assert not st["indirect_asg"].any() and st["indirect_asg"].shape == (12, 32, 5)
    assert torch.equal(st.max_radii2D, model_args[17]) and torch.equal(st.xyz_gradient_accum, model_args[18])
    assert torch.equal(st.denom, model_args[19])
    opt = model_args[20]
    by_name = {g["name"]: g for g in opt["param_groups"]}
    assert {"env", "env2"} <= set(by_name)            # the reference's optimizer carries the environment maps too
    for n, f in sm.FIELDS.items():
        saved = by_name[f.group]
        group = st._group(n)
        assert group["lr"] == saved["lr"]
        state = opt["state"].get(saved["params"][0])
        if state is None:
            assert st.optimizer.state.get(st[n]) is None
        else:
            assert torch.equal(st.optimizer.state[st[n]]["exp_avg"], state["exp_avg"]), n
            assert torch.equal(st.optimizer.state[st[n]]["exp_avg_sq"], state["exp_avg_sq"]), n
    again = st.capture()
    assert len(again) == 22 and again[0] == 2 and again[21] == 3.5
    for i, n in enumerate(names):
        assert again[1 + i] is st[n]
    # capture() writes the optimizer state in the REFERENCE's layout: an optimizer built like training_setup
    # (gaussian_model.py:422-447: 18 groups, env/env2 in 7th/8th place) loads it positionally, moments on the right tensors
    opt2 = again[20]
    assert [g["name"] for g in opt2["param_groups"]] == [g["name"] for g in opt["param_groups"]] == list(sm.REFERENCE_GROUP_ORDER)
    by_group = {f.group: [torch.nn.Parameter(st[n].detach().clone())] for n, f in sm.FIELDS.items()}
    for n in ("env", "env2"):                          # as many parameters as the saved EnvLight module had
        by_group[n] = [torch.nn.Parameter(torch.zeros(6, 4, 4, 3)) for _ in by_name[n]["params"]]
    like_ref = torch.optim.Adam([{"params": by_group[n], "lr": 0.0, "name": n} for n in sm.REFERENCE_GROUP_ORDER], lr=0.0, eps=1e-15)
    like_ref.load_state_dict(opt2)
    for n, f in sm.FIELDS.items():
        mine = st.optimizer.state.get(st[n])
        theirs = like_ref.state.get(by_group[f.group][0])
        assert (mine is None or not mine) == (theirs is None or not theirs), n
        if mine:
            assert torch.equal(mine["exp_avg"], theirs["exp_avg"]) and torch.equal(mine["exp_avg_sq"], theirs["exp_avg_sq"]), n
    for n in ("env", "env2"):                          # carried through restore() -> capture() untouched
        for k, i in enumerate(by_name[n]["params"]):
            src, dst = opt["state"].get(i), like_ref.state.get(by_group[n][k])
            assert (not src) == (not dst)
            if src:
                assert torch.equal(src["exp_avg"], dst["exp_avg"])
    # a store that never saw the environment-map groups cannot write a reference-loadable tuple: it says so
    fresh = sm.SurfelStore({n: st[n].detach().clone() for n in sm.FIELDS})
    with pytest.raises(ValueError):
        fresh.capture()
    hyper = {k: v for k, v in by_name["env"].items() if k != "params"}
    assert len(fresh.capture(env_groups={"env": (hyper, [None]), "env2": (dict(hyper, name="env2"), [None])})[20]["param_groups"]) == 18
    # the restored optimizer steps
    for p in st.params.values():
        p.grad = torch.ones_like(p)
    st.optimizer.step()


def test_learning_rate_schedule_matches_reference_function():
    z = np.load(ROOT / "tests" / "golden" / "lr_schedule.npz")
    f1 = sm.get_expon_lr_func(lr_init=1.6e-4 * 3.5, lr_final=1.6e-6 * 3.5, lr_delay_mult=0.01, max_steps=30000)
    f2 = sm.get_expon_lr_func(lr_init=0.01, lr_final=0.001, lr_delay_steps=1000, lr_delay_mult=0.1, max_steps=20000)
    assert np.array_equal(np.array([f1(int(s)) for s in z["steps"]]), z["plain"])
    assert np.array_equal(np.array([f2(int(s)) for s in z["steps"]]), z["delayed"])
    st = sm.SurfelStore(random_fields(5), xyz_schedule=dict(lr_init=1.6e-4, lr_final=1.6e-6, lr_delay_mult=0.01, max_steps=30000))
    assert st.update_learning_rate(0) == pytest.approx(1.6e-4) and st._group("xyz")["lr"] == pytest.approx(1.6e-4)
    assert st.update_learning_rate(30000) == pytest.approx(1.6e-6)
    st.oneupSHdegree(); st.oneupSHdegree(); st.oneupSHdegree(); st.oneupSHdegree()
    assert st.active_sh_degree == 3


# ---- invariants over arbitrary densification sequences (hypothesis) ----------------------------------------------
from hypothesis import given, settings, strategies as st_  # noqa: E402


@settings(max_examples=15, deadline=None)
@given(seed=st_.integers(0, 10_000), P=st_.integers(1, 60), max_grad=st_.floats(1e-5, 5e-3), min_opacity=st_.floats(0.0, 0.6),
       extent=st_.floats(0.5, 6.0), screen=st_.one_of(st_.none(), st_.integers(1, 40)), rounds=st_.integers(1, 3))
def test_store_invariants_hold_after_any_densification_sequence(seed, P, max_grad, min_opacity, extent, screen, rounds):
    """Whatever gets cloned, split or pruned (including everything or nothing): every field, both Adam moments and the three
    statistics keep one row per surfel, the optimizer still steps, and a .ply round trip returns the same cloud."""
    g = torch.Generator().manual_seed(seed)
    fields = random_fields(P, seed)
    fields["scaling"] = torch.log(0.02 * torch.exp(0.8 * torch.randn((P, 2), generator=g)))
    store = sm.SurfelStore(fields, lrs={f.group: 1e-3 for f in sm.FIELDS.values()})
    for _ in range(rounds):
        n = store.num_points
        for p in store.params.values():
            p.grad = 0.01 * torch.randn(p.shape, generator=g)
        store.optimizer.step()
        if n:
            grad = 1e-3 * torch.randn(n, 3, generator=g)
            radii = (torch.randint(0, 50, (n,), generator=g) * (torch.rand(n, generator=g) > 0.3)).int()
            store.add_densification_stats(grad, radii > 0, radii)
        store.densify_and_prune(max_grad, min_opacity, extent, screen, generator=g)
        n = store.num_points
        for name, f in sm.FIELDS.items():
            p = store[name]
            assert p.shape == (n, *f.shape) and p.requires_grad, name
            state = store.optimizer.state.get(p)
            if state:
                assert state["exp_avg"].shape == p.shape and state["exp_avg_sq"].shape == p.shape, name
            assert store._group(name)["params"][0] is p
        assert store.xyz_gradient_accum.shape == (n, 1) and store.denom.shape == (n, 1) and store.max_radii2D.shape == (n,)
        assert not store.xyz_gradient_accum.any() and not store.max_radii2D.any()      # reset by densification_postfix
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "cloud.ply"
        store.save_ply(path)
        back = sm.load_ply(path)
    for name in sm.FIELDS:
        assert np.array_equal(back[name], store[name].detach().numpy()), name


def test_reflection_aware_resets_match_reference_methods():
    """reset_opacity0 -> reset_refl -> reset_opacity1 -> dist_color -> reset_scale as train_refnerf.py:1439-1455 chains
    them, without and with an exclusion mask, against the reference's own methods (tests/golden/resets.npz): parameters
    bit-equal, Adam moments zeroed the same way."""
    z = np.load(ROOT / "tests" / "golden" / "resets.npz")
    st = store_from_golden(_with_percent_dense(z))
    msk = torch.from_numpy(z["mask"])
    for tag, mask in (("plain", None), ("masked", msk)):
        st.reset_opacity0()
        st.reset_refl(exclusive_msk=mask, rst_value=0.1 if mask is not None else None)
        st.reset_opacity1(exclusive_msk=mask)
        torch.manual_seed(int(z["seed"]) + 7)
        st.dist_color(exclusive_msk=mask)
        st.reset_scale(exclusive_msk=mask)
        for name, group in (("opacity", "opacity"), ("refl_strength", "refl_strength"), ("features_dc", "f_dc"), ("scaling", "scaling")):
            assert np.array_equal(st[name].detach().numpy(), z[f"{tag}_p_{group}"]), (tag, name)
            state = st.optimizer.state[st[name]]
            assert np.array_equal(state["exp_avg"].numpy(), z[f"{tag}_m_{group}"]), (tag, name)
            assert np.array_equal(state["exp_avg_sq"].numpy(), z[f"{tag}_v_{group}"]), (tag, name)


class _with_percent_dense:
    """An npz view that also answers `percent_dense` (store_from_golden reads it from the densify vectors)."""
    def __init__(self, z):
        self.z, self.files = z, list(z.files) + ["percent_dense"]

    def __getitem__(self, k):
        return np.float64(0.01) if k == "percent_dense" else self.z[k]

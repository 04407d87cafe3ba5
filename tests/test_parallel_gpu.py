"""GPU: the device side of the view-sharded step — the fused densification-statistics kernel
(mrgs_densify_stats) against the torch expression of scene/gaussian_model.py:1059-1061, and autograd
accumulating the rasterizer's gradients in place into a bound GradArena."""
import pytest
import torch

from materialrefgs_b200 import parallel, synthetic

pytestmark = pytest.mark.gpu


def test_densify_stats_kernel_matches_torch():
    dev = torch.device("cuda:0")
    P = 100_003
    g = torch.Generator().manual_seed(5)
    arena = parallel.GradArena.create(P, dev)
    stats = torch.zeros(P, 2)
    mx = torch.zeros(P, dtype=torch.int32)
    for _ in range(3):
        vg = torch.randn(P, 3, generator=g)
        radii = torch.randint(-3, 40, (P,), generator=g, dtype=torch.int32).clamp_min(0)
        arena.accumulate_view({}, vg.to(dev), radii.to(dev))
        vis = radii > 0
        stats[:, 0] += torch.linalg.norm(vg, dim=-1) * vis
        stats[:, 1] += vis.float()
        mx = torch.maximum(mx, radii)
    assert torch.allclose(arena.stats.cpu(), stats, rtol=1e-6, atol=1e-6)
    assert torch.equal(arena.max_radii.cpu(), mx)
    assert float(arena.grads.abs().max()) == 0.0     # the statistics tail does not spill into the gradients


def test_rasterizer_grads_accumulate_in_place_into_bound_arena():
    from materialrefgs_b200.diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    dev = torch.device("cuda:0")
    P, S, W, H = 5000, 8, 160, 120
    cloud = synthetic.make_cloud(P, S=S, seed=3).to(dev)
    names = ("means3D", "scales", "rotations", "opacities", "shs", "features")
    leaves = {k: getattr(cloud, k).clone().requires_grad_(True) for k in names}
    free = {k: getattr(cloud, k).clone().requires_grad_(True) for k in names}
    arena = parallel.GradArena.create(P, dev)
    arena.bind(leaves)
    arena.zero_()
    bg = torch.zeros(3, device=dev)
    for v in (1, 4):
        cam = synthetic.orbit_camera(v, 8, W, H).to(dev)
        rs = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, cam.world_view_transform,
                                           cam.full_proj_transform, 3, cam.camera_center, False, False)
        for L in (leaves, free):
            m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
            _, color, feat, radii, allmap = GaussianRasterizer(rs)(
                means3D=L["means3D"], means2D=m2, opacities=L["opacities"], shs=L["shs"], features=L["features"],
                scales=L["scales"], rotations=L["rotations"])
            (color.sum() + feat.square().sum() + allmap[1].sum()).backward()
    assert arena.bound(leaves)
    for k in names:   # the backward's atomics make two runs equal only up to summation order
        ref = free[k].grad.reshape(P, -1)
        assert float((arena.views[k] - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-12, k


def test_grad_sink_fused_accumulation_matches_autograd():
    """GaussianRasterizer(..., grad_sink=arena.views): the per-surfel backward ADDS every view's parameter
    gradients into the arena itself (MrgsBackwardArgs.accumulate); result == autograd accumulation, the leaves'
    bound .grad sees it, means2D still gets its per-view gradient, non-leaf inputs are refused."""
    from materialrefgs_b200.diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    dev = torch.device("cuda:0")
    P, S, W, H = 6000, 8, 176, 128
    cloud = synthetic.make_cloud(P, S=S, seed=4).to(dev)
    names = ("means3D", "scales", "rotations", "opacities", "shs", "features")
    sunk = {k: getattr(cloud, k).clone().requires_grad_(True) for k in names}
    free = {k: getattr(cloud, k).clone().requires_grad_(True) for k in names}
    arena = parallel.GradArena.create(P, dev)
    arena.bind(sunk)
    arena.zero_()
    bg = torch.zeros(3, device=dev)
    m2_grads = {}
    for v in (0, 3, 6):
        cam = synthetic.orbit_camera(v, 8, W, H).to(dev)
        rs = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, cam.world_view_transform,
                                           cam.full_proj_transform, 3, cam.camera_center, False, False)
        for tag, L, sink in (("sunk", sunk, arena.views), ("free", free, None)):
            m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
            _, color, feat, radii, allmap = GaussianRasterizer(rs, grad_sink=sink)(
                means3D=L["means3D"], means2D=m2, opacities=L["opacities"], shs=L["shs"], features=L["features"],
                scales=L["scales"], rotations=L["rotations"])
            (color.sum() + feat.square().sum() + allmap[1].sum() + allmap[0].sum()).backward()
            m2_grads[tag] = m2.grad
        assert float((m2_grads["sunk"] - m2_grads["free"]).abs().max()) <= 1e-5 * float(m2_grads["free"].abs().max())
    assert arena.bound(sunk)
    for k in names:
        ref = free[k].grad.reshape(P, -1)
        assert float((arena.views[k] - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-12, k
        assert sunk[k].grad.data_ptr() == arena.views[k].data_ptr()
    # a computed (non-leaf) input cannot bypass autograd
    with pytest.raises(RuntimeError):
        GaussianRasterizer(rs, grad_sink=arena.views)(
            means3D=sunk["means3D"], means2D=None, opacities=torch.sigmoid(sunk["opacities"]), shs=sunk["shs"],
            features=sunk["features"], scales=sunk["scales"], rotations=sunk["rotations"])

"""EnvLight.build_mips as precomputed sparse operators ("prefilter plans"), backed by csrc/prefilter.cu.

The reference runs `build_mips()` every training iteration (train_refnerf.py:1155-1163 -> scene/light.py:72-86):
cubemap_mip down to min_res, `ru.diffuse_cubemap` on the smallest level, `ru.specular_cubemap(level, roughness,
cutoff)` on every level (scene/renderutils/ops.py:391-458). For a fixed (res, roughness, cutoff) the last two are
fixed linear maps; `plan_pair()` builds them once per key and device (like the reference caches `__ndfBounds`,
ops.py:428-443) and `MipChain` applies a whole chain in two launches forward (mip pyramid + one gather over all
levels and the diffuse map) and 1 + (levels-1) launches backward (one gather with the transposed operators + the
reference's non-adjoint cubemap_mip backward from the coarsest level down).

HBM: a plan stores one float per (destination texel, tap) per orientation — 2 x 4.7 GB for the 6x512^2, 6-level
chain of BASELINE.json's C1/C3, 2 x 0.05 GB for the reference's default 128^2 chain. Plans beyond
MRGS_PREFILTER_BUDGET_GB (default 40) are refused with PrefilterTooLarge; `cubemap.specular_cubemap` then uses the
direct per-texel kernels of csrc/cubemap.cu.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


class PrefilterTooLarge(RuntimeError):
    pass


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def ndf_cutoff_costheta(roughness: float, cutoff: float) -> float:
    """cos of the cone angle that keeps `cutoff` of the GGX lobe's energy: the host-side search of
    __ndfBounds (scene/renderutils/ops.py:428-441), same 1e6-sample cumsum in float64."""
    def ndfGGX(alphaSqr, costheta):
        costheta = np.clip(costheta, 0.0, 1.0)
        d = (costheta * alphaSqr - costheta) * costheta + 1.0
        return alphaSqr / (d * d * np.pi)
    nSamples = 1000000
    costheta = np.cos(np.linspace(0, np.pi / 2.0, nSamples))
    D = np.cumsum(ndfGGX(roughness ** 4, costheta))
    idx = np.argmax(D >= D[..., -1] * cutoff)
    return float(costheta[idx])


_ct_cache: dict = {}


def cutoff_costheta(roughness: float, cutoff: float) -> float:
    key = (float(roughness), float(cutoff))
    if key not in _ct_cache:
        _ct_cache[key] = ndf_cutoff_costheta(*key)
    return _ct_cache[key]


def specular_bounds(res: int, costheta_cutoff: float, device) -> torch.Tensor:
    """int32 [6,res,res,6,4] = (xmin,xmax,ymin,ymax) per source face: SpecularBoundsKernel, c_src/cubemap.cu:183-246."""
    lib = _lib.load()
    bounds = torch.empty((6, res, res, 6, 4), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.mrgs_specular_bounds(res, float(costheta_cutoff), bounds.data_ptr(), _stream(device)),
                   "mrgs_specular_bounds")
    return bounds


@dataclass
class Plan:
    kind: int
    res: int
    rows_per_lane: int
    patch_width: int
    patch_seg_begin: torch.Tensor
    patch_slot_begin: torch.Tensor
    seg_desc: torch.Tensor
    spans: torch.Tensor
    weights: torch.Tensor
    taps: int
    rows: int               # weight rows of 32 * rows_per_lane floats
    wsum: torch.Tensor | None = None   # [texels] un-normalised weight sum (specular kinds)

    def struct(self) -> _lib.PrefilterPlan:
        s = _lib.PrefilterPlan()
        s.res, s.rows_per_lane, s.patch_width = self.res, self.rows_per_lane, self.patch_width
        s.patch_seg_begin, s.patch_slot_begin = self.patch_seg_begin.data_ptr(), self.patch_slot_begin.data_ptr()
        s.seg_desc, s.spans, s.weights = self.seg_desc.data_ptr(), self.spans.data_ptr(), self.weights.data_ptr()
        return s

    @property
    def patches(self) -> int:
        return self.patch_seg_begin.numel() - 1

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.patch_seg_begin, self.patch_slot_begin, self.seg_desc,
                                                           self.spans, self.weights))


_table_cache: dict = {}
_plan_cache: dict = {}
_shape_cache: dict = {}
_plan_bytes: dict = {}


def budget_bytes() -> int:
    return int(float(os.environ.get("MRGS_PREFILTER_BUDGET_GB", "40")) * (1 << 30))


def texel_table(res: int, device) -> torch.Tensor:
    key = (res, str(device))
    if key not in _table_cache:
        lib = _lib.load()
        tab = torch.empty((6, res, res, 4), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _lib.check(lib.mrgs_prefilter_texel_table(res, tab.data_ptr(), _stream(device)), "mrgs_prefilter_texel_table")
        _table_cache[key] = tab
    return _table_cache[key]


SHAPES = ((32, 1), (32, 2), (16, 1), (16, 2), (8, 1), (8, 2))     # (patch width in lanes, texel rows per lane)


def _count(kind, res, shape, roughness, ct, tab, bounds, wsum, full_search, device):
    """Count pass for one patch shape: (args, per-patch counts [3, patches], (segments, weight rows, taps))."""
    lib = _lib.load()
    PW, G = shape
    patches = lib.mrgs_prefilter_patch_count(res, G, PW)
    if patches < 0:
        raise RuntimeError(f"prefilter plan: patch shape {PW}x{G} does not tile a {res}^2 face")
    counts = torch.zeros((3, patches), dtype=torch.int32, device=device)
    a = _lib.PrefilterBuildArgs()
    a.kind, a.res, a.rows_per_lane, a.patch_width, a.full_search = kind, res, G, PW, int(full_search)
    a.roughness, a.costheta_cutoff = float(roughness), float(ct)
    a.texel_table = tab.data_ptr()
    a.bounds = None if bounds is None else bounds.data_ptr()
    a.wsum = None if wsum is None else wsum.data_ptr()
    a.seg_count, a.slot_count, a.tap_count = (counts[i].data_ptr() for i in range(3))
    with torch.cuda.device(device):
        _lib.check(lib.mrgs_prefilter_plan_count(C.byref(a), _stream(device)), "mrgs_prefilter_plan_count")
    sums = counts.to(torch.int64).cumsum(1)
    totals = tuple(int(v) for v in sums[:, -1].tolist())           # one-time host sync
    return a, sums, totals


def choose_shape(res, roughness, ct, tab, bounds, device):
    """The patch shape with the fewest padded weight slots for this level (exact counts from the cheap,
    membership-only count pass). MRGS_PREFILTER_SHAPE=WxG forces one (falls back to 32x1 where it does not tile)."""
    lib = _lib.load()
    forced = os.environ.get("MRGS_PREFILTER_SHAPE") or (
        "32x2" if os.environ.get("MRGS_PREFILTER_ROWS") == "2" else "32x1" if os.environ.get("MRGS_PREFILTER_ROWS") == "1" else None)
    if forced:
        PW, G = (int(v) for v in forced.lower().split("x"))
        return ((PW, G) if lib.mrgs_prefilter_patch_count(res, G, PW) >= 0 else (32, 1)), True
    # two rows per lane halve the source-texel reads per tap, which is what keeps the gather HBM-bound instead of
    # L1-bound (measured on the 6x512^2 chain: 1.16-1.35 ms with G = 2 against 1.39-1.87 ms with G = 1 although G = 2
    # pads 4-12 % more slots, profiles/r02_prefilter.md): among the shapes that tile this face take G = 2 if any does
    cands = [sh for sh in SHAPES if lib.mrgs_prefilter_patch_count(res, sh[1], sh[0]) >= 0]
    if any(sh[1] == 2 for sh in cands):
        cands = [sh for sh in cands if sh[1] == 2]
    # ... and weigh the padded slots with the per-byte cost measured for each width on the 6x512^2 chain (all levels in
    # one shape: 8 lanes wide streams at 5.8 TB/s, 16 at 5.25, 32 at 4.9; profiles/r02_prefilter.md)
    width_cost = {8: 1.0, 16: 1.11, 32: 1.19}
    best, best_score = (32, 1), None
    for shape in cands:
        _, _, (_, rows, _) = _count(_lib.PREFILTER_SPECULAR, res, shape, roughness, ct, tab, bounds, None, False, device)
        score = rows * shape[1] * width_cost[shape[0]]
        if best_score is None or score < best_score:
            best, best_score = shape, score
    return best, False


def _build(kind, res, shape, roughness, ct, tab, bounds, wsum, full_search, device) -> Plan:
    lib = _lib.load()
    PW, G = shape
    a, sums, (segs, rows, taps) = _count(kind, res, shape, roughness, ct, tab, bounds, wsum, full_search, device)
    if segs >= (1 << 31) or rows >= (1 << 31):
        raise PrefilterTooLarge(f"prefilter plan res={res} roughness={roughness}: {rows} weight rows")
    need = rows * G * 32 * 4 + segs * (8 + 64)
    used = _plan_bytes.get(str(device), 0)
    if used + need > budget_bytes():
        raise PrefilterTooLarge(f"prefilter plan res={res} roughness={roughness} needs {need / 2**30:.1f} GiB "
                                f"({used / 2**30:.1f} GiB of plans already resident, budget {budget_bytes() / 2**30:.0f} GiB; "
                                "raise MRGS_PREFILTER_BUDGET_GB)")
    zero = torch.zeros((1,), dtype=torch.int64, device=device)
    seg_begin = torch.cat([zero, sums[0]]).to(torch.int32)
    slot_begin = torch.cat([zero, sums[1]]).to(torch.int32)
    seg_desc = torch.empty((max(segs, 1), 2), dtype=torch.int32, device=device)
    spans = torch.empty((max(segs, 1), 32), dtype=torch.int16, device=device)
    weights = torch.empty((max(rows, 1) * G * 32,), dtype=torch.float32, device=device)
    plan = Plan(kind, res, G, PW, seg_begin, slot_begin, seg_desc, spans, weights, taps, rows, wsum)
    a.plan = plan.struct()
    with torch.cuda.device(device):
        _lib.check(lib.mrgs_prefilter_plan_fill(C.byref(a), _stream(device)), "mrgs_prefilter_plan_fill")
    _plan_bytes[str(device)] = used + plan.nbytes
    return plan


def estimate_bytes(res: int, ct: float | None) -> int:
    texels = 6 * res * res
    frac = 1.0 if ct is None else (1.0 - ct) / 2.0
    return int(2 * 1.15 * 4 * frac * texels * texels)


def plan_pair(kind: str, res: int, roughness: float, ct: float | None, device, shape=None):
    """(forward plan, transposed plan) of `kind` in {"specular", "diffuse"} for one level; cached per device.
    `ct` is the cos-theta cutoff (ndf_cutoff_costheta), None for the diffuse map; shape = (patch width, rows per
    lane) or None for the automatic choice."""
    device = torch.device(device)
    base_key = (kind, res, None if kind == "diffuse" else (float(roughness), float(ct)), str(device))
    forced_env = os.environ.get("MRGS_PREFILTER_SHAPE") or os.environ.get("MRGS_PREFILTER_ROWS")
    if shape is None and not forced_env and base_key in _shape_cache:
        shape = _shape_cache[base_key]
    if shape is not None and (*base_key, tuple(shape)) in _plan_cache:
        return _plan_cache[(*base_key, tuple(shape))]
    if _plan_bytes.get(str(device), 0) + estimate_bytes(res, ct if kind == "specular" else None) > budget_bytes():
        raise PrefilterTooLarge(f"{kind} prefilter plan for res={res}, roughness={roughness} would exceed "
                                f"MRGS_PREFILTER_BUDGET_GB={budget_bytes() / 2**30:.0f}")
    tab = texel_table(res, device)
    if kind == "specular":
        bounds = specular_bounds(res, ct, device)
        if shape is None:
            shape, forced = choose_shape(res, roughness, ct, tab, bounds, device)
            if not forced:
                _shape_cache[base_key] = shape
        wsum = torch.zeros((6 * res * res,), dtype=torch.float32, device=device)
        fwd = _build(_lib.PREFILTER_SPECULAR, res, shape, roughness, ct, tab, bounds, wsum, False, device)
        bwd = _build(_lib.PREFILTER_SPECULAR_T, res, shape, roughness, ct, tab, bounds, wsum, res % 32 != 0, device)
        if bwd.taps != fwd.taps:
            # the reference's 16x16 tile culling dropped taps asymmetrically: search the whole cube per texel
            _plan_bytes[str(device)] -= bwd.nbytes
            bwd = _build(_lib.PREFILTER_SPECULAR_T, res, shape, roughness, ct, tab, bounds, wsum, True, device)
        if bwd.taps != fwd.taps:
            raise RuntimeError(f"prefilter plan res={res} roughness={roughness}: transposed operator has {bwd.taps} taps, "
                               f"forward has {fwd.taps}")
    elif kind == "diffuse":
        shape = (32, 1) if shape is None else shape     # every texel taps every texel: all shapes cost the same
        fwd = _build(_lib.PREFILTER_DIFFUSE, res, shape, 0.0, 0.0, tab, None, None, True, device)
        bwd = _build(_lib.PREFILTER_DIFFUSE_T, res, shape, 0.0, 0.0, tab, None, None, True, device)
    else:
        raise ValueError(kind)
    _plan_cache[(*base_key, tuple(shape))] = (fwd, bwd)
    return fwd, bwd


def apply_jobs(jobs, backward: bool, device, shard=None, max_ctas: int = 0) -> None:
    """jobs: list of (plan, src, src_stride, dst, dst_stride, nan_where_zero or None); ONE launch per 8 jobs,
    longest tap lists first. shard = (rank, world): only this rank's contiguous share of every plan's patches.
    max_ctas > 0 caps the grid (background execution next to issue-bound kernels, include/mrgs.h)."""
    lib = _lib.load()
    jobs = sorted(jobs, key=lambda j: -(j[0].rows / max(j[0].patches, 1)))
    if shard is not None:
        rank, world = shard
        ranged = []
        for j in jobs:
            n = j[0].patches
            b, e = n * rank // world, n * (rank + 1) // world
            if e > b:
                ranged.append((*j, b, e))
        jobs = ranged
    for i in range(0, len(jobs), _lib.PREFILTER_MAX_JOBS):
        chunk = jobs[i:i + _lib.PREFILTER_MAX_JOBS]
        arr = (_lib.PrefilterJob * len(chunk))()
        for k, (plan, src, ss, dst, ds, nz, *rng) in enumerate(chunk):
            arr[k].plan = plan.struct()
            arr[k].src, arr[k].dst = src.data_ptr(), dst.data_ptr()
            arr[k].nan_where_zero = None if nz is None else nz.data_ptr()
            arr[k].src_stride, arr[k].dst_stride = ss, ds
            if rng:
                arr[k].patch_begin, arr[k].patch_end = rng
        with torch.cuda.device(device):
            _lib.check(lib.mrgs_prefilter_apply(arr, len(chunk), int(backward), int(max_ctas), _stream(device)),
                       "mrgs_prefilter_apply")


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class MipChain:
    """Device state of EnvLight.build_mips for one (max_res, levels, roughnesses, cutoff, device)."""

    def __init__(self, res: int, num_levels: int, roughnesses, cutoff: float, device, shape=None):
        if num_levels > _lib.MAX_MIP_LEVELS or res % (1 << (num_levels - 1)) != 0:
            raise RuntimeError(f"mip chain of {num_levels} levels is not available for res={res}")
        self.res, self.n, self.device = res, num_levels, torch.device(device)
        self.sizes = [res >> l for l in range(num_levels)]
        self.counts = [6 * r * r for r in self.sizes]
        self.offsets = [0]
        for c in self.counts:
            self.offsets.append(self.offsets[-1] + c)
        self.texels = self.offsets[-1]
        self.roughnesses = [float(r) for r in roughnesses]
        total = sum(estimate_bytes(r, cutoff_costheta(ro, cutoff)) for r, ro in zip(self.sizes, self.roughnesses))
        if _plan_bytes.get(str(self.device), 0) + total > budget_bytes():
            raise PrefilterTooLarge(f"prefilter plans of the {res}^2 chain need about {total / 2**30:.1f} GiB")
        self.spec = [plan_pair("specular", r, ro, cutoff_costheta(ro, cutoff), self.device, shape)
                     for r, ro in zip(self.sizes, self.roughnesses)]
        self.diff = plan_pair("diffuse", self.sizes[-1], 0.0, None, self.device)
        self.raw4 = torch.zeros((self.texels + 64, 4), dtype=torch.float32, device=self.device)

    def level_views(self, flat, width):
        return [flat[self.offsets[l]:self.offsets[l + 1]].view(6, r, r, width) for l, r in enumerate(self.sizes)]

    _static = None

    def forward(self, base: torch.Tensor, static: bool = False, shard=None, max_ctas: int = 0):
        """base [6,res,res,3] -> ([prefiltered level l: [6,res>>l,res>>l,3]], diffuse [6,rmin,rmin,3]).
        static: write into the chain's own persistent output buffers (every call returns tensors over the SAME storage),
        which is what consumers captured in a CUDA graph need.
        shard = (rank, world, group): view-sharded steps replicate the cubemap on every rank; instead of every rank
        filtering all of it, each applies its share of every level's patches and ONE allreduce assembles the chain
        (every rank must call)."""
        lib = _lib.load()
        dev = self.device
        base = base.detach().contiguous()
        raw = self.level_views(self.raw4, 4)
        with torch.cuda.device(dev):
            _lib.check(lib.mrgs_mip_pyramid_forward(base.data_ptr(), self.res, self.n, _ptr_array(raw), _stream(dev)),
                       "mrgs_mip_pyramid_forward")
        if static and self._static is not None:
            levels, diffuse, flat = self._static
        else:
            # one flat buffer for all outputs: a sharded build sums it across ranks with ONE collective
            n_d = 6 * self.sizes[-1] * self.sizes[-1]
            flat = torch.empty((self.texels + n_d, 3), dtype=torch.float32, device=dev)
            levels = [flat[self.offsets[l]:self.offsets[l + 1]].view(6, r, r, 3) for l, r in enumerate(self.sizes)]
            diffuse = flat[self.texels:].view(6, self.sizes[-1], self.sizes[-1], 3)
            if static:
                self._static = (levels, diffuse, flat)
        if shard is not None:
            flat.zero_()
        jobs = [(self.spec[l][0], raw[l], 4, levels[l], 3, self.spec[l][0].wsum) for l in range(self.n)]
        jobs.append((self.diff[0], raw[-1], 4, diffuse, 3, None))
        apply_jobs(jobs, False, dev, None if shard is None else shard[:2], max_ctas)
        if shard is not None:
            # every rank filtered its share of every level's patches into the zero-filled buffer: the sum is the chain
            # (x + 0 is exact, so the result equals the unsharded build bit for bit)
            import torch.distributed as dist
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=shard[2])
        return levels, diffuse

    def backward(self, grad4: torch.Tensor | None, grad_diffuse: torch.Tensor | None, shard=None, max_ctas: int = 0) -> torch.Tensor:
        """grad4: [texels,4] gradient of the prefiltered levels (rgb + pad, level after level: the layout of
        EnvLight.level_grad_sink), grad_diffuse [6,rmin,rmin,3]; returns the gradient of the base cubemap."""
        lib = _lib.load()
        dev = self.device
        n_d = 6 * self.sizes[-1] * self.sizes[-1]
        gflat = torch.empty((self.texels + n_d, 3), dtype=torch.float32, device=dev)
        grads = [gflat[self.offsets[l]:self.offsets[l + 1]].view(6, r, r, 3) for l, r in enumerate(self.sizes)]
        g0 = grads[0]
        if shard is not None:
            gflat.zero_()
        jobs = []
        if grad4 is not None:
            if tuple(grad4.shape) != (self.texels, 4) or not grad4.is_contiguous():
                raise RuntimeError(f"expected a contiguous [{self.texels}, 4] level-gradient buffer")
            for l in range(self.n):
                jobs.append((self.spec[l][1], grad4[self.offsets[l]:self.offsets[l + 1]], 4, grads[l], 3, None))
        else:
            for g in grads:
                g.zero_()
        extra = None
        if grad_diffuse is not None:
            extra = gflat[self.texels:].view(6, self.sizes[-1], self.sizes[-1], 3)
            jobs.append((self.diff[1], grad_diffuse.contiguous(), 3, extra, 3, None))
        if jobs:
            apply_jobs(jobs, True, dev, None if shard is None else shard[:2], max_ctas)
            if shard is not None:      # the ranks' shares of d(raw levels) (+ the diffuse map's) in ONE collective
                import torch.distributed as dist
                dist.all_reduce(gflat, op=dist.ReduceOp.SUM, group=shard[2])
        if self.n == 1 and extra is not None:
            g0.add_(extra)
        with torch.cuda.device(dev):
            _lib.check(lib.mrgs_mip_chain_backward(self.res, self.n, _ptr_array(grads),
                                                   None if extra is None else extra.data_ptr(), _stream(dev)),
                       "mrgs_mip_chain_backward")
        return g0


_chain_cache: dict = {}


def get_chain(res, num_levels, roughnesses, cutoff, device, shape=None) -> MipChain:
    forced = os.environ.get("MRGS_PREFILTER_SHAPE"), os.environ.get("MRGS_PREFILTER_ROWS")
    key = (res, num_levels, tuple(float(r) for r in roughnesses), float(cutoff), str(torch.device(device)), shape, forced)
    if key not in _chain_cache:
        _chain_cache[key] = MipChain(res, num_levels, roughnesses, cutoff, device, shape)
    return _chain_cache[key]


class _BuildMips(torch.autograd.Function):
    """base -> (level 0, ..., level n-1, diffuse): EnvLight.build_mips (scene/light.py:72-86) as one autograd node."""

    @staticmethod
    def forward(ctx, base, chain: MipChain, static: bool = False, shard=None, max_ctas: int = 0):
        levels, diffuse = chain.forward(base, static, shard, max_ctas)
        ctx.chain, ctx.shard, ctx.max_ctas = chain, shard, max_ctas
        ctx.set_materialize_grads(False)
        if static:      # fresh tensor objects over the persistent storage (an autograd output must not be reused)
            levels, diffuse = [t.view(t.shape) for t in levels], diffuse.view(diffuse.shape)
        return (*levels, diffuse)

    @staticmethod
    def backward(ctx, *grads):
        chain: MipChain = ctx.chain
        g_levels, g_diffuse = grads[:-1], grads[-1]
        if g_diffuse is None and all(g is None for g in g_levels):
            return None, None, None, None, None  # e.g. the texel gradients went to EnvLight's sink instead of through autograd
        grad4 = None
        if any(g is not None for g in g_levels):
            grad4 = torch.zeros((chain.texels, 4), dtype=torch.float32, device=chain.device)
            for l, g in enumerate(g_levels):
                if g is not None:
                    grad4[chain.offsets[l]:chain.offsets[l + 1], :3] = g.reshape(-1, 3)
        return chain.backward(grad4, g_diffuse, ctx.shard, ctx.max_ctas), None, None, None, None


def build_mips(base, chain: MipChain, static: bool = False, shard=None, max_ctas: int = 0):
    outs = _BuildMips.apply(base, chain, static, shard, max_ctas)
    return list(outs[:-1]), outs[-1]

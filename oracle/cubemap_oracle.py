"""numpy restatement of the reference's cubemap prefilter ops. TEST INFRASTRUCTURE ONLY.

Follows scene/renderutils/c_src/cubemap.cu: pixel_area :17-30, cube_to_dir :32-46, ndfGGX :176-181,
SpecularBoundsKernel :183-246 (including its non-conservative 16x16 interval culling), SpecularCubemapFwd/Bwd
:248-354, DiffuseCubemapFwd/Bwd :110-171, and the host-side cutoff search of
scene/renderutils/ops.py:428-441. Pinned by tests/golden/cubemap_*.npz, produced on a B200 by the
unmodified reference plugin (oracle/build_ref_renderutils.py + tests/golden/make_golden.py)."""
from __future__ import annotations

import numpy as np


def pixel_area(x, y, N):
    if N <= 1:
        return np.ones_like(x, dtype=np.float32)
    H = N // 2
    x = np.abs(x - H).astype(np.float32)
    y = np.abs(y - H).astype(np.float32)
    Hf = np.float32(H)
    dx = np.arctan((x + 1) / Hf) - np.arctan(x / Hf)
    dy = np.arctan((y + 1) / Hf) - np.arctan(y / Hf)
    return (dx * dy).astype(np.float32)


def texel_dirs(N):
    """[6,N,N,3] unit directions of all texel centres (cube_to_dir)."""
    c = (2.0 * ((np.arange(N, dtype=np.float32) + 0.5) / np.float32(N)) - 1.0).astype(np.float32)
    fy, fx = np.meshgrid(c, c, indexing="ij")
    one = np.ones_like(fx)
    faces = [(one, -fy, -fx), (-one, -fy, fx), (fx, one, fy), (fx, -one, -fy), (fx, -fy, one), (-fx, -fy, -one)]
    d = np.stack([np.stack(f, -1) for f in faces], 0).astype(np.float32)
    return d / np.linalg.norm(d, axis=-1, keepdims=True)


def ndf_ggx(alphaSqr, cos):
    # The GPU binary contracts (c*a2 - c)*c + 1 into two fused multiply-adds; at roughness 0.08
    # (a2 = 4e-5) the expression cancels catastrophically, so the rounding placement is visible at
    # the 1e-3 level and is emulated here (float64 product/sum, rounded once to float32).
    c = np.clip(cos, 0.0, 1.0).astype(np.float32)
    c64 = c.astype(np.float64)
    t = (c64 * np.float64(np.float32(alphaSqr)) - c64).astype(np.float32)
    d = (t.astype(np.float64) * c64 + 1.0).astype(np.float32)
    return (np.float64(alphaSqr) / ((d * d).astype(np.float64) * np.pi)).astype(np.float32)


def ndf_cutoff_costheta(roughness, cutoff):  # ops.py:428-441
    nSamples = 1000000
    costheta = np.cos(np.linspace(0, np.pi / 2.0, nSamples))
    c = np.clip(costheta, 0.0, 1.0)
    d = (c * roughness ** 4 - c) * c + 1.0
    D = np.cumsum(roughness ** 4 / (d * d * np.pi))
    idx = np.argmax(D >= D[..., -1] * cutoff)
    return float(costheta[idx])


def _dir_grid(N):
    """[6,N+1,N+1,3] unit directions of cube_to_dir for texel indices 0..N (index N lies one texel
    outside the face; SpecularBoundsKernel evaluates it for the far corners of its culling tiles)."""
    c = (2.0 * ((np.arange(N + 1, dtype=np.float32) + 0.5) / np.float32(N)) - 1.0).astype(np.float32)
    fy, fx = np.meshgrid(c, c, indexing="ij")
    one = np.ones_like(fx)
    faces = [(one, -fy, -fx), (-one, -fy, fx), (fx, one, fy), (fx, -one, -fy), (fx, -fy, one), (-fx, -fy, -one)]
    d = np.stack([np.stack(f, -1) for f in faces], 0).astype(np.float32)
    return d / np.linalg.norm(d, axis=-1, keepdims=True)


def specular_bounds(N, cutoff_cos, TS=16):
    """int32 [6,N,N,6,4] (xmin,xmax,ymin,ymax); empty = (N-1,0,N-1,0). Restates SpecularBoundsKernel
    INCLUDING its 16x16-tile interval test (c_src/cubemap.cu:203-220), which is not conservative
    (tile corners are texel-centre directions, the far ones one texel outside the tile), so it is
    part of the observable result."""
    cutoff = np.float32(cutoff_cos)
    d = texel_dirs(N).reshape(-1, 3)                      # output normals V
    ext = _dir_grid(N)
    inside = ((d @ d.T).astype(np.float32) >= cutoff).reshape(6 * N * N, 6, N, N)
    nt = (N + TS - 1) // TS
    for s in range(6):
        for tx in range(nt):
            for ty in range(nt):
                tsx, tsy = tx * TS, ty * TS
                tex, tey = min((tx + 1) * TS, N), min((ty + 1) * TS, N)
                L = np.stack([ext[s, tsy, tsx], ext[s, tsy, tex], ext[s, tey, tsx], ext[s, tey, tex]], 0)
                lo, hi = L.min(0), L.max(0)                                     # [3]
                maxdp = (np.maximum(lo[None] * d, hi[None] * d)).astype(np.float32)
                maxdp = (maxdp[:, 0] + maxdp[:, 1]) + maxdp[:, 2]
                culled = ~(maxdp >= cutoff)
                inside[culled, s, tsy:tey, tsx:tex] = False
    out = np.zeros((6 * N * N, 6, 4), np.int32)
    for o in range(6 * N * N):
        for s in range(6):
            ys, xs = np.nonzero(inside[o, s])
            out[o, s] = (xs.min(), xs.max(), ys.min(), ys.max()) if xs.size else (N - 1, 0, N - 1, 0)
    return out.reshape(6, N, N, 6, 4)


def specular_weights(N, roughness, cutoff_cos):
    """Dense [6N^2, 6N^2] float32 weight matrix W[out, in] of the specular prefilter (small N only):
    texels inside the per-face bounds AND inside the cone (SpecularCubemapFwdKernel :267-289)."""
    d = texel_dirs(N).reshape(-1, 3)
    ys, xs = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    area = np.tile(pixel_area(xs, ys, N).reshape(-1), 6)
    dots = (d @ d.T).astype(np.float32)                       # L.V
    Hh = d[None, :, :] + d[:, None, :]
    Hn = np.linalg.norm(Hh, axis=-1, keepdims=True)
    Hh = np.where(Hn > 0, Hh / np.maximum(Hn, 1e-30), 0).astype(np.float32)
    VdotH = np.maximum(np.einsum("oik,ok->oi", Hh, d), 0.0).astype(np.float32)
    alphaSqr = np.float32(np.float32(roughness * roughness) ** 2)
    w = np.maximum(dots, 0) * ndf_ggx(alphaSqr, VdotH) * area[None, :] / np.float32(4.0)
    b = specular_bounds(N, cutoff_cos).reshape(6 * N * N, 6, 4)
    X = xs[None, None]
    Y = ys[None, None]
    in_box = ((X >= b[:, :, 0, None, None]) & (X <= b[:, :, 1, None, None]) &
              (Y >= b[:, :, 2, None, None]) & (Y <= b[:, :, 3, None, None])).reshape(6 * N * N, -1)
    return np.where((dots >= np.float32(cutoff_cos)) & in_box, w, 0).astype(np.float32), dots


def specular_cubemap(cubemap, roughness, cutoff=0.99):
    """Returns (rgb / wsum [6,N,N,3], out4 [6,N,N,4], cutoff_cos)."""
    N = cubemap.shape[1]
    ct = ndf_cutoff_costheta(roughness, cutoff)
    W, _ = specular_weights(N, roughness, ct)
    rgb = (W.astype(np.float64) @ cubemap.reshape(-1, 3).astype(np.float64)).astype(np.float32)
    wsum = W.astype(np.float64).sum(1).astype(np.float32)
    out4 = np.concatenate([rgb, wsum[:, None]], 1).reshape(6, N, N, 4)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (rgb / wsum[:, None]).reshape(6, N, N, 3), out4, ct


def specular_cubemap_backward(N, roughness, cutoff_cos, dout_rgb):
    W, _ = specular_weights(N, roughness, cutoff_cos)
    return (W.T.astype(np.float64) @ dout_rgb.reshape(-1, 3).astype(np.float64)).astype(np.float32).reshape(6, N, N, 3)


def diffuse_weights(N):
    d = texel_dirs(N).reshape(-1, 3)
    ys, xs = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    area = np.tile(pixel_area(xs, ys, N).reshape(-1), 6)
    cos = np.clip((d @ d.T).astype(np.float32), 0.0, np.float32(0.999))
    return (cos * area[None, :] / np.float32(3.141592)).astype(np.float32)


def diffuse_cubemap(cubemap):
    N = cubemap.shape[1]
    W = diffuse_weights(N)
    return (W.astype(np.float64) @ cubemap.reshape(-1, 3).astype(np.float64)).astype(np.float32).reshape(6, N, N, 3)


def diffuse_cubemap_backward(N, dout):
    W = diffuse_weights(N)
    return (W.T.astype(np.float64) @ dout.reshape(-1, 3).astype(np.float64)).astype(np.float32).reshape(6, N, N, 3)

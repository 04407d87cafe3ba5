"""Deterministic synthetic surfel clouds and cameras.

Mirrors the reference's random initialisation and camera conventions without importing it:
  positions U[-1.3,1.3]^3                      scene/dataset_readers.py:321
  scales from the 3-nearest-neighbour law       scene/gaussian_model.py:367-368 (simple_knn)
  rotations normalize(U[0,1)^4)                 scene/gaussian_model.py:369, :242
  opacity 0.1 ("init") or sigmoid(N(0,2^2))     scene/gaussian_model.py:371
  camera matrices                               scene/cameras.py:70-84, utils/graphics_utils.py:38-71
Everything is generated with numpy on the host so CPU oracle and GPU path see identical bits.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

SH_C0 = 0.28209479177387814
SEED = 3407  # train_refnerf.py:1777


@dataclass
class Camera:
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    R: np.ndarray               # 3x3 camera-to-world rotation (3DGS convention)
    T: np.ndarray               # world-to-camera translation
    world_view_transform: torch.Tensor   # [4,4], transposed (row-vector convention)
    full_proj_transform: torch.Tensor    # [4,4]
    camera_center: torch.Tensor          # [3]
    K: np.ndarray               # 3x3 intrinsics (dataset_readers.py:283-290)

    @property
    def tanfovx(self):
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self):
        return math.tan(self.FoVy * 0.5)

    @property
    def HWK(self):
        return (self.image_height, self.image_width, self.K)

    def to(self, device):
        return Camera(self.image_width, self.image_height, self.FoVx, self.FoVy, self.R, self.T,
                      self.world_view_transform.to(device), self.full_proj_transform.to(device),
                      self.camera_center.to(device), self.K)


def _world2view(R, t):
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    return np.float32(Rt)


def _projection(znear, zfar, fovX, fovY):
    tanHalfFovY = math.tan(fovY / 2)
    tanHalfFovX = math.tan(fovX / 2)
    top, right = tanHalfFovY * znear, tanHalfFovX * znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def orbit_camera(i: int = 0, n: int = 8, width: int = 800, height: int = 800, fovx: float = 0.6911,
                 radius: float = 4.0, elevation_deg: float = 30.0, znear: float = 0.01,
                 zfar: float = 100.0) -> Camera:
    """View i of n on a circle around the origin (Shiny-Blender-like), looking at the origin."""
    az = 2.0 * math.pi * i / n
    el = math.radians(elevation_deg)
    c = radius * np.array([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)])
    f = -c / np.linalg.norm(c)                       # +z of the camera looks at the origin
    r = np.cross(f, np.array([0.0, 0.0, 1.0]))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)                               # +y of the camera points down
    R = np.stack([r, d, f], axis=1)                  # camera-to-world rotation
    T = -R.T @ c
    fovy = 2.0 * math.atan(math.tan(fovx / 2) * height / width)
    w2v = torch.tensor(_world2view(R, T)).transpose(0, 1).contiguous()
    proj = _projection(znear, zfar, fovx, fovy).transpose(0, 1)
    full = (w2v.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    center = w2v.inverse()[3, :3].contiguous()
    fo = width / (2 * math.tan(fovx / 2))
    fo_y = height / (2 * math.tan(fovy / 2))
    K = np.array([[fo, 0, width / 2], [0, fo_y, height / 2], [0, 0, 1]], dtype=np.float32)
    return Camera(width, height, fovx, fovy, R.astype(np.float32), T.astype(np.float32), w2v, full,
                  center, K)


@dataclass
class SurfelCloud:
    means3D: torch.Tensor    # [P,3]
    scales: torch.Tensor     # [P,2]
    rotations: torch.Tensor  # [P,4] normalised
    opacities: torch.Tensor  # [P,1]
    shs: torch.Tensor        # [P,16,3]
    features: torch.Tensor   # [P,S]

    def to(self, device):
        return SurfelCloud(*(t.to(device) for t in (self.means3D, self.scales, self.rotations,
                                                     self.opacities, self.shs, self.features)))

    @property
    def P(self):
        return self.means3D.shape[0]


def make_cloud(P: int, S: int = 8, opacity: str = "trained", seed: int = SEED, extent: float = 1.3,
               scale_mult: float = 1.0, color: str = "bright", unbounded: bool = False) -> SurfelCloud:
    rng = np.random.RandomState(seed)
    if unbounded:  # Ref-Real-like: contracted cloud + 30 % background shell (config C4)
        n_bg = int(0.3 * P)
        core = rng.normal(0.0, 2.0, size=(P - n_bg, 3))
        dirs = rng.normal(size=(n_bg, 3))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        shell = dirs * rng.uniform(5.0, 20.0, size=(n_bg, 1))
        means = np.concatenate([core, shell], 0).astype(np.float32)
        vol = (4.0 / 3.0) * math.pi * 6.0 ** 3
    else:
        means = rng.uniform(-extent, extent, size=(P, 3)).astype(np.float32)
        vol = (2 * extent) ** 3
    # sqrt(mean of the 3 nearest squared distances) for a Poisson cloud: 0.7525 * (V/P)^(1/3)
    s0 = 0.7525 * (vol / max(P, 1)) ** (1.0 / 3.0) * scale_mult
    jitter = np.exp(rng.normal(0.0, 0.3, size=(P, 1)))
    scales = np.maximum(s0 * jitter * np.ones((1, 2)), math.sqrt(1e-7)).astype(np.float32)
    rot = rng.uniform(0.0, 1.0, size=(P, 4))
    rot = (rot / np.linalg.norm(rot, axis=1, keepdims=True)).astype(np.float32)
    if opacity == "init":
        opa = np.full((P, 1), 0.1, dtype=np.float32)
    elif opacity == "trained":
        opa = (1.0 / (1.0 + np.exp(-rng.normal(0.0, 2.0, size=(P, 1))))).astype(np.float32)
    else:
        raise ValueError(opacity)
    base = rng.uniform(0.0, 1.0, size=(P, 3))
    if color == "init":  # scene/dataset_readers.py:321 random init is nearly black
        base = base / 255.0
    shs = np.zeros((P, 16, 3), dtype=np.float32)
    shs[:, 0, :] = (base - 0.5) / SH_C0
    shs[:, 1:, :] = rng.normal(0.0, 0.05, size=(P, 15, 3))
    feats = np.zeros((P, S), dtype=np.float32)
    if S > 0:
        raw = rng.normal(0.0, 1.0, size=(P, S))
        n_sig = min(S, 5)
        feats[:, :n_sig] = 1.0 / (1.0 + np.exp(-raw[:, :n_sig]))     # refl, rough, albedo
        if S > 5:
            feats[:, 5:] = np.maximum(0.3 * raw[:, 5:], 0.0)          # indirect
    t = torch.from_numpy
    return SurfelCloud(t(means), t(scales), t(rot), t(opa), t(shs), t(feats))


def upstream_grads(S: int, H: int, W: int, seed: int = SEED + 1):
    """dL/dcolor [3,H,W], dL/dfeature [S,H,W], dL/dallmap [7,H,W] ~ N(0,1)/N (all branches live)."""
    rng = np.random.RandomState(seed)
    N = H * W
    mk = lambda c: torch.from_numpy((rng.normal(size=(c, H, W)) / N).astype(np.float32))
    return mk(3), mk(S), mk(7)

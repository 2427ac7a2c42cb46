"""Seeded synthetic inputs for the Hair-GS render path and the callers either side of
the rasterizer that Hair-GS keeps in torch:

  * strand scenes shaped like USC-HairSalon / Cem Yuksel hair (roots on a scalp hemisphere, random-walk
    growth with gravity) and Stage-I free-Gaussian "blob" scenes;
  * the camera rig of utils/camera.py:41-100 with the matrix conventions of scene/cameras.py:87-108
    (world_view_transform = W2C^T, full_proj_transform = W2C^T P^T, camera_center from the inverse,
    znear 0.01, zfar 100) and utils/graphics.py:51-71 (getProjectionMatrix);
  * the strand-aligned Gaussian parameterisation of scene/hair_gaussian_model.py:134-206 with
    utils/transform.py:54-86, including a restatement of pytorch3d.transforms.matrix_to_quaternion
    (pytorch3d is an un-vendored, un-pinned dependency of the reference: environment.yml:11; the
    algorithm restated here is the published one — four candidates from sqrt(max(0, 1 +- m00 +- m11 +- m22)),
    best-conditioned pick, divide by 2*max(q_abs, 0.1); pinned by R(q) round-trip tests, "parity unpinned"
    against the reference itself).

Everything here is deterministic given the seed and runs on CPU or CUDA tensors alike.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch

SH_C0 = 0.28209479177387814
DIST_TO_SCALE = 0.5102133812190369  # scene/gaussian_model.py:35
MIN_VAL = 1e-7


# ---------------------------------------------------------------------------------------------------
# cameras
# ---------------------------------------------------------------------------------------------------
@dataclass
class Camera:
    """What gaussian_renderer.render() reads from a viewpoint camera (gaussian_renderer/__init__.py:53-66)."""
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: torch.Tensor  # [4,4] = W2C^T
    full_proj_transform: torch.Tensor   # [4,4] = W2C^T @ P^T
    camera_center: torch.Tensor         # [3]

    @property
    def tanfovx(self):
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self):
        return math.tan(self.FoVy * 0.5)

    def to(self, device):
        return Camera(self.image_width, self.image_height, self.FoVx, self.FoVy,
                      self.world_view_transform.to(device), self.full_proj_transform.to(device),
                      self.camera_center.to(device))


def projection_matrix(znear, zfar, fovx, fovy):
    """utils/graphics.py:51-71."""
    tan_y = math.tan(fovy / 2)
    tan_x = math.tan(fovx / 2)
    top, right = tan_y * znear, tan_x * znear
    bottom, left = -top, -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def camera_from_w2c(w2c, width, height, fovx, fovy, device="cpu"):
    """scene/cameras.py:87-108 with trans = 0, scale = 1."""
    wv = torch.tensor(np.float32(w2c)).transpose(0, 1)
    proj = projection_matrix(0.01, 100.0, fovx, fovy).transpose(0, 1)
    full = wv.unsqueeze(0).bmm(proj.unsqueeze(0)).squeeze(0)
    center = wv.inverse()[3, :3]
    return Camera(int(width), int(height), float(fovx), float(fovy), wv.contiguous().to(device),
                  full.contiguous().to(device), center.contiguous().to(device))


def _look_at_w2c(eye, target, up):
    f = target - eye
    f = f / np.linalg.norm(f)
    x = np.cross(f, up)
    x = x / np.linalg.norm(x)
    y = np.cross(f, x)  # image "down" (COLMAP / OpenCV camera: +z forward, +y down)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = x, y, f, eye
    return np.linalg.inv(c2w)


def orbit_cameras(n_views, width, height, center=(0.0, -0.05, 0.0), radius=0.5, device="cpu"):
    """n_views-1 cameras on a circle around the y axis + one top view (utils/camera.py:41-100);
    pinhole with f = width/2 px, i.e. FoV 90 degrees (the reference's 500 px at 1000^2)."""
    center = np.asarray(center, dtype=np.float64)
    focal = width / 2.0
    fovx = 2 * math.atan(width / (2 * focal))
    fovy = 2 * math.atan(height / (2 * focal))
    cams = []
    ring = max(n_views - 1, 1)
    for i in range(n_views - 1 if n_views > 1 else 1):
        a = 2 * math.pi * i / ring
        eye = center + radius * np.array([math.sin(a), 0.0, math.cos(a)])
        cams.append(camera_from_w2c(_look_at_w2c(eye, center, np.array([0.0, 1.0, 0.0])), width, height, fovx, fovy,
                                    device))
    if n_views > 1:
        eye = center + np.array([0.0, radius, 0.0])
        cams.append(camera_from_w2c(_look_at_w2c(eye, center, np.array([0.0, 0.0, -1.0])), width, height, fovx, fovy,
                                    device))
    return cams


# ---------------------------------------------------------------------------------------------------
# scenes
# ---------------------------------------------------------------------------------------------------
@dataclass
class StrandScene:
    endpoints: torch.Tensor       # [E,3] joints (shared between consecutive segments)
    endpoint_pairs: torch.Tensor  # [P,2] int64
    width: torch.Tensor           # [P,1] log sigma_yz
    opacity_logit: torch.Tensor   # [P,1]
    mask_logit: torch.Tensor      # [P,1]
    features_dc: torch.Tensor     # [P,1,3]
    features_rest: torch.Tensor   # [P,M-1,3]
    n_strands: int

    def to(self, device):
        return StrandScene(*[t.to(device) if torch.is_tensor(t) else t for t in
                             (self.endpoints, self.endpoint_pairs, self.width, self.opacity_logit, self.mask_logit,
                              self.features_dc, self.features_rest)], self.n_strands)


def strand_scene(n_strands, n_vertices, seed=0, sh_coeffs=1):
    """Roots uniform on the upper hemisphere of a 0.10 m sphere; v_{i+1} = v_i + l*normalize(d_i),
    d_{i+1} = normalize(d_i + 0.15 N(0,I) + 0.02 g), g = (0,-1,0), strand length L ~ U(0.10, 0.30) m."""
    rng = np.random.default_rng(seed)
    S, V = n_strands, n_vertices
    z = rng.uniform(0.0, 1.0, S)
    phi = rng.uniform(0.0, 2 * np.pi, S)
    rxy = np.sqrt(1 - z * z)
    normal = np.stack([rxy * np.cos(phi), z, rxy * np.sin(phi)], 1)  # y-up
    L = rng.uniform(0.10, 0.30, S)
    step = (L / (V - 1))[:, None]
    verts = np.empty((S, V, 3))
    verts[:, 0] = 0.10 * normal
    d = normal.copy()
    g = np.array([0.0, -1.0, 0.0])
    for i in range(V - 1):
        verts[:, i + 1] = verts[:, i] + step * d
        d = d + 0.15 * rng.standard_normal((S, 3)) + 0.02 * g
        d /= np.linalg.norm(d, axis=1, keepdims=True)
    endpoints = verts.reshape(S * V, 3)
    base = (np.arange(S) * V)[:, None] + np.arange(V - 1)[None, :]
    pairs = np.stack([base, base + 1], -1).reshape(-1, 2)
    P = pairs.shape[0]
    width = np.log(rng.uniform(1e-4, 3e-4, (P, 1)))
    opac = rng.uniform(0.3, 0.95, (P, 1))
    mask = rng.uniform(0.6, 0.99, (P, 1))
    hue = rng.uniform(0.0, 1.0, S)
    rgb = 0.5 + 0.4 * np.stack([np.cos(2 * np.pi * (hue + k / 3.0)) for k in range(3)], 1)  # per-strand colour
    rgb_seg = np.repeat(rgb, V - 1, axis=0)
    f_dc = ((rgb_seg - 0.5) / SH_C0)[:, None, :]
    f_rest = 0.05 * rng.standard_normal((P, sh_coeffs - 1, 3))
    t = lambda a, dt=torch.float32: torch.tensor(np.ascontiguousarray(a), dtype=dt)  # noqa: E731
    return StrandScene(t(endpoints), t(pairs, torch.int64), t(width), t(np.log(opac / (1 - opac))),
                       t(np.log(mask / (1 - mask))), t(f_dc), t(f_rest), S)


@dataclass
class BlobScene:
    means3D: torch.Tensor    # [P,3]
    scales: torch.Tensor     # [P,3] (already activated)
    rotations: torch.Tensor  # [P,4] unit quaternions (w,x,y,z)
    opacities: torch.Tensor  # [P,1] in (0,1)
    shs: torch.Tensor        # [P,M,3]

    def to(self, device):
        return BlobScene(*[t.to(device) for t in (self.means3D, self.scales, self.rotations, self.opacities, self.shs)])


def blob_scene(P, seed=0, sh_coeffs=16):
    """Stage-I style free Gaussians: points in a 0.25 m ball biased to the hair shell,
    scale = exp(N(log 2e-3, 0.5^2)) per axis, random unit quaternions, M SH coefficients."""
    rng = np.random.default_rng(seed)
    dirs = rng.standard_normal((P, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    r = 0.25 * np.clip(0.55 + 0.25 * rng.standard_normal(P), 0.02, 1.0)
    means = dirs * r[:, None] + np.array([0.0, -0.05, 0.0])
    scales = np.exp(rng.normal(math.log(2e-3), 0.5, (P, 3)))
    q = rng.standard_normal((P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opac = rng.uniform(0.05, 0.95, (P, 1))
    sh = 0.05 * rng.standard_normal((P, sh_coeffs, 3))
    sh[:, 0, :] = (rng.uniform(0.1, 0.9, (P, 3)) - 0.5) / SH_C0
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32)  # noqa: E731
    return BlobScene(t(means), t(scales), t(q), t(opac), t(sh))


# ---------------------------------------------------------------------------------------------------
# strand-aligned parameterisation (the torch glue upstream of the rasterizer)
# ---------------------------------------------------------------------------------------------------
def matrix_to_quaternion(matrix):
    """Rotation matrices [...,3,3] -> quaternions (w,x,y,z); restates pytorch3d.transforms.matrix_to_quaternion
    (the call at utils/transform.py:84-85).  Sign is standardised to w >= 0 as in recent pytorch3d."""
    m00, m01, m02 = matrix[..., 0, 0], matrix[..., 0, 1], matrix[..., 0, 2]
    m10, m11, m12 = matrix[..., 1, 0], matrix[..., 1, 1], matrix[..., 1, 2]
    m20, m21, m22 = matrix[..., 2, 0], matrix[..., 2, 1], matrix[..., 2, 2]
    x = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], -1)
    q_abs = torch.zeros_like(x)
    pos = x > 0
    q_abs[pos] = torch.sqrt(x[pos])
    quat_by_rijk = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1),
    ], -2)
    flr = torch.tensor(0.1, dtype=q_abs.dtype, device=q_abs.device)
    cand = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    idx = q_abs.argmax(-1)
    out = torch.gather(cand, -2, idx[..., None, None].expand(*idx.shape, 1, 4)).squeeze(-2)
    return torch.where(out[..., :1] < 0, -out, out)


def rotation_from_x_axis(v2, eps=1e-7):
    """Rotation taking +x onto normalize(v2): R = I + K + K^2/(1 + x.d) (utils/transform.py:69-86).
    The norm is floored at 1e-30 so collapsed segments give R = I instead of NaN (the reference filters them
    out with a boolean mask before calling; callers here select with torch.where, which needs finite values)."""
    v2 = v2 / torch.norm(v2, dim=1, keepdim=True).clamp_min(1e-30)
    v1 = torch.zeros_like(v2)
    v1[:, 0] = 1.0
    dot = torch.clamp(torch.sum(v1 * v2, dim=1), -1 + eps, 1 - eps)
    cross = torch.cross(v1, v2, dim=1)
    K = torch.zeros(v2.shape[0], 3, 3, dtype=v2.dtype, device=v2.device)
    K[:, 0, 1] = -cross[:, 2]
    K[:, 0, 2] = cross[:, 1]
    K[:, 1, 2] = -cross[:, 0]
    K[:, 1, 0] = -K[:, 0, 1]
    K[:, 2, 0] = -K[:, 0, 2]
    K[:, 2, 1] = -K[:, 1, 2]
    eye = torch.eye(3, dtype=v2.dtype, device=v2.device).repeat(K.shape[0], 1, 1)
    return eye + K + torch.bmm(K, K) / (1 + dot)[:, None, None]


def strand_xyz(endpoints, endpoint_pairs):
    """HairGaussianModel.get_xyz (scene/hair_gaussian_model.py:167-172): segment centres."""
    return torch.mean(endpoints[endpoint_pairs], dim=1)


def strand_scaling(endpoints, endpoint_pairs, width):
    """get_scaling (:134-145): sigma_x = max(|e1-e0|/2 * k, 1e-7), sigma_yz = exp(width)."""
    pairs = endpoints[endpoint_pairs]
    dist = torch.norm(pairs[:, 1] - pairs[:, 0], p=2, dim=1, keepdim=True)
    scale_x = torch.clamp(dist / 2 * DIST_TO_SCALE, min=MIN_VAL)
    return torch.cat((scale_x, torch.exp(width.repeat(1, 2))), dim=1)


def strand_rotation(endpoints, endpoint_pairs):
    """get_rotation (:147-165): quaternion (w,x,y,z) rotating +x onto the segment; identity when collapsed;
    NOT re-normalised (neither does the reference, :164, nor the rasterizer, forward.cu:127).
    Sync-free: the reference's boolean-mask indexing is replaced by torch.where."""
    pairs = endpoints[endpoint_pairs]
    v2 = pairs[:, 1] - pairs[:, 0]
    valid = torch.norm(v2, p=2, dim=1) > MIN_VAL
    q = matrix_to_quaternion(rotation_from_x_axis(v2))
    ident = torch.zeros_like(q)
    ident[:, 0] = 1.0
    return torch.where(valid[:, None], q, ident)


def strand_orientation(endpoints, endpoint_pairs):
    """get_orientation (:188-201): unit segment direction, +x when collapsed."""
    pairs = endpoints[endpoint_pairs]
    d = pairs[:, 1] - pairs[:, 0]
    norm = torch.norm(d, p=2, dim=1, keepdim=True)
    xaxis = torch.zeros_like(d)
    xaxis[:, 0] = 1.0
    return torch.where(norm >= MIN_VAL, d / norm.clamp_min(1e-30), xaxis)


def strand_gaussians(endpoints, endpoint_pairs, width):
    """All four derived quantities: (means3D[P,3], scales[P,3], rotations[P,4], orientation[P,3])."""
    return (strand_xyz(endpoints, endpoint_pairs), strand_scaling(endpoints, endpoint_pairs, width),
            strand_rotation(endpoints, endpoint_pairs), strand_orientation(endpoints, endpoint_pairs))

"""Synthetic RGB-D frame pairs with analytic ground truth.

The reference ships no test inputs for the registration seat besides one Kinect pair without
ground truth (reference data/exp1), so benchmarks and parity tests run on rendered scenes:

* camera 640x480, fx=fy=525, cx=319.5, cy=235.5, factor=1000 -- the constants of reference
  src/convert2PCD.cpp:19-23; clouds are back-projected with the formula and the row-major,
  holes-skipped ordering of src/convert2PCD.cpp:54-80.
* scene S1 "room corner": floor y=+1.2 m, back wall z=4 m, side wall x=-2 m (three non-parallel
  planes = ``max_planes: 3`` of reference parameters.yaml:47); camera yawed 25 deg towards the side
  wall and pitched 10 deg down.  Scene S0: one fronto-parallel wall (degenerate for point-to-plane).
* relative pose of pair i: rotation U[0.01,0.05] rad about a uniform axis, translation U[0.01,0.05] m
  in a uniform direction, drawn from the counter-based stream splitmix64(base_seed + i).

All randomness is counter based (splitmix64), so any pair can be generated independently on any
rank.  numpy only; this module is product-side (used by bench.py and the tests), not oracle code.
"""
from __future__ import annotations

from dataclasses import dataclass, replace
import numpy as np

BASE_SEED = 20140501

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def rand_u64(seed: int, a, b=0, c=0) -> np.ndarray:
    """Same stream as orc_rand()/s3d_rand(): mix(mix(mix(seed+a)+b)+c)."""
    with np.errstate(over="ignore"):
        s = np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
        x = _mix(s + np.asarray(a, dtype=np.uint64))
        x = _mix(x + np.asarray(b, dtype=np.uint64))
        x = _mix(x + np.asarray(c, dtype=np.uint64))
        return x


def uniform01(seed: int, a, b=0, c=0) -> np.ndarray:
    return (rand_u64(seed, a, b, c) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def gaussian(seed: int, n: int, stream: int) -> np.ndarray:
    """n standard normals via Box-Muller from two counter-based uniform streams."""
    i = np.arange(n, dtype=np.uint64)
    u1 = uniform01(seed, i, stream, 1)
    u2 = uniform01(seed, i, stream, 2)
    u1 = np.maximum(u1, 1e-300)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


@dataclass(frozen=True)
class Camera:
    fx: float = 525.0
    fy: float = 525.0
    cx: float = 319.5
    cy: float = 235.5
    factor: float = 1000.0
    width: int = 640
    height: int = 480

    def scaled(self, s: float) -> "Camera":
        """Camera of the same field of view at s times the resolution."""
        return replace(self, fx=self.fx * s, fy=self.fy * s, cx=(self.cx + 0.5) * s - 0.5,
                       cy=(self.cy + 0.5) * s - 0.5, width=int(round(self.width * s)),
                       height=int(round(self.height * s)))


def rot_axis_angle(axis, angle: float) -> np.ndarray:
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def make_T(R, t) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def scene_planes(name: str = "S1"):
    """World-frame planes n.X + d = 0 (camera convention: x right, y down, z forward)."""
    if name == "S1":
        N = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])
        d = np.array([1.2, 4.0, 2.0])  # y=1.2 floor, z=4 back wall, x=-2 side wall
    elif name == "S0":
        N = np.array([[0.0, 0.0, -1.0]])
        d = np.array([3.0])
    elif name == "S2":  # two planes only: one sliding direction left (degenerate)
        N = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0]])
        d = np.array([1.2, 4.0])
    else:
        raise ValueError(name)
    return N, d


def base_pose(scene: str = "S1") -> np.ndarray:
    """Camera-to-world pose of frame 1."""
    if scene == "S0":
        return np.eye(4)
    yaw = rot_axis_angle([0, 1, 0], -np.deg2rad(25.0))   # towards -x (side wall)
    pitch = rot_axis_angle([1, 0, 0], -np.deg2rad(10.0))  # look down (y is down)
    return make_T(yaw @ pitch, np.zeros(3))


def random_rel_pose(seed: int, rot_range=(0.01, 0.05), trans_range=(0.01, 0.05)) -> np.ndarray:
    """T that maps frame-1 coordinates into frame-2 coordinates."""
    u = uniform01(seed, np.arange(8, dtype=np.uint64), 7, 0)
    def sphere(a, b):
        z = 2 * a - 1
        ph = 2 * np.pi * b
        r = np.sqrt(max(0.0, 1 - z * z))
        return np.array([r * np.cos(ph), r * np.sin(ph), z])
    axis = sphere(u[0], u[1])
    ang = rot_range[0] + (rot_range[1] - rot_range[0]) * u[2]
    tdir = sphere(u[3], u[4])
    tmag = trans_range[0] + (trans_range[1] - trans_range[0]) * u[5]
    return make_T(rot_axis_angle(axis, ang), tdir * tmag)


def render_depth(C: np.ndarray, cam: Camera, scene: str = "S1", sigma: float = 0.002,
                 seed: int = 0, stream: int = 0):
    """Analytic ray-plane depth (metres, float64, HxW) + plane id per pixel, noise N(0,sigma) on z."""
    N, d = scene_planes(scene)
    v, u = np.mgrid[0:cam.height, 0:cam.width]
    dirs = np.stack([(u - cam.cx) / cam.fx, (v - cam.cy) / cam.fy, np.ones_like(u, dtype=np.float64)], -1)
    R, o = C[:3, :3], C[:3, 3]
    wd = dirs @ R.T
    num = -(N @ o + d)                      # (k,)
    den = wd @ N.T                          # (H,W,k)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = num / den
    s = np.where((s > 1e-6) & np.isfinite(s), s, np.inf)
    pid = np.argmin(s, -1)
    z = np.min(s, -1)
    z = np.where(np.isfinite(z), z, 0.0)
    if sigma > 0:
        z = z + sigma * gaussian(seed, z.size, stream).reshape(z.shape) * (z > 0)
    return z, pid.astype(np.int32)


def quantize_depth(z: np.ndarray, cam: Camera) -> np.ndarray:
    return np.clip(np.rint(z * cam.factor), 0, 65535).astype(np.uint16)


def backproject(z: np.ndarray, cam: Camera):
    """(N,4) float32 cloud of the non-zero pixels, row-major, + flat pixel index of every point.
    x=(n-cx)z/fx, y=(m-cy)z/fy computed in double and stored as float (src/convert2PCD.cpp:64-68)."""
    z = np.asarray(z)
    zz = z.astype(np.float64) / cam.factor if z.dtype == np.uint16 else z.astype(np.float64)
    v, u = np.mgrid[0:cam.height, 0:cam.width]
    m = (z != 0).ravel()
    x = ((u - cam.cx) * zz / cam.fx).ravel()[m]
    y = ((v - cam.cy) * zz / cam.fy).ravel()[m]
    out = np.empty((int(m.sum()), 4), dtype=np.float32)
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = x, y, zz.ravel()[m], 1.0
    return out, np.flatnonzero(m)


def analytic_normals(C: np.ndarray, pid: np.ndarray, scene: str, pix: np.ndarray) -> np.ndarray:
    """(N,4) float32: plane normal of every point in the camera frame of pose C, w = 1 (valid)."""
    N, _ = scene_planes(scene)
    nc = (N @ C[:3, :3]).astype(np.float32)      # n_c = R^T n_w
    out = np.ones((pix.size, 4), dtype=np.float32)
    out[:, :3] = nc[pid.ravel()[pix]]
    return out


def punch_holes(z: np.ndarray, seed: int, fraction: float) -> np.ndarray:
    """Zero out random 16x16 blocks (sensor holes) until ~fraction of the pixels are gone."""
    if fraction <= 0:
        return z
    z = z.copy()
    H, W = z.shape
    nb = max(1, int(fraction * H * W / 256.0))
    r = rand_u64(seed, np.arange(nb, dtype=np.uint64), 99, 0)
    ys = (r % np.uint64(max(1, H - 16))).astype(np.int64)
    xs = ((r >> np.uint64(32)) % np.uint64(max(1, W - 16))).astype(np.int64)
    for y0, x0 in zip(ys, xs):
        z[y0:y0 + 16, x0:x0 + 16] = 0
    return z


def make_pair(i: int = 0, base_seed: int = BASE_SEED, cam: Camera = Camera(), scene: str = "S1",
              sigma: float = 0.002, quantize: bool = False, holes: float = 0.0,
              rot_range=(0.01, 0.05), trans_range=(0.01, 0.05)):
    """Frame pair i.  Returns a dict with
       src, tgt        (N,4)/(M,4) float32 clouds (frame 1 / frame 2)
       T_gt            4x4 float64, X2 = T_gt X1
       tgt_normals     (M,4) float32 analytic plane normals of the target (w = 1)
       src_depth, tgt_depth   uint16 depth images when quantize=True else None
    """
    seed = base_seed + i
    T = random_rel_pose(seed, rot_range, trans_range)
    C1 = base_pose(scene)
    C2 = C1 @ np.linalg.inv(T)
    z1, pid1 = render_depth(C1, cam, scene, sigma, seed, 11)
    z2, pid2 = render_depth(C2, cam, scene, sigma, seed, 12)
    z1 = punch_holes(z1, seed + 1000003, holes)
    z2 = punch_holes(z2, seed + 2000003, holes)
    d1 = d2 = None
    if quantize:
        d1, d2 = quantize_depth(z1, cam), quantize_depth(z2, cam)
        src, pix1 = backproject(d1, cam)
        tgt, pix2 = backproject(d2, cam)
    else:
        src, pix1 = backproject(z1, cam)
        tgt, pix2 = backproject(z2, cam)
    return dict(src=src, tgt=tgt, T_gt=T, tgt_normals=analytic_normals(C2, pid2, scene, pix2),
                src_normals=analytic_normals(C1, pid1, scene, pix1),
                src_depth=d1, tgt_depth=d2, cam=cam, seed=seed)


def make_map(base_seed: int = BASE_SEED, cam: Camera = Camera(), scene: str = "S1", sigma: float = 0.002):
    """Config 5: fused map of 4 frames (+-0.15 rad yaw, +-0.2 m lateral) expressed in frame-0
    coordinates and concatenated without voxel filtering (like reference src/saveOutput.cpp:87-88),
    and a 5th incoming frame.  Returns dict(map, map_normals, frame, T_gt) where T_gt maps the
    incoming frame's coordinates into map (frame-0) coordinates."""
    C0 = base_pose(scene)
    clouds, normals = [], []
    offs = [(-0.15, -0.2), (-0.15, 0.2), (0.15, -0.2), (0.15, 0.2)]
    for k, (yaw, lat) in enumerate(offs):
        D = make_T(rot_axis_angle([0, 1, 0], yaw), [lat, 0, 0])   # frame-k -> frame-0
        Ck = C0 @ D
        z, pid = render_depth(Ck, cam, scene, sigma, base_seed, 20 + k)
        pts, pix = backproject(z, cam)
        nrm = analytic_normals(Ck, pid, scene, pix)
        p0 = pts.copy()
        p0[:, :3] = (pts[:, :3].astype(np.float64) @ D[:3, :3].T + D[:3, 3]).astype(np.float32)
        n0 = nrm.copy()
        n0[:, :3] = (nrm[:, :3].astype(np.float64) @ D[:3, :3].T).astype(np.float32)
        clouds.append(p0)
        normals.append(n0)
    Tin = random_rel_pose(base_seed + 77)           # frame-0 -> incoming frame
    Cin = C0 @ np.linalg.inv(Tin)
    z, _ = render_depth(Cin, cam, scene, sigma, base_seed, 30)
    frame, _ = backproject(z, cam)
    return dict(map=np.concatenate(clouds), map_normals=np.concatenate(normals), frame=frame,
                T_gt=np.linalg.inv(Tin))


def pose_error(T_a: np.ndarray, T_b: np.ndarray):
    """(rotation angle [rad], translation distance [m]) between two poses; the metric of reference
    tools/evaluate_rpe.py:162-170 and src/exp1/exp1_2.cpp:167-170 applied to T_a^-1 T_b."""
    E = np.linalg.inv(T_a) @ T_b
    c = np.clip((np.trace(E[:3, :3]) - 1.0) / 2.0, -1.0, 1.0)
    s = 0.5 * np.linalg.norm([E[2, 1] - E[1, 2], E[0, 2] - E[2, 0], E[1, 0] - E[0, 1]])
    return float(np.arctan2(s, c)), float(np.linalg.norm(T_a[:3, 3] - T_b[:3, 3]))

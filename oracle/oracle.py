"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (slam3d_gx_b200) never imports this module.
PARITY UNPINNED: see oracle/oracle_common.h.
"""
import ctypes as C
import os
import subprocess
import numpy as np

from slam3d_gx_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("icp_oracle.c", "plane_oracle.c", "filter_oracle.c", "oracle_common.h")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       env={**os.environ, "CC": "gcc"})
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_pose_norm.restype = C.c_double
        _LIB.oracle_pose_norm.argtypes = [C.POINTER(C.c_double)]
    return _LIB


def _f4(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def nn(src, tgt, T=None, brute=False, nthreads=0):
    """Exact NN (float32 d^2, lowest-index ties) of T*src in tgt -> (idx int32, d2 float32)."""
    src, tgt = _f4(src), _f4(tgt)
    idx = np.empty(len(src), np.int32)
    d2 = np.empty(len(src), np.float32)
    T12 = None if T is None else np.ascontiguousarray(np.asarray(T, np.float64)[:3, :4], dtype=np.float32)
    tp = _fp(T12) if T12 is not None else None
    ip, dp = idx.ctypes.data_as(C.POINTER(C.c_int)), _fp(d2)
    if brute:
        lib().oracle_nn_brute(_fp(src), len(src), _fp(tgt), len(tgt), tp, ip, dp)
    else:
        lib().oracle_nn_kdtree(_fp(src), len(src), _fp(tgt), len(tgt), tp, ip, dp, int(nthreads))
    return idx, d2


def icp(src, tgt, tgt_normals=None, guess=None, params=None, nthreads=1, want_nn=False):
    src, tgt = _f4(src), _f4(tgt)
    params = params or _abi.icp_params()
    res = _abi.Result()
    nrm = _f4(tgt_normals) if tgt_normals is not None else None
    g = None if guess is None else np.ascontiguousarray(guess, dtype=np.float64)
    nn_out = np.empty(len(src), np.int32) if want_nn else None
    rc = lib().oracle_icp(_fp(src), len(src), _fp(tgt), _fp(nrm) if nrm is not None else None, len(tgt),
                          g.ctypes.data_as(C.POINTER(C.c_double)) if g is not None else None,
                          C.byref(params), C.byref(res),
                          nn_out.ctypes.data_as(C.POINTER(C.c_int)) if want_nn else None, int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle_icp failed rc={rc}")
    out = _abi.result_to_dict(res)
    if want_nn:
        out["nn"] = nn_out
    return out


def segment_planes(xyzw, params=None, want_counts=False):
    xyzw = _f4(xyzw)
    params = params or _abi.plane_params()
    n = len(xyzw)
    planes = (_abi.Plane * max(1, params.max_planes))()
    labels = np.empty(n, np.int32)
    normals = np.zeros((n, 4), np.float32)
    ncand = params.max_iterations + _abi.PLANE_CANDIDATES_EXTRA
    counts = np.zeros((max(1, params.max_planes), ncand), np.int32)
    k = lib().oracle_segment_planes(_fp(xyzw), n, C.byref(params), planes,
                                    labels.ctypes.data_as(C.POINTER(C.c_int32)), _fp(normals),
                                    counts.ctypes.data_as(C.POINTER(C.c_int32)))
    out = dict(planes=[dict(coef=np.array(list(planes[i].coef), np.float32), inliers=planes[i].inliers,
                            hypotheses=planes[i].hypotheses) for i in range(k)],
               labels=labels, normals=normals)
    if want_counts:
        out["cand_counts"] = counts[:k]
    return out


def planar_keypoints(depth, cam, uv, threshold=0.01, min_inliers=40, seed=12345):
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    uv = np.ascontiguousarray(uv, dtype=np.int32).reshape(-1, 2)
    flags = np.zeros(len(uv), np.uint8)
    camc = _abi.camera_c(cam)
    lib().oracle_planar_keypoints(depth.ctypes.data_as(C.POINTER(C.c_uint16)), depth.shape[1], depth.shape[0],
                                  C.byref(camc), uv.ctypes.data_as(C.POINTER(C.c_int32)), len(uv),
                                  C.c_float(threshold), int(min_inliers), C.c_uint64(seed),
                                  flags.ctypes.data_as(C.POINTER(C.c_uint8)))
    return flags


def backproject(depth, cam, z_max=0.0):
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    out = np.empty((depth.size, 4), np.float32)
    camc = _abi.camera_c(cam)
    k = lib().oracle_backproject(depth.ctypes.data_as(C.POINTER(C.c_uint16)), depth.shape[1], depth.shape[0],
                                 C.byref(camc), C.c_float(z_max), _fp(out))
    return out[:k].copy()


def passthrough_z(pts, z_min, z_max):
    """pcl::PassThrough on z (reference src/GraphicEnd.cpp:283-285)."""
    pts = _f4(pts)
    out = np.empty((max(len(pts), 1), 4), np.float32)
    k = lib().oracle_passthrough_z(_fp(pts), len(pts), C.c_float(z_min), C.c_float(z_max), _fp(out))
    return out[:k].copy()


def voxel_grid(pts, leaf):
    """pcl::VoxelGrid with a cubic leaf (reference src/GraphicEnd.cpp:287-295); None when PCL would refuse (index overflow)."""
    pts = _f4(pts)
    out = np.empty((max(len(pts), 1), 4), np.float32)
    k = lib().oracle_voxel_grid(_fp(pts), len(pts), C.c_float(leaf), _fp(out))
    return None if k < 0 else out[:k].copy()


def transform(pts, T):
    """pcl::transformPointCloud with the float32 cast of T (reference src/saveOutput.cpp:87)."""
    pts = _f4(pts)
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
    out = np.empty((max(len(pts), 1), 4), np.float32)
    lib().oracle_transform(_fp(pts), len(pts), T.ctypes.data_as(C.POINTER(C.c_double)), _fp(out))
    return out[:len(pts)].copy()


def map_fuse(clouds, poses, leaf, z_max):
    """The key-frame fusion loop of saveOutput (reference src/saveOutput.cpp:47-95)."""
    parts = [transform(passthrough_z(voxel_grid(c, leaf), 0.0, z_max), T) for c, T in zip(clouds, poses)]
    return voxel_grid(np.concatenate(parts, axis=0), leaf)


def backproject_normals(depth, cam, z_max=0.0, step=1, max_jump=0.05):
    """Depth -> (cloud (n,4), normals (n,4) with w = valid) like s3d_cloud_from_depth_normals."""
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    pts = np.empty((depth.size, 4), np.float32)
    nrm = np.empty((depth.size, 4), np.float32)
    camc = _abi.camera_c(cam)
    k = lib().oracle_backproject_normals(depth.ctypes.data_as(C.POINTER(C.c_uint16)), depth.shape[1], depth.shape[0], C.byref(camc),
                                         C.c_float(z_max), int(step), C.c_float(max_jump), _fp(pts), _fp(nrm))
    return pts[:k].copy(), nrm[:k].copy()


def pose_norm(T):
    T = np.ascontiguousarray(T, dtype=np.float64)
    return lib().oracle_pose_norm(T.ctypes.data_as(C.POINTER(C.c_double)))


def max_threads() -> int:
    return lib().oracle_max_threads()

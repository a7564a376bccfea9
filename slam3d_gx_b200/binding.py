"""ctypes binding of libslam3d_b200.so (include/slam3d_b200.h).  No fallbacks: a missing library or a
missing GPU is an error."""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class LibraryMissing(RuntimeError):
    pass


class S3DError(RuntimeError):
    pass


def library_path() -> str:
    # S3D_LIBRARY: developer override to A/B-test an alternative build of the same CUDA library
    return os.environ.get("S3D_LIBRARY") or os.path.join(_HERE, "libslam3d_b200.so")


# every symbol include/slam3d_b200.h declares (tests check that the .so exports all of them)
EXPORTS = [
    "s3d_abi_version", "s3d_create", "s3d_destroy", "s3d_last_error", "s3d_set_stream", "s3d_device_sm_count",
    "s3d_launch_count", "s3d_cloud_upload", "s3d_cloud_from_device", "s3d_cloud_from_depth", "s3d_cloud_set_normals",
    "s3d_cloud_set_normals_device", "s3d_cloud_size", "s3d_cloud_has_normals", "s3d_cloud_download",
    "s3d_cloud_drop_index", "s3d_cloud_free", "s3d_segment_planes", "s3d_register_batch", "s3d_register_pair",
    "s3d_last_correspondences", "s3d_last_timing", "s3d_icp_params_default", "s3d_plane_params_default",
    "s3d_planar_keypoints", "s3d_gather_results", "s3d_cloud_passthrough_z", "s3d_cloud_voxel_grid", "s3d_cloud_transform",
    "s3d_cloud_concat", "s3d_map_fuse", "s3d_cloud_from_depth_normals", "s3d_cloud_upload_async", "s3d_cloud_wait", "s3d_host_alloc", "s3d_host_free",
    "s3d_memory_stats", "s3d_last_plane_timing", "s3d_comm_unique_id", "s3d_comm_create", "s3d_comm_destroy", "s3d_register_batch_gather", "s3d_register_enqueue", "s3d_register_drain", "s3d_segment_planes_enqueue", "s3d_segment_planes_drain", "s3d_cloud_release", "s3d_batch_shape",
]
ASYNC_DEPTH = 64          # S3D_ASYNC_DEPTH
MAX_PLANES = 16           # S3D_MAX_PLANES
COMM_ID_BYTES = 128


def load_library():
    """dlopen the CUDA library.  Raises LibraryMissing if it was not built (run __graft_entry__.build())."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise LibraryMissing(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, ci = C.c_void_p, C.c_int
    lib.s3d_abi_version.restype = ci
    lib.s3d_create.argtypes = [C.POINTER(vp), ci]
    lib.s3d_destroy.argtypes = [vp]
    lib.s3d_destroy.restype = None
    lib.s3d_last_error.argtypes = [vp]
    lib.s3d_last_error.restype = C.c_char_p
    lib.s3d_set_stream.argtypes = [vp, vp]
    lib.s3d_device_sm_count.argtypes = [vp]
    lib.s3d_launch_count.argtypes = [vp]
    lib.s3d_launch_count.restype = C.c_int64
    lib.s3d_cloud_upload.argtypes = [vp, vp, ci, ci, C.POINTER(vp)]
    lib.s3d_cloud_upload_async.argtypes = [vp, vp, ci, ci, C.POINTER(vp)]
    lib.s3d_cloud_wait.argtypes = [vp, vp]
    lib.s3d_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.s3d_host_free.argtypes = [vp, vp]
    lib.s3d_host_free.restype = None
    lib.s3d_cloud_from_device.argtypes = [vp, vp, ci, C.POINTER(vp)]
    lib.s3d_cloud_from_depth.argtypes = [vp, vp, ci, ci, C.POINTER(_abi.CameraC), C.c_float, C.POINTER(vp)]
    lib.s3d_cloud_set_normals.argtypes = [vp, vp, vp, ci, ci]
    lib.s3d_cloud_set_normals_device.argtypes = [vp, vp, vp, ci]
    lib.s3d_cloud_size.argtypes = [vp]
    lib.s3d_cloud_has_normals.argtypes = [vp]
    lib.s3d_cloud_download.argtypes = [vp, vp, vp, vp, vp]
    lib.s3d_cloud_drop_index.argtypes = [vp, vp]
    lib.s3d_cloud_free.argtypes = [vp, vp]
    lib.s3d_cloud_free.restype = None
    lib.s3d_cloud_release.argtypes = [vp, vp]
    lib.s3d_cloud_release.restype = None
    lib.s3d_batch_shape.argtypes = [ci, ci, ci, C.POINTER(ci), C.POINTER(ci)]
    lib.s3d_segment_planes.argtypes = [vp, vp, C.POINTER(_abi.PlaneParams), C.POINTER(_abi.Plane), C.POINTER(ci)]
    lib.s3d_register_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), vp, ci, C.POINTER(_abi.IcpParams),
                                       C.POINTER(_abi.Result)]
    lib.s3d_register_pair.argtypes = [vp, vp, vp, vp, C.POINTER(_abi.IcpParams), C.POINTER(_abi.Result)]
    lib.s3d_last_correspondences.argtypes = [vp, vp, ci]
    lib.s3d_last_timing.argtypes = [vp, C.POINTER(_abi.Timing)]
    lib.s3d_register_enqueue.argtypes = [vp, vp, vp, vp, C.POINTER(_abi.IcpParams)]
    lib.s3d_register_drain.argtypes = [vp, C.POINTER(_abi.Result), C.POINTER(_abi.Timing), ci, C.POINTER(ci)]
    lib.s3d_segment_planes_enqueue.argtypes = [vp, vp, C.POINTER(_abi.PlaneParams)]
    lib.s3d_segment_planes_drain.argtypes = [vp, C.POINTER(_abi.Plane), C.POINTER(ci), ci, C.POINTER(ci)]
    lib.s3d_icp_params_default.argtypes = [C.POINTER(_abi.IcpParams)]
    lib.s3d_icp_params_default.restype = None
    lib.s3d_plane_params_default.argtypes = [C.POINTER(_abi.PlaneParams)]
    lib.s3d_plane_params_default.restype = None
    lib.s3d_planar_keypoints.argtypes = [vp, vp, ci, ci, C.POINTER(_abi.CameraC), vp, ci, C.c_float, ci, C.c_uint64, vp]
    lib.s3d_gather_results.argtypes = [vp, vp, C.POINTER(_abi.Result), ci, ci, C.POINTER(_abi.Result)]
    lib.s3d_last_plane_timing.argtypes = [vp, C.POINTER(_abi.PlaneTiming)]
    lib.s3d_memory_stats.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.s3d_comm_unique_id.argtypes = [vp]
    lib.s3d_comm_create.argtypes = [vp, vp, ci, ci, C.POINTER(vp)]
    lib.s3d_comm_destroy.argtypes = [vp, vp]
    lib.s3d_comm_destroy.restype = None
    lib.s3d_register_batch_gather.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp), vp, ci, ci, C.POINTER(_abi.IcpParams), ci,
                                              C.POINTER(_abi.Result)]
    lib.s3d_cloud_from_depth_normals.argtypes = [vp, vp, ci, ci, C.POINTER(_abi.CameraC), C.c_float, ci, C.c_float, C.POINTER(vp)]
    lib.s3d_cloud_passthrough_z.argtypes = [vp, vp, C.c_float, C.c_float, C.POINTER(vp)]
    lib.s3d_cloud_voxel_grid.argtypes = [vp, vp, C.c_float, C.POINTER(vp)]
    lib.s3d_cloud_transform.argtypes = [vp, vp, vp, C.POINTER(vp)]
    lib.s3d_cloud_concat.argtypes = [vp, C.POINTER(vp), ci, C.POINTER(vp)]
    lib.s3d_map_fuse.argtypes = [vp, C.POINTER(vp), vp, ci, C.c_float, C.c_float, C.POINTER(vp)]
    _LIB = lib
    return lib


class Cloud:
    """Device-resident cloud handle (s3d_cloud*)."""

    def __init__(self, ctx: "Context", handle):
        self.ctx, self.handle = ctx, handle

    def __len__(self):
        return self.ctx.lib.s3d_cloud_size(self.handle)

    @property
    def has_normals(self) -> bool:
        return bool(self.ctx.lib.s3d_cloud_has_normals(self.handle))

    def wait(self):
        """Block until an asynchronous upload of this cloud has finished."""
        self.ctx._check(self.ctx.lib.s3d_cloud_wait(self.ctx.h, self.handle))

    def set_normals(self, normals: np.ndarray):
        a = np.ascontiguousarray(normals, dtype=np.float32)
        assert a.ndim == 2 and a.shape[1] >= 3 and a.shape[0] == len(self)
        self.ctx._check(self.ctx.lib.s3d_cloud_set_normals(self.ctx.h, self.handle, a.ctypes.data, a.shape[1], a.shape[0]))

    def set_normals_device(self, dptr: int, n: int):
        self.ctx._check(self.ctx.lib.s3d_cloud_set_normals_device(self.ctx.h, self.handle, C.c_void_p(dptr), n))

    def download(self, xyz=True, normals=False, labels=False):
        n = len(self)
        out = {}
        ax = np.empty((n, 3), np.float32) if xyz else None
        an = np.empty((n, 3), np.float32) if normals else None
        al = np.empty(n, np.int32) if labels else None
        self.ctx._check(self.ctx.lib.s3d_cloud_download(self.ctx.h, self.handle, ax.ctypes.data if xyz else None,
                                                        an.ctypes.data if normals else None,
                                                        al.ctypes.data if labels else None))
        if xyz:
            out["xyz"] = ax
        if normals:
            out["normals"] = an
        if labels:
            out["labels"] = al
        return out

    def segment_planes(self, params: _abi.PlaneParams | None = None):
        """GraphicEnd::extractPlanesAndGenerateImage (reference src/GraphicEnd.cpp:353-430) on the device."""
        params = params or _abi.plane_params()
        planes = (_abi.Plane * max(1, params.max_planes))()
        k = C.c_int(0)
        self.ctx._check(self.ctx.lib.s3d_segment_planes(self.ctx.h, self.handle, C.byref(params), planes, C.byref(k)))
        return [dict(coef=np.array(list(planes[i].coef), np.float32), inliers=planes[i].inliers,
                     hypotheses=planes[i].hypotheses) for i in range(k.value)]

    def segment_planes_enqueue(self, params: _abi.PlaneParams | None = None):
        """The extraction issued without waiting for it (s3d_segment_planes_enqueue); planes come back from Context.planes_drain."""
        params = params or _abi.plane_params()
        self.ctx._check(self.ctx.lib.s3d_segment_planes_enqueue(self.ctx.h, self.handle, C.byref(params)))

    def drop_index(self):
        self.ctx.lib.s3d_cloud_drop_index(self.ctx.h, self.handle)

    def passthrough_z(self, z_min: float, z_max: float) -> "Cloud":
        """pcl::PassThrough on z (reference src/GraphicEnd.cpp:283-285)."""
        h = C.c_void_p()
        self.ctx._check(self.ctx.lib.s3d_cloud_passthrough_z(self.ctx.h, self.handle, z_min, z_max, C.byref(h)))
        return Cloud(self.ctx, h)

    def voxel_grid(self, leaf: float) -> "Cloud":
        """pcl::VoxelGrid with a cubic leaf (reference src/GraphicEnd.cpp:287-295, parameter grid_leaf)."""
        h = C.c_void_p()
        self.ctx._check(self.ctx.lib.s3d_cloud_voxel_grid(self.ctx.h, self.handle, leaf, C.byref(h)))
        return Cloud(self.ctx, h)

    def transform(self, T) -> "Cloud":
        """pcl::transformPointCloud (reference src/saveOutput.cpp:87)."""
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        h = C.c_void_p()
        self.ctx._check(self.ctx.lib.s3d_cloud_transform(self.ctx.h, self.handle, T.ctypes.data, C.byref(h)))
        return Cloud(self.ctx, h)

    def free(self):
        if self.handle is not None:
            self.ctx.lib.s3d_cloud_free(self.ctx.h, self.handle)
            self.handle = None

    def release(self):
        """Like free() but without waiting for the context's stream (s3d_cloud_release): for clouds with work still enqueued."""
        if self.handle is not None:
            self.ctx.lib.s3d_cloud_release(self.ctx.h, self.handle)
            self.handle = None


def batch_shape(n_pairs: int, n_points_max: int, resident_ctas: int = 148):
    """(groups, CTAs per group) of a batch (s3d_batch_shape; no GPU needed)."""
    g, c = C.c_int(0), C.c_int(0)
    rc = load_library().s3d_batch_shape(n_pairs, n_points_max, resident_ctas, C.byref(g), C.byref(c))
    if rc != 0:
        raise S3DError(f"s3d_batch_shape rc={rc}")
    return g.value, c.value


class Context:
    """Per-GPU context (s3d_ctx*)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.s3d_create(C.byref(h), device)
        if rc != 0:
            raise S3DError(f"s3d_create(device={device}) failed rc={rc}: a CUDA device is required (no CPU fallback)")
        self.h = h

    def _check(self, rc: int):
        if rc != 0:
            raise S3DError(f"rc={rc}: {self.lib.s3d_last_error(self.h).decode()}")

    def close(self):
        if self.h is not None:
            self.lib.s3d_destroy(self.h)
            self.h = None

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.s3d_set_stream(self.h, C.c_void_p(cuda_stream)))

    @property
    def sm_count(self) -> int:
        return self.lib.s3d_device_sm_count(self.h)

    @property
    def launch_count(self) -> int:
        return self.lib.s3d_launch_count(self.h)

    def upload(self, xyz: np.ndarray, normals: np.ndarray | None = None) -> Cloud:
        a = np.ascontiguousarray(xyz, dtype=np.float32)
        assert a.ndim == 2 and a.shape[1] >= 3
        h = C.c_void_p()
        self._check(self.lib.s3d_cloud_upload(self.h, a.ctypes.data, a.shape[1], a.shape[0], C.byref(h)))
        c = Cloud(self, h)
        if normals is not None:
            c.set_normals(normals)
        return c

    def upload_async(self, xyz: np.ndarray) -> Cloud:
        """Upload on the context's copy stream without waiting (s3d_cloud_upload_async): the next frame crosses PCIe while
        the current one is registered.  `xyz` must be float32, C-contiguous (ideally page-locked) and is kept alive by
        the returned Cloud; do not modify it before the cloud has been used or `wait()`ed for."""
        assert xyz.dtype == np.float32 and xyz.ndim == 2 and xyz.shape[1] >= 3 and xyz.flags["C_CONTIGUOUS"]
        h = C.c_void_p()
        self._check(self.lib.s3d_cloud_upload_async(self.h, xyz.ctypes.data, xyz.shape[1], xyz.shape[0], C.byref(h)))
        c = Cloud(self, h)
        c._host = xyz
        return c

    def from_device(self, dptr: int, n: int) -> Cloud:
        h = C.c_void_p()
        self._check(self.lib.s3d_cloud_from_device(self.h, C.c_void_p(dptr), n, C.byref(h)))
        return Cloud(self, h)

    def from_depth(self, depth: np.ndarray, cam, z_max: float = 0.0) -> Cloud:
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        camc = _abi.camera_c(cam)
        h = C.c_void_p()
        self._check(self.lib.s3d_cloud_from_depth(self.h, d.ctypes.data, d.shape[1], d.shape[0], C.byref(camc),
                                                  C.c_float(z_max), C.byref(h)))
        return Cloud(self, h)

    def concat(self, clouds) -> Cloud:
        n = len(clouds)
        arr = (C.c_void_p * max(n, 1))(*[c.handle for c in clouds])
        h = C.c_void_p()
        self._check(self.lib.s3d_cloud_concat(self.h, arr, n, C.byref(h)))
        return Cloud(self, h)

    def map_fuse(self, clouds, poses, leaf: float, z_max: float) -> Cloud:
        """The key-frame fusion loop of saveOutput (reference src/saveOutput.cpp:47-95)."""
        n = len(clouds)
        arr = (C.c_void_p * n)(*[c.handle for c in clouds])
        P = np.ascontiguousarray(poses, dtype=np.float64).reshape(n, 16)
        h = C.c_void_p()
        self._check(self.lib.s3d_map_fuse(self.h, arr, P.ctypes.data, n, leaf, z_max, C.byref(h)))
        return Cloud(self, h)

    def from_depth_normals(self, depth: np.ndarray, cam, z_max: float = 0.0, step: int = 1, max_jump: float = 0.05) -> Cloud:
        """Depth image -> cloud with per-point normals from the organised image."""
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        camc = _abi.camera_c(cam)
        h = C.c_void_p()
        self._check(self.lib.s3d_cloud_from_depth_normals(self.h, d.ctypes.data, d.shape[1], d.shape[0], C.byref(camc),
                                                          C.c_float(z_max), step, C.c_float(max_jump), C.byref(h)))
        return Cloud(self, h)

    def register_batch(self, srcs, tgts, guess=None, params: _abi.IcpParams | None = None, raw: bool = False):
        """Batched GraphicEnd::multiPnP (reference src/GraphicEnd.cpp:557-659): one result per (src, tgt)."""
        params = params or _abi.icp_params()
        n = len(srcs)
        assert n == len(tgts) and n > 0
        sa = (C.c_void_p * n)(*[s.handle for s in srcs])
        ta = (C.c_void_p * n)(*[t.handle for t in tgts])
        g = None
        if guess is not None:
            g = np.ascontiguousarray(guess, dtype=np.float64).reshape(n, 16)
        res = (_abi.Result * n)()
        self._check(self.lib.s3d_register_batch(self.h, sa, ta, g.ctypes.data if g is not None else None, n,
                                                C.byref(params), res))
        if raw:
            return res
        return [_abi.result_to_dict(res[i]) for i in range(n)]

    def register_enqueue(self, src, tgt, guess=None, params: _abi.IcpParams | None = None):
        """One registration issued without waiting for it (s3d_register_enqueue); at most ASYNC_DEPTH outstanding."""
        params = params or _abi.icp_params()
        g = None
        if guess is not None:
            g = np.ascontiguousarray(guess, dtype=np.float64).reshape(16)
        self._check(self.lib.s3d_register_enqueue(self.h, src.handle, tgt.handle, g.ctypes.data if g is not None else None, C.byref(params)))

    def planes_drain(self):
        """Planes of every extraction enqueued since the last drain, in enqueue order (s3d_segment_planes_drain)."""
        planes = (_abi.Plane * (ASYNC_DEPTH * MAX_PLANES))()
        counts = (C.c_int * ASYNC_DEPTH)()
        n = C.c_int(0)
        self._check(self.lib.s3d_segment_planes_drain(self.h, planes, counts, ASYNC_DEPTH, C.byref(n)))
        return [[dict(coef=np.array(list(planes[i * MAX_PLANES + k].coef), np.float32), inliers=planes[i * MAX_PLANES + k].inliers,
                      hypotheses=planes[i * MAX_PLANES + k].hypotheses) for k in range(counts[i])] for i in range(n.value)]

    def register_drain(self, raw: bool = False):
        """Results (enqueue order) and device timings of everything enqueued since the last drain (s3d_register_drain)."""
        res = (_abi.Result * ASYNC_DEPTH)()
        tms = (_abi.Timing * ASYNC_DEPTH)()
        n = C.c_int(0)
        self._check(self.lib.s3d_register_drain(self.h, res, tms, ASYNC_DEPTH, C.byref(n)))
        timings = [dict(index_ms=tms[i].index_ms, iterate_ms=tms[i].iterate_ms, iter_launches=tms[i].iter_launches,
                        total_launches=tms[i].total_launches) for i in range(n.value)]
        if raw:
            out = (_abi.Result * n.value)()
            for i in range(n.value):
                out[i] = res[i]
            return out, timings
        return [_abi.result_to_dict(res[i]) for i in range(n.value)], timings

    # ---- multi-GPU: one shard per rank, one pose gather -------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte NCCL unique id (create on one rank, hand to the others, e.g. with torch.distributed broadcast)."""
        buf = C.create_string_buffer(COMM_ID_BYTES)
        rc = load_library().s3d_comm_unique_id(buf)
        if rc != 0:
            raise S3DError(f"s3d_comm_unique_id rc={rc}: NCCL not loadable")
        return buf.raw

    def comm_create(self, uid: bytes, world: int, rank: int):
        assert len(uid) == COMM_ID_BYTES
        h = C.c_void_p()
        self._check(self.lib.s3d_comm_create(self.h, C.c_char_p(uid), world, rank, C.byref(h)))
        return h

    def comm_destroy(self, comm):
        self.lib.s3d_comm_destroy(self.h, comm)

    def gather_results(self, comm, local, world: int):
        """s3d_gather_results: all-gather already computed records (ctypes array of s3d_result)."""
        n = len(local)
        out = (_abi.Result * (n * world))()
        self._check(self.lib.s3d_gather_results(self.h, comm, local, n, world, out))
        return out

    def register_batch_gather(self, comm, srcs, tgts, world: int, n_slot: int | None = None, guess=None,
                              params: _abi.IcpParams | None = None, raw: bool = False):
        """This rank's shard of a sharded batch + the pose gather in one device-side sequence (s3d_register_batch_gather).
        Returns world * n_slot records in rank order (padding slots have status PAIR_ABSENT)."""
        params = params or _abi.icp_params()
        n = len(srcs)
        assert n == len(tgts)
        n_slot = n if n_slot is None else n_slot
        sa = (C.c_void_p * max(n, 1))(*[s.handle for s in srcs])
        ta = (C.c_void_p * max(n, 1))(*[t.handle for t in tgts])
        g = None
        if guess is not None:
            g = np.ascontiguousarray(guess, dtype=np.float64).reshape(n, 16)
        res = (_abi.Result * (n_slot * world))()
        self._check(self.lib.s3d_register_batch_gather(self.h, comm, sa, ta, g.ctypes.data if g is not None else None, n, n_slot,
                                                       C.byref(params), world, res))
        if raw:
            return res
        return [_abi.result_to_dict(res[i]) for i in range(n_slot * world)]

    def memory_stats(self) -> dict:
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._check(self.lib.s3d_memory_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(live_bytes=a.value, peak_live_bytes=b.value, cached_bytes=c.value)

    def host_alloc(self, nbytes: int) -> int:
        h = C.c_void_p()
        self._check(self.lib.s3d_host_alloc(self.h, nbytes, C.byref(h)))
        return h.value

    def host_free(self, ptr: int):
        self.lib.s3d_host_free(self.h, C.c_void_p(ptr))

    def register(self, src: Cloud, tgt: Cloud, guess=None, params: _abi.IcpParams | None = None) -> dict:
        return self.register_batch([src], [tgt], None if guess is None else [guess], params)[0]

    def last_correspondences(self, n: int) -> np.ndarray:
        out = np.empty(n, np.int32)
        self._check(self.lib.s3d_last_correspondences(self.h, out.ctypes.data, n))
        return out

    def last_timing(self) -> dict:
        t = _abi.Timing()
        self.lib.s3d_last_timing(self.h, C.byref(t))
        return dict(index_ms=t.index_ms, iterate_ms=t.iterate_ms, iter_launches=t.iter_launches,
                    total_launches=t.total_launches)

    def last_plane_timing(self) -> dict:
        t = _abi.PlaneTiming()
        self.lib.s3d_last_plane_timing(self.h, C.byref(t))
        return dict(total_ms=t.total_ms, eval_ms=t.eval_ms, rounds=t.rounds, eval_passes_per_round=t.eval_passes_per_round,
                    points_scanned=t.points_scanned)

    def planar_keypoints(self, depth: np.ndarray, cam, uv: np.ndarray, threshold=0.01, min_inliers=40, seed=12345):
        """isPlanar (reference src/planarFeatures.cpp:88-136) for a batch of keypoints."""
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        uv = np.ascontiguousarray(uv, dtype=np.int32).reshape(-1, 2)
        flags = np.zeros(len(uv), np.uint8)
        camc = _abi.camera_c(cam)
        self._check(self.lib.s3d_planar_keypoints(self.h, d.ctypes.data, d.shape[1], d.shape[0], C.byref(camc),
                                                  uv.ctypes.data, len(uv), C.c_float(threshold), int(min_inliers),
                                                  C.c_uint64(seed), flags.ctypes.data))
        return flags

"""ctypes mirror of include/slam3d_b200.h (struct layouts shared by the product binding and the oracle binding)."""
import ctypes as C


class IcpParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("max_corr_dist", C.c_float), ("estimator", C.c_int32),
                ("search", C.c_int32), ("grid_cell", C.c_float), ("min_correspondences", C.c_int32),
                ("pivot_eps", C.c_double), ("reuse_index", C.c_int32), ("reserved", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("T", C.c_double * 16), ("norm", C.c_double), ("fitness", C.c_double),
                ("inliers", C.c_int32), ("iterations", C.c_int32), ("status", C.c_int32),
                ("reserved", C.c_int32)]


class PlaneParams(C.Structure):
    _fields_ = [("distance_threshold", C.c_float), ("plane_percent", C.c_float), ("max_planes", C.c_int32),
                ("max_iterations", C.c_int32), ("probability", C.c_float), ("reserved", C.c_int32),
                ("seed", C.c_uint64)]


class Plane(C.Structure):
    _fields_ = [("coef", C.c_float * 4), ("inliers", C.c_int32), ("hypotheses", C.c_int32)]


class CameraC(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("factor", C.c_double)]


class Timing(C.Structure):
    _fields_ = [("index_ms", C.c_float), ("iterate_ms", C.c_float), ("iter_launches", C.c_int32),
                ("total_launches", C.c_int32)]


class PlaneTiming(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("eval_ms", C.c_float), ("rounds", C.c_int32), ("eval_passes_per_round", C.c_int32),
                ("points_scanned", C.c_int64), ("reserved", C.c_int64)]


ESTIMATOR_POINT_TO_PLANE, ESTIMATOR_SVD = 0, 1
SEARCH_GRID, SEARCH_BRUTE, SEARCH_GRID_LANE = 0, 1, 2
PAIR_OK, PAIR_FEW, PAIR_DEGENERATE, PAIR_NONFINITE = 0, 1, 2, 3
PAIR_ABSENT = -1
PLANE_CANDIDATES_EXTRA = 14
RESULT_BYTES = C.sizeof(Result)


def icp_params(max_iterations=10, max_corr_dist=0.0, estimator=ESTIMATOR_POINT_TO_PLANE, search=SEARCH_GRID,
               grid_cell=0.0, min_correspondences=3, pivot_eps=0.0, reuse_index=1) -> IcpParams:
    return IcpParams(max_iterations, max_corr_dist, estimator, search, grid_cell, min_correspondences,
                     pivot_eps, reuse_index, 0)


def plane_params(distance_threshold=0.08, plane_percent=0.2, max_planes=3, max_iterations=50,
                 probability=0.99, seed=12345, timed=False) -> PlaneParams:
    """Defaults: reference parameters.yaml:41-47 and PCL-1.7 SACSegmentation.  timed: per-pass CUDA events instead of the graph replay."""
    return PlaneParams(distance_threshold, plane_percent, max_planes, max_iterations, probability, 1 if timed else 0, seed)


def camera_c(cam) -> CameraC:
    return CameraC(cam.fx, cam.fy, cam.cx, cam.cy, cam.factor)


def result_to_dict(r: Result) -> dict:
    import numpy as np
    return dict(T=np.array(list(r.T), dtype=np.float64).reshape(4, 4), norm=r.norm, fitness=r.fitness,
                inliers=r.inliers, iterations=r.iterations, status=r.status)

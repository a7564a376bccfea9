"""slam3d_gx_b200 -- B200-native planar-ICP registration path of gaoxiang12/slam3d_gx.

The product is the C-ABI shared library ``libslam3d_b200.so`` (hand-written sm_100a CUDA, declared in
``include/slam3d_b200.h``) plus the C++ host shell under ``slam3d_gx_b200/host`` that keeps the reference's
``GraphicEnd`` / ``ParameterReader`` surface.  This Python package is a thin ctypes binding used by the
tests and by ``bench.py``; it contains no CPU implementation of the path and raises if the CUDA library is
missing or no GPU is present.
"""
from .binding import (Context, Cloud, load_library, library_path, LibraryMissing, S3DError)  # noqa: F401
from ._abi import (icp_params, plane_params, ESTIMATOR_POINT_TO_PLANE, ESTIMATOR_SVD, SEARCH_GRID,  # noqa: F401
                   SEARCH_BRUTE, PAIR_OK, PAIR_FEW, PAIR_DEGENERATE, PAIR_NONFINITE)

__all__ = ["Context", "Cloud", "load_library", "library_path", "LibraryMissing", "S3DError", "icp_params",
           "plane_params"]

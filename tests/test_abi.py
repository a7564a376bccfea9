"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/slam3d_b200.h
declares, and the ctypes struct mirrors have the C layout.  No compute call is made (no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import _abi, binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "slam3d_b200.h")


def _declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(s3d_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = s3d.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/slam3d_b200.h but not exported"
    assert sorted(binding.EXPORTS) == declared
    assert lib.s3d_abi_version() == 1


def test_struct_layouts_match_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "slam3d_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(s3d_icp_params),sizeof(s3d_result),sizeof(s3d_plane_params),sizeof(s3d_plane),sizeof(s3d_camera),'
                   'sizeof(s3d_timing),offsetof(s3d_result,inliers),offsetof(s3d_icp_params,pivot_eps),sizeof(s3d_plane_timing),'
                   'offsetof(s3d_plane_timing,points_scanned));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True,
                   env={**os.environ, "CC": "gcc"})
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_abi.IcpParams), C.sizeof(_abi.Result), C.sizeof(_abi.PlaneParams), C.sizeof(_abi.Plane),
            C.sizeof(_abi.CameraC), C.sizeof(_abi.Timing), _abi.Result.inliers.offset, _abi.IcpParams.pivot_eps.offset,
            C.sizeof(_abi.PlaneTiming), _abi.PlaneTiming.points_scanned.offset]
    assert got == want


def test_defaults_come_from_the_library():
    lib = s3d.load_library()
    p = _abi.IcpParams()
    lib.s3d_icp_params_default(C.byref(p))
    assert (p.max_iterations, p.estimator, p.search, p.min_correspondences, p.reuse_index) == (10, 0, 0, 3, 1)
    q = _abi.PlaneParams()
    lib.s3d_plane_params_default(C.byref(q))
    # reference parameters.yaml: distance_threshold 0.08, plane_percent 0.2, max_planes 3; PCL: 50 iterations, p=0.99
    assert abs(q.distance_threshold - 0.08) < 1e-7 and abs(q.plane_percent - 0.2) < 1e-7
    assert (q.max_planes, q.max_iterations) == (3, 50)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(s3d.S3DError):
        s3d.Context(0)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "slam3d_gx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in txt.replace("oracle/oracle_common.h", "").replace("oracle/icp_oracle.c", "").replace("oracle/plane_oracle.c", "") \
                    or f.endswith((".cu", ".cuh", ".h")), f
                assert "import oracle" not in txt and "from oracle" not in txt, f"{f} imports the oracle"
                assert "liboracle" not in txt, f"{f} links the oracle"


def test_only_the_checkers_touch_the_oracle():
    """oracle/ is test infrastructure: besides tests/, only bench.py (cpu_baseline / reference arm) and
    __graft_entry__.py (smoke check, build of the checker) may import it -- tools/ and the host shell may not."""
    for sub in ("tools", "include", os.path.join("slam3d_gx_b200", "host")):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cpp", ".h", ".hpp", ".c")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, os.path.join(dirpath, f)


def test_batch_shape_rule():
    """s3d_batch_shape (host arithmetic of s3d_register_batch): a lone pair gets the whole chip, up to 16 pairs
    run all at once on equal groups, larger batches on four groups that walk the list; small clouds are not spread wider than
    two chunks of 32 queries per warp; the groups always fit the resident CTAs."""
    n = 307200
    assert binding.batch_shape(1, n) == (1, 148)
    assert binding.batch_shape(2, n) == (2, 74)
    assert binding.batch_shape(16, n) == (16, 9)
    assert binding.batch_shape(17, n) == (4, 37)
    assert binding.batch_shape(64, n) == (4, 37)
    assert binding.batch_shape(512, n) == (4, 37)
    assert binding.batch_shape(5, 1000) == (5, 1)            # a 1000-point cloud is one CTA's worth
    assert binding.batch_shape(200, 1000) == (148, 1)        # ... and then every CTA takes pairs
    for pairs in (1, 3, 16, 17, 100, 1000):
        for pts in (0, 1, 5000, 307200, 2000000):
            for res in (1, 7, 132, 148):
                g, c = binding.batch_shape(pairs, pts, res)
                assert 1 <= g <= pairs and c >= 1 and g * c <= max(res, 1)
    with pytest.raises(s3d.S3DError):
        binding.batch_shape(0, n)

"""Scratch: where the end-to-end step (host buffers -> pose) spends its time."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
ctx = s3d.Context(0)
p = synth.make_pair(0)
a = torch.from_numpy(p["src"].copy()).pin_memory(); b = torch.from_numpy(p["tgt"].copy()).pin_memory()
pp = _abi.plane_params(); prm = _abi.icp_params(30, reuse_index=0)
def step():
    t = [time.perf_counter()]
    cs = ctx.upload(a.numpy()); t.append(time.perf_counter())
    ct = ctx.upload(b.numpy()); t.append(time.perf_counter())
    planes = ct.segment_planes(pp); t.append(time.perf_counter())
    r = ctx.register_batch([cs], [ct], None, prm, raw=True)[0]; t.append(time.perf_counter())
    tm = ctx.last_timing()
    cs.free(); ct.free(); t.append(time.perf_counter())
    return np.diff(t) * 1e3, tm
for i in range(6):
    d, tm = step()
    print("upload src %.2f  upload tgt %.2f  segment %.2f  register %.2f (index %.2f iterate %.2f)  free %.2f   total %.2f ms" % (d[0], d[1], d[2], d[3], tm["index_ms"], tm["iterate_ms"], d[4], d.sum()), flush=True)

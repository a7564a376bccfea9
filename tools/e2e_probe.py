"""Scratch: where the end-to-end step (host buffers -> pose) spends its time, blocking and pipelined upload."""
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
def step(nxt, mode):
    t = [time.perf_counter()]
    if mode == "block":
        cs = ctx.upload(a.numpy()); ct = ctx.upload(b.numpy())
    else:
        cs, ct = nxt
        nxt = (ctx.upload_async(a.numpy()), ctx.upload_async(b.numpy()))
    t.append(time.perf_counter())
    planes = ct.segment_planes(pp); t.append(time.perf_counter())
    r = ctx.register_batch([cs], [ct], None, prm, raw=True)[0]; t.append(time.perf_counter())
    tm = ctx.last_timing()
    cs.free(); ct.free(); t.append(time.perf_counter())
    return np.diff(t) * 1e3, tm, nxt
for mode in ("block", "async"):
    nxt = (ctx.upload_async(a.numpy()), ctx.upload_async(b.numpy())) if mode == "async" else None
    for i in range(14):
        d, tm, nxt = step(nxt, mode)
        if i >= 8:
            print("%s: upload %.3f  segment %.3f  register %.3f (index %.3f iterate %.3f)  free %.3f   total %.3f ms" % (mode, d[0], d[1], d[2], tm["index_ms"], tm["iterate_ms"], d[3], d.sum()), flush=True)
    if nxt:
        nxt[0].free(); nxt[1].free()

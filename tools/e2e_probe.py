"""Where the end-to-end step (host buffers -> pose) spends its time, blocking and pipelined upload.
argv: [npairs] [torch]  -- 'torch' runs the ctx on torch's current stream like bench.py does."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = s3d.Context(0)
if "torch" in sys.argv:
    torch.cuda.set_device(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
pin = []
for k in range(npairs):
    p = synth.make_pair(k)
    pin.append((torch.from_numpy(p["src"].copy()).pin_memory(), torch.from_numpy(p["tgt"].copy()).pin_memory()))
pp = _abi.plane_params(); prm = _abi.icp_params(30, reuse_index=0)
def up(i, mode):
    a, b = pin[i % npairs]
    if mode == "block":
        return ctx.upload(a.numpy()), ctx.upload(b.numpy())
    return ctx.upload_async(a.numpy()), ctx.upload_async(b.numpy())
def step(i, nxt, mode):
    t = [time.perf_counter()]
    if mode == "block":
        cs, ct = up(i, mode)
    else:
        cs, ct = nxt
        nxt = up(i + 1, mode)
    t.append(time.perf_counter())
    planes = ct.segment_planes(pp); t.append(time.perf_counter())
    r = ctx.register_batch([cs], [ct], None, prm, raw=True)[0]; t.append(time.perf_counter())
    tm = ctx.last_timing()
    cs.free(); ct.free(); t.append(time.perf_counter())
    return np.diff(t) * 1e3, tm, nxt
for mode in ("block", "async"):
    nxt = up(0, mode) if mode == "async" else None
    acc = []
    t0 = None
    for i in range(8 + 16):
        if i == 8:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        d, tm, nxt = step(i, nxt, mode)
        if i >= 8:
            acc.append(list(d) + [tm["index_ms"], tm["iterate_ms"]])
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 16 * 1e3
    m = np.mean(acc, axis=0)
    print("%s: upload %.3f  segment %.3f  register %.3f (index %.3f iterate %.3f)  free %.3f   sum %.3f ms  wall/step %.3f ms" % (mode, m[0], m[1], m[2], m[4], m[5], m[3], m[:4].sum(), wall), flush=True)
    if nxt:
        nxt[0].free(); nxt[1].free()

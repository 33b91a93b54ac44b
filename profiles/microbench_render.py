#!/usr/bin/env python
"""render_spherical forward: the plain kernel against the empty-space-skipping one on three kinds of input at B=16
(GenRe-like shells from cam_bp of sphere depth maps, cam_bp of per-pixel random depth = a volume with occupied voxels in
every brick, and an empty volume).  NCU=1: run ONE skip launch on the shells between cudaProfilerStart/Stop."""
import json, os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import genre_shapehd_b200
genre_shapehd_b200.install()
from genre_shapehd_b200 import _lib
from genre_shapehd_b200.synth import sphere_depth, uniform_depth
from toolbox.cam_bp.cam_bp.modules.camera_backprojection_module import Camera_back_projection_layer
from toolbox.spherical_proj import render_spherical, render_forward
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
B, R, S, Z = 16, 128, 128, 256
m = render_spherical().to(dev)
dirs = m._dirs_on(dev)
layer = Camera_back_projection_layer()

def vol(kind):
    if kind == "empty":
        return torch.full((B, 1, R, R, R), 1e-5, device=dev)
    d = np.stack([sphere_depth(radius=0.3 + 0.01 * i) if kind == "shells" else uniform_depth(i) for i in range(B)])[:, None]
    return torch.clamp(layer(torch.from_numpy(d).to(dev)) * 50, 1e-5, 1 - 1e-5)

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

out = torch.empty(B, 1, S, S, device=dev)
st = torch.cuda.current_stream().cuda_stream
res = {}
if os.environ.get("NCU"):
    v = vol("shells")
    for _ in range(2): render_forward(v, B, R, dirs, S, Z, m.depth_weight, out)
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    render_forward(v, B, R, dirs, S, Z, m.depth_weight, out)
    _lib.call("genre_b200_render_spherical_forward", v.data_ptr(), B, R, dirs.data_ptr(), S, Z, m.depth_weight.data_ptr(), out.data_ptr(), st)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
    sys.exit(0)
for kind in ("shells", "noise", "empty"):
    v = vol(kind)
    occ_frac = float((v > 1e-5).float().mean())
    plain = timeit(lambda: _lib.call("genre_b200_render_spherical_forward", v.data_ptr(), B, R, dirs.data_ptr(), S, Z,
                                     m.depth_weight.data_ptr(), out.data_ptr(), st))
    a = out.clone()
    skip = timeit(lambda: render_forward(v, B, R, dirs, S, Z, m.depth_weight, out))
    nbytes = _lib.load().genre_b200_render_spherical_workspace_bytes(B, R)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    render_forward(v, B, R, dirs, S, Z, m.depth_weight, out)
    # brick occupancy actually marked by the pre-pass
    _lib.call("genre_b200_render_spherical_forward_skip", v.data_ptr(), B, R, dirs.data_ptr(), S, Z, m.depth_weight.data_ptr(), 0, 1.0, 0.0, 0.0,
              out.data_ptr(), ws.data_ptr(), nbytes, st)
    bits = ws.view(torch.int32)
    marked = sum(bin(int(x) & 0xffffffff).count("1") for x in bits.cpu().tolist()) / (B * 4096)
    res[kind] = {"plain_us": plain, "skip_us": skip, "max_abs_diff": float((out - a).abs().max()), "occupied_voxel_frac": occ_frac,
                 "marked_brick_frac": marked, "alg_GBps_skip": B * (4 * R ** 3 + 4 * S * S) / (skip * 1e-6) / 1e9}
print(json.dumps(res))

#!/usr/bin/env python
"""Per-convolution device time of a voxel net's eval forward (B=16): custom kernels vs cuDNN (TF32 allowed)."""
import json, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import genre_shapehd_b200
genre_shapehd_b200.install()
from genre_shapehd_b200 import ops_conv
import networks.networks as nets
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
B = int(os.environ.get("B", 16))
out = {}
for name in sys.argv[1:] or ["VoxelGenerator", "VoxelDecoder", "VoxelDiscriminator"]:
    net = getattr(nets, name)().to(dev).eval()
    if name == "VoxelDiscriminator":
        x = torch.rand(B, 1, 128, 128, 128, device=dev)
    elif name == "VoxelGenerator":
        x = torch.randn(B, 200, 1, 1, 1, device=dev)
    else:
        x = torch.randn(B, 200, device=dev)
    times = {}
    def pre(n):
        def f(m, i):
            e = torch.cuda.Event(enable_timing=True); e.record(); times.setdefault(n, []).append([e, None])
        return f
    def post(n):
        def f(m, i, o):
            e = torch.cuda.Event(enable_timing=True); e.record(); times[n][-1][1] = e
        return f
    for n, m in net.named_modules():
        if not list(m.children()) and not isinstance(m, torch.nn.Sequential):
            m.register_forward_pre_hook(pre("%s:%s" % (n, type(m).__name__))); m.register_forward_hook(post("%s:%s" % (n, type(m).__name__)))
    res = {}
    for mode, enabled in (("custom", True), ("cudnn_tf32", False)):
        ops_conv.ENABLED = enabled; torch.backends.cudnn.allow_tf32 = True
        with torch.no_grad():
            for _ in range(3): net(x)
            times.clear()
            for _ in range(5): net(x)
        torch.cuda.synchronize()
        res[mode] = {k: round(sum(a.elapsed_time(b) for a, b in v) / len(v), 3) for k, v in times.items()}
        res[mode]["total"] = round(sum(res[mode].values()), 3)
    out[name] = res
print(json.dumps(out))

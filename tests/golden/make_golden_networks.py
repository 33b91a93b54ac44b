#!/usr/bin/env python
"""Golden digests of the reference's 3D networks (networks/networks.py), generated HERE (authoring container) by
importing the reference module from /root/reference on torch CPU fp32 — it is pure torch.nn and runs unmodified.
A network is built under a fixed seed, fed a seeded input, and its state_dict layout, parameter bytes (sha256) and
a sampled signature of the output (eval and train mode) are recorded.  tests/test_networks.py rebuilds the same
network from this repo's drop-in under the same seed and must reproduce all of it.

    python tests/golden/make_golden_networks.py        # writes tests/golden/networks_digest.json
"""
import hashlib
import importlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GENRE_REF", "/root/reference")
pkg = types.ModuleType("refpkg")
pkg.__path__ = [os.path.join(REF, "networks")]
sys.modules["refpkg"] = pkg
ref = importlib.import_module("refpkg.networks")

torch.set_num_threads(max(1, os.cpu_count() or 1))
torch.backends.mkldnn.enabled = True

CASES = {
    "Unet_3D": (dict(), (2, 2, 128, 128, 128)),
    "VoxelDecoder": (dict(), (2, 200)),
    "VoxelGenerator": (dict(), (2, 200, 1, 1, 1)),
    "VoxelDiscriminator": (dict(), (1, 1, 128, 128, 128)),
    "VoxelGenerator_res64": (dict(nz=200, nf=64, bias=False, res=64), (2, 200, 1, 1, 1)),
    "VoxelDiscriminator_res64": (dict(nf=64, bias=False, res=64), (2, 1, 64, 64, 64)),
}


def signature(t, n=64):
    flat = t.detach().reshape(-1).double()
    idx = torch.linspace(0, flat.numel() - 1, n).long()
    return {"shape": list(t.shape), "sum": float(flat.sum()), "abs_sum": float(flat.abs().sum()),
            "samples": [float(v) for v in flat[idx]]}


out = {"torch": torch.__version__, "source": "reference networks/networks.py on torch CPU fp32", "cases": {}}
for name, (kw, in_shape) in CASES.items():
    cls = getattr(ref, name.split("_res")[0])
    torch.manual_seed(1234)
    net = cls(**kw)
    sd = net.state_dict()
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v.numpy()).tobytes())
    torch.manual_seed(99)
    x = torch.rand(*in_shape) if "Discriminator" in name or name == "Unet_3D" else torch.randn(*in_shape)
    net.eval()
    with torch.no_grad():
        y_eval = net(x)
    net.train()
    with torch.no_grad():
        y_train = net(x)
    out["cases"][name] = {"kwargs": kw, "input_shape": list(in_shape),
                          "state_dict": [[k, list(v.shape)] for k, v in sd.items()], "params_sha256": h.hexdigest(),
                          "n_params": int(sum(p.numel() for p in net.parameters())),
                          "eval": signature(y_eval), "train": signature(y_train)}
    print(name, out["cases"][name]["n_params"], out["cases"][name]["eval"]["sum"])
json.dump(out, open(os.path.join(HERE, "networks_digest.json"), "w"), indent=1)

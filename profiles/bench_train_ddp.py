#!/usr/bin/env python
"""BASELINE configs[3] (ShapeHD fine-tune step and 3D-WGAN-GP critic step, batch 8 per GPU, DDP over NCCL) on the
networks drop-in.  The 2D ImageEncoder (ResNet-18) is outside the hot path and not available on the GPU box, so the
decoder is fed random 200-d codes directly; everything 3D is the reference's graph:
  shapehd step : models/shapehd.py:67-79,113-118  VoxelDecoder -> sigmoid -> frozen D; loss = BCE + w * -mean(D(.))
  wgangp D step: models/wgangp.py:77-142,144-164  D(real), D(G(z)) and the gradient penalty (double backward through D),
                 accumulated with no_sync() so the three backward() calls cost one all-reduce
Under autograd the covered layers run their forward on the custom kernels and their backward on cuDNN / the custom
weight-gradient kernels (ops_conv._ConvForward, _ConvTC1Train); GENRE_B200_CONV_TRAIN_FORWARD=0 gives the all-cuDNN step.
NCU=shapehd|wgan: run one warm step of that kind between cudaProfilerStart/Stop (for `ncu --profile-from-start off`).

    torchrun --nproc-per-node N profiles/bench_train_ddp.py [--steps K]          # one JSON line from rank 0
"""
import argparse, contextlib, json, os, sys, time
import torch
import torch.nn as nn
import torch.nn.functional as F
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import genre_shapehd_b200
genre_shapehd_b200.install()
from genre_shapehd_b200 import dist_util
import networks.networks as nets
from torch.nn.parallel import DistributedDataParallel as DDP

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--batch", type=int, default=8)
args = ap.parse_args()
world, rank, local = dist_util.env_world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist_util.init("nccl", dev)
torch.manual_seed(1 + rank)
B = args.batch
wrap = (lambda m: DDP(m, device_ids=[local])) if world > 1 else (lambda m: m)

# ---- ShapeHD fine-tune step ---------------------------------------------------------------------------------------
dec = wrap(nets.VoxelDecoder().to(dev))
D_frozen = nets.VoxelDiscriminator().to(dev).eval()
for p in D_frozen.parameters():
    p.requires_grad_(False)
opt = torch.optim.Adam(dec.parameters(), lr=1e-3)
codes = torch.randn(B, 200, device=dev)
target = (torch.rand(B, 1, 128, 128, 128, device=dev) < 0.05).float()

def shapehd_step():
    opt.zero_grad(set_to_none=True)
    logits = dec(codes)
    loss = F.binary_cross_entropy_with_logits(logits, target) - 1e-3 * D_frozen(torch.sigmoid(logits)).mean()
    loss.backward()
    opt.step()
    return loss

# ---- WGAN-GP critic step ---------------------------------------------------------------------------------------------
G = nets.VoxelGenerator().to(dev)
Dn = wrap(nets.VoxelDiscriminator().to(dev))
opt_d = torch.optim.Adam(Dn.parameters(), lr=1e-4, betas=(0.5, 0.9))
real = target
noise = torch.randn(B, 200, 1, 1, 1, device=dev)

def wgan_d_step():
    opt_d.zero_grad(set_to_none=True)
    nosync = Dn.no_sync if world > 1 else contextlib.nullcontext
    with nosync():
        (-Dn(real).mean()).backward()
        with torch.no_grad():
            fake = G(noise)
        Dn(fake).mean().backward()
    eps = torch.rand(B, 1, 1, 1, 1, device=dev)
    inter = (eps * real + (1 - eps) * fake).requires_grad_(True)
    out = Dn(inter)
    grad, = torch.autograd.grad(out.sum(), inter, create_graph=True)
    gp = 10.0 * ((grad.reshape(B, -1).norm(2, dim=1) - 1) ** 2).mean()
    gp.backward()
    opt_d.step()
    return gp

def bench(fn):
    for _ in range(args.warmup):
        fn()
    dist_util.barrier(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        last = fn()
    e1.record()
    dist_util.barrier(dev)
    ms = dist_util.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    return ms, float(last)

if os.environ.get("NCU"):
    fn = shapehd_step if os.environ["NCU"] == "shapehd" else wgan_d_step
    for _ in range(2): fn()
    torch.cuda.synchronize(); torch.cuda.profiler.start(); fn(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
    sys.exit(0)
from genre_shapehd_b200 import ops_conv
res = {"n_gpus": world, "batch_per_gpu": B, "steps": args.steps, "train_forward_custom": ops_conv.TRAIN_FORWARD}
ms, loss = bench(shapehd_step)
res["shapehd_step_ms"] = ms
res["shapehd_shapes_per_s"] = world * B / ms * 1e3
res["shapehd_loss_finite"] = bool(loss == loss)
ms, gp = bench(wgan_d_step)
res["wgangp_d_step_ms"] = ms
res["wgangp_shapes_per_s"] = world * B / ms * 1e3
res["wgangp_gp_finite"] = bool(gp == gp)
if rank == 0:
    print(json.dumps(res), flush=True)
dist_util.finalize()

"""Synthetic GenRe workload for BASELINE configs[2]: inputs and a random initialisation of the frozen
``models/genre_full_model.Net`` under which the predicted depth actually lands inside the voxel grid.

A plain random init makes net1's min/max head predict ~(0, 0): every pixel then unprojects 2.2 units in front of the
grid, cam_bp hits nothing and the 3D path runs on an empty volume.  There are no checkpoints offline, so the head's
last bias is set to the dataset's depth range and the depth decoder's last layer is scaled up so that the relative
depth varies over the silhouette (gain 10: a surface with ~2 voxels of relief; a trained net1 predicts smooth surfaces).  Everything else keeps PyTorch's default initialisation.
"""
import types

import torch


def genre_opt(joint_train=False):
    """the option fields models/genre_full_model.Net and depth_pred_with_sph_inpaint.Net read (their argparse defaults)"""
    import argparse
    return argparse.Namespace(joint_train=joint_train, padding_margin=16, inpaint_path=None, pred_depth_minmax=True,
                              net1_path=None, load_offline=False)


def init_genre_net_for_bench(net, depth_range=(1.85, 2.55), depth_gain=10.0):
    """net: models.genre_full_model.Net (reference class, unmodified).  In-place; returns net."""
    net1 = net.depth_and_inpaint.net1
    with torch.no_grad():
        head = net1.decoder_minmax[-1]                       # nn.Linear(128, 2): (min, max) of the absolute depth (marrnet1.py:142-152)
        head.weight.mul_(0.05)
        head.bias.copy_(torch.tensor(depth_range, dtype=head.bias.dtype))
        last = net1.decoder_depth[-1][-1]                    # revresnet deconv2: the relative-depth map, later / scale_25d
        last.weight.mul_(depth_gain)
    return net


def genre_inputs(batch, device=None, seed=0, pin=False):
    """rgb [B,3,256,256] ~ N(0,1); silhou [B,1,256,256] = 100 * disc mask (scale_25d = 100, marrnetbase.py:17)."""
    g = torch.Generator().manual_seed(seed)
    rgb = torch.randn(batch, 3, 256, 256, generator=g)
    yy, xx = torch.meshgrid(torch.arange(256.0), torch.arange(256.0), indexing="ij")
    r2 = (yy - 127.5) ** 2 + (xx - 127.5) ** 2
    sil = torch.stack([(r2 < (70.0 + 3 * (i % 16)) ** 2).float() for i in range(batch)])[:, None] * 100
    if pin:
        rgb, sil = rgb.pin_memory(), sil.pin_memory()
    if device is not None:
        rgb, sil = rgb.to(device), sil.to(device)
    return types.SimpleNamespace(rgb=rgb, silhou=sil)

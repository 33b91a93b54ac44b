"""CPU stand-in of toolbox/spherical_proj.py:6-72 on the oracle (see ../README.md)."""
import numpy as np
import torch

from oracle import oracle
from ._pool import per_map


def gen_sph_grid(res=128):          # spherical_proj.py:6-18
    phi = np.linspace(0, 180, res * 2 + 1)[1::2] * np.pi / 180
    theta = np.linspace(0, 360, res + 1)[:-1] * np.pi / 180
    grid = np.zeros([res, res, 3])
    s = np.sin(phi)[:, None]
    grid[:, :, 0], grid[:, :, 1], grid[:, :, 2] = s * np.cos(theta)[None, :], s * np.sin(theta)[None, :], np.cos(phi)[:, None]
    return torch.from_numpy(grid.reshape(1, 1, res, res, 3)).float()


def sph_pad(sph_tensor, padding_margin=16):      # spherical_proj.py:21-28
    m = padding_margin
    out = torch.nn.functional.pad(sph_tensor, (m, m, m, m), mode='replicate')
    _, _, h, w = out.shape
    out[:, :, :, 0:m] = out[:, :, :, w - 2 * m:w - m]
    out[:, :, :, h - m:] = out[:, :, :, m:2 * m]
    return out


class render_spherical(torch.nn.Module):         # spherical_proj.py:31-72
    def __init__(self, sph_res=128, z_res=256):
        super().__init__()
        self.sph_res, self.z_res = sph_res, z_res
        # same arithmetic as spherical_proj.py:39-57: fp64 table * 2 * (1 - alpha), rounded to fp32 once
        phi = np.linspace(0, 180, sph_res * 2 + 1)[1::2] * np.pi / 180
        theta = np.linspace(0, 360, sph_res + 1)[:-1] * np.pi / 180
        d = np.zeros([sph_res, sph_res, 3])
        s = np.sin(phi)[:, None]
        d[:, :, 0], d[:, :, 1], d[:, :, 2] = s * np.cos(theta)[None, :], s * np.sin(theta)[None, :], np.cos(phi)[:, None]
        alpha = np.linspace(0, 1, z_res).reshape(1, 1, z_res, 1)
        grid = (d * 2)[:, :, None, :] * (1 - alpha)
        self.register_buffer('depth_weight', torch.linspace(0, 1, z_res))
        self.register_buffer('grid', torch.from_numpy(grid).float())

    def forward(self, vox):
        v = vox.detach().cpu().numpy()
        g, w = self.grid.numpy(), self.depth_weight.numpy()
        outs = per_map(lambda i: oracle.render_spherical(v[i:i + 1], g, w), v.shape[0])
        return torch.from_numpy(np.concatenate(outs, axis=0))

"""CPU stand-ins of CameraBackProjection / SphericalBackProjection (forward only) on the oracle."""
import numpy as np
import torch

from oracle import oracle
from ...._pool import per_map


class CameraBackProjection(torch.autograd.Function):      # cam_back_projection.py:12-30
    @staticmethod
    def forward(ctx, depth_t, fl, cam_dist, res=128):
        d = depth_t.detach().cpu().numpy()
        f, c = fl.detach().cpu().numpy(), cam_dist.detach().cpu().numpy()
        outs = per_map(lambda i: oracle.cam_bp_forward(d[i:i + 1], f[i:i + 1], c[i:i + 1], res)[0], d.shape[0])
        return torch.from_numpy(np.concatenate(outs, axis=0))


class SphericalBackProjection(torch.autograd.Function):   # sperical_to_tdf.py:13-31
    @staticmethod
    def forward(ctx, spherical, grid, res=128):
        s = spherical.detach().cpu().numpy()
        g = grid.detach().cpu().numpy()
        outs = per_map(lambda i: oracle.sph_bp_forward(s[i:i + 1], g[i:i + 1], res), s.shape[0])
        tdf = torch.from_numpy(np.concatenate([o[0] for o in outs], axis=0))
        cnt = torch.from_numpy(np.concatenate([o[1] for o in outs], axis=0))
        ctx.mark_non_differentiable(cnt)
        return tdf, cnt

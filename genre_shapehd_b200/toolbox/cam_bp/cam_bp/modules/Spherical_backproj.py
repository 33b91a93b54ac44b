"""spherical_backprojection — kept importable for API parity with
toolbox/cam_bp/cam_bp/modules/Spherical_backproj.py:7-17 (whose super() call names the wrong class
and therefore cannot be constructed in the reference)."""
import torch
from torch import nn

from ..functions import SphericalBackProjection


class spherical_backprojection(nn.Module):

    def __init__(self, grid, vox_res=128):
        super(spherical_backprojection, self).__init__()
        self.vox_res = vox_res
        self.register_buffer('grid', grid.float())

    def forward(self, spherical):
        grid = self.grid.expand(spherical.shape[0], -1, -1, -1, -1)
        return SphericalBackProjection.apply(spherical, grid, self.vox_res)

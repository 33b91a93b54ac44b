"""CPU stand-in of camera_backprojection_module.py:6-28."""
import torch
from torch.nn import Module

from ..functions import CameraBackProjection


class Camera_back_projection_layer(Module):
    def __init__(self, res=128):
        super().__init__()
        self.res = res

    def forward(self, depth_t, fl=418.3, cam_dist=2.2, shift=True):
        n = depth_t.size(0)
        if type(fl) == float:
            fl = torch.full((n, 1), fl, dtype=depth_t.dtype)
        if type(cam_dist) == float:
            cam_dist = torch.full((n, 1), cam_dist, dtype=depth_t.dtype)
        df = CameraBackProjection.apply(depth_t, fl, cam_dist, self.res)
        return self.shift_tdf(df) if shift else df

    @staticmethod
    def shift_tdf(input_tdf, res=128):
        return 1 - res * input_tdf

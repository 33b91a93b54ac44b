from .nnd import nndistance, nndistance_w_idx, nndistance_score

"""The reference's frozen model files import and build on top of this package's toolbox / nndistance / networks.
Runs only where a GenRe-ShapeHD checkout exists (the authoring container: /root/reference); the GPU box has none."""
import argparse
import os
import sys

import pytest
import torch

REF = os.environ.get("GENRE_REF", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="no reference checkout")


@pytest.fixture(scope="module")
def ref_root():
    from genre_shapehd_b200 import compat
    from conftest import REPO
    root = compat.bootstrap(REF)
    yield root, REPO


def _mine(obj, repo):
    mod = sys.modules[obj.__module__]
    return os.path.abspath(mod.__file__).startswith(os.path.join(repo, "genre_shapehd_b200"))


def test_frozen_models_import_on_top_of_this_package(ref_root):
    root, repo = ref_root
    import models.genre_full_model as gfm
    import models.shapehd as shd
    import models.wgangp as wg
    import models.marrnet2 as m2
    import models.depth_pred_with_sph_inpaint as dpi
    assert os.path.abspath(gfm.__file__).startswith(root)            # the caller is the reference's file, unchanged
    assert _mine(gfm.Camera_back_projection_layer, repo)
    assert _mine(gfm.SphericalBackProjection, repo)
    assert _mine(gfm.Unet_3D, repo)
    assert gfm.gen_sph_grid.__module__ == "toolbox.spherical_proj" and _mine(gfm.gen_sph_grid, repo)
    assert _mine(dpi.render_spherical, repo)
    assert _mine(wg.VoxelGenerator, repo) and _mine(wg.VoxelDiscriminator, repo)
    assert _mine(m2.VoxelDecoder, repo)
    assert shd is not None


def test_genre_net_builds_with_reference_constructor(ref_root):
    root, repo = ref_root
    import models.genre_full_model as gfm
    opt = argparse.Namespace(joint_train=False, padding_margin=16, inpaint_path=None, pred_depth_minmax=True,
                             net1_path=None, load_offline=False)
    net = gfm.Net(opt, gfm.Model)
    assert _mine(net.refine_net, repo) and _mine(net.proj_depth, repo)
    keys = list(net.state_dict().keys())
    assert "grid" in keys and any(k.startswith("refine_net.enc1.net.0.") for k in keys)
    assert any(k.startswith("depth_and_inpaint.render_spherical.") or "depth_weight" in k for k in keys)
    # the spherical back-projection glue of the frozen file runs on the new op when a GPU is present
    if torch.cuda.is_available():
        net = net.cuda()
        sph = torch.rand(2, 1, 160, 160, device="cuda") * 0.4 + 0.3
        out = net.backproject_spherical(sph)
        assert out.shape == (2, 1, 128, 128, 128)

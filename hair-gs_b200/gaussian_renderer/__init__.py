"""`render()` — the glue Hair-GS calls 9 times per code base (train.py:109,146; loss/losses.py:247,312;
render.py:60-80): same signature, same switches, same returned dict as gaussian_renderer/__init__.py:24-127 of
the reference, on top of the drop-in `diff_gaussian_rasterization` package of this repository.

`pc` is any object with the getter protocol of scene/gaussian_model.py:118-157 /
scene/hair_gaussian_model.py:134-206 (get_xyz, get_opacity, get_scaling, get_rotation, get_features,
get_covariance, active_sh_degree, max_sh_degree); hairgs_b200.models provides two such classes.
"""
import math

import torch

from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from hairgs_b200.sh import eval_sh


def render(viewpoint_camera, pc, bg_color, scaling_modifier=1.0, override_color=None, debug=False,
           compute_cov3D_python=False, convert_SHs_python=False):
    """Render the scene.  Background tensor (bg_color) must be on GPU!"""
    xyz = pc.get_xyz
    # zero tensor whose .grad receives the screen-space mean gradients (densification statistics)
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass

    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx, tanfovy=tanfovy, bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform, projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree, campos=viewpoint_camera.camera_center, prefiltered=False, debug=debug)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)

    means3D = xyz
    means2D = screenspace_points
    opacity = pc.get_opacity

    scales = rotations = cov3D_precomp = None
    if compute_cov3D_python:
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales = pc.get_scaling
        rotations = pc.get_rotation

    shs = colors_precomp = None
    if override_color is None:
        if convert_SHs_python:
            feats = pc.get_features
            shs_view = feats.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
            dir_pp = xyz - viewpoint_camera.camera_center.repeat(feats.shape[0], 1)
            dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
            sh2rgb = eval_sh(pc.active_sh_degree, shs_view, dir_pp_normalized)
            colors_precomp = torch.clamp_min(sh2rgb + 0.5, 0.0)
        else:
            shs = pc.get_features
    else:
        colors_precomp = override_color

    rendered_image, radii = rasterizer(means3D=means3D, means2D=means2D, shs=shs, colors_precomp=colors_precomp,
                                       opacities=opacity, scales=scales, rotations=rotations,
                                       cov3D_precomp=cov3D_precomp)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii}

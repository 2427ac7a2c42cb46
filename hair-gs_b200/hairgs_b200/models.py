"""Minimal parameter containers with the getter protocol `render()` relies on — the input builders of the hot
path, not the reference's training/topology machinery (densify/merge/grow/PLY are out of
scope).

  StrandModel : scene/hair_gaussian_model.py:134-206 — Gaussians derived from segment endpoints + width
  BlobModel   : scene/gaussian_model.py:118-157 — Stage-I free Gaussians (exp / normalize / sigmoid activations)
"""
import torch
import torch.nn as nn

from . import scenes
from .sh import build_covariance_from_scaling_rotation


class StrandModel(nn.Module):
    def __init__(self, scene: scenes.StrandScene, sh_degree=0):
        super().__init__()
        self.max_sh_degree = sh_degree
        self.active_sh_degree = sh_degree
        self._endpoints = nn.Parameter(scene.endpoints.clone())
        self.register_buffer("endpoint_pairs", scene.endpoint_pairs.clone())
        self._width = nn.Parameter(scene.width.clone())
        self._opacity = nn.Parameter(scene.opacity_logit.clone())
        self._mask = nn.Parameter(scene.mask_logit.clone())
        self._features_dc = nn.Parameter(scene.features_dc.clone())
        self._features_rest = nn.Parameter(scene.features_rest.clone())
        self._cache = None

    # every getter recomputes from the parameters on each access, as the reference's properties do
    @property
    def get_xyz(self):
        return scenes.strand_xyz(self._endpoints, self.endpoint_pairs)

    @property
    def get_scaling(self):
        return scenes.strand_scaling(self._endpoints, self.endpoint_pairs, self._width)

    @property
    def get_rotation(self):
        return scenes.strand_rotation(self._endpoints, self.endpoint_pairs)

    @property
    def get_orientation(self):
        return scenes.strand_orientation(self._endpoints, self.endpoint_pairs)

    @property
    def get_features(self):
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    @property
    def get_mask(self):
        return torch.sigmoid(self._mask)

    def get_covariance(self, scaling_modifier=1.0):
        return build_covariance_from_scaling_rotation(self.get_scaling, scaling_modifier, self.get_rotation)


class BlobModel(nn.Module):
    def __init__(self, scene: scenes.BlobScene, sh_degree=3):
        super().__init__()
        self.max_sh_degree = int(round(scene.shs.shape[1] ** 0.5)) - 1
        self.active_sh_degree = sh_degree
        self._xyz = nn.Parameter(scene.means3D.clone())
        self._scaling = nn.Parameter(torch.log(scene.scales))
        self._rotation = nn.Parameter(scene.rotations.clone())
        self._opacity = nn.Parameter(torch.logit(scene.opacities))
        self._features_dc = nn.Parameter(scene.shs[:, :1].clone())
        self._features_rest = nn.Parameter(scene.shs[:, 1:].clone())

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_scaling(self):
        return torch.exp(self._scaling)

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self._rotation)

    @property
    def get_features(self):
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    def get_covariance(self, scaling_modifier=1.0):
        return build_covariance_from_scaling_rotation(self.get_scaling, scaling_modifier, self._rotation)

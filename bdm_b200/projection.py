"""PC^2 projection conditioning on the B200 kernels.

reference: experiments/model/projection_model.py:127-157 (surface_projection) and :179-231
(get_input_with_conditioning).  The reference rasterises one sample at a time through pytorch3d
(PointsRasterizer, K=1, radius 0.0075, naive) inside a Python loop over a list of cameras; here a
batch of cameras is four small tensors and the whole batch is one call into libbdm_b200.so.

pytorch3d is not vendored in the reference (version unpinned): camera semantics are restated from
its published conventions (PerspectiveCameras in NDC: X_view = X R + T, ndc = f * xy / z + p; +X left,
+Y up) -- see oracle/bdm_oracle.c `orc_project_points` for the exact arithmetic.  PARITY UNPINNED.
"""
import math
from dataclasses import dataclass

import torch

from . import backend as _backend


@dataclass
class Cameras:
    """A batch of perspective cameras in NDC (what `PerspectiveCameras(R, T, focal_length,
    principal_point)` holds in the reference: dataset/shapenet_r2n2.py:86-93)."""
    R: torch.Tensor          # f32[B,3,3]  row-vector convention
    T: torch.Tensor          # f32[B,3]
    focal: torch.Tensor      # f32[B,2]
    principal: torch.Tensor  # f32[B,2]

    def to(self, device):
        return Cameras(self.R.to(device), self.T.to(device), self.focal.to(device), self.principal.to(device))

    def __len__(self):
        return self.R.shape[0]

    def slice(self, lo, hi):
        return Cameras(self.R[lo:hi].contiguous(), self.T[lo:hi].contiguous(), self.focal[lo:hi].contiguous(),
                       self.principal[lo:hi].contiguous())


def look_at_cameras(azimuth_deg, elevation_deg, distance, focal=2.1875):
    """R2N2-style cameras looking at the origin (dataset/shapenet_r2n2.py:46-53, :374-384:
    azimuth U[0,360), elevation U[25,30], distance U[0.65,0.95]*1.75, NDC focal 2.1875, pp 0),
    following pytorch3d's look_at_view_transform conventions (camera at C, +Z forward, +Y up, +X left)."""
    az = torch.as_tensor(azimuth_deg, dtype=torch.float64) * math.pi / 180.0
    el = torch.as_tensor(elevation_deg, dtype=torch.float64) * math.pi / 180.0
    d = torch.as_tensor(distance, dtype=torch.float64)
    C = torch.stack([d * torch.cos(el) * torch.sin(az), d * torch.sin(el), d * torch.cos(el) * torch.cos(az)], -1)
    z = torch.nn.functional.normalize(-C, dim=-1)
    up = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64).expand_as(z)
    x = torch.nn.functional.normalize(torch.cross(up, z, dim=-1), dim=-1)
    y = torch.nn.functional.normalize(torch.cross(z, x, dim=-1), dim=-1)
    R = torch.stack([x, y, z], dim=-1)                      # columns = camera axes -> X_view = X R + T
    T = -torch.einsum('bi,bij->bj', C, R)
    b = C.shape[0]
    return Cameras(R.float().contiguous(), T.float().contiguous(), torch.full((b, 2), float(focal)),
                   torch.zeros((b, 2)))


class ProjectionConditioner:
    """Step-invariant conditioning state + the per-step projection.

    `local_features` f32[B,C,H,W] is what the reference's get_local_conditioning returns
    (projection_model.py:110-125: normalised RGB + ViT feature map [+ mask channels]); the reference
    recomputes it -- including a ViT forward -- on every one of the 1000 steps although it does not
    depend on the step.  Here it is given once; a channel-last copy is kept so the per-point gather
    reads C contiguous floats."""

    def __init__(self, local_features, cameras, radius=0.0075, scale_factor=1.0, channel_last=True):
        self.cameras = Cameras(cameras.R.clone().contiguous(), (cameras.T * scale_factor).contiguous(),  # :136-137
                               cameras.focal.clone().contiguous(), cameras.principal.clone().contiguous())
        self.radius = float(radius)
        self.scale_factor = float(scale_factor)
        self.channel_last = channel_last
        self.C = local_features.shape[1]
        self.feat = local_features.permute(0, 2, 3, 1).contiguous() if channel_last else local_features.contiguous()

    def load(self, local_features, cameras):
        """Refill this conditioner IN PLACE with the feature maps and cameras of another batch of the same
        shape (device tensors, or pinned host tensors: the copies are stream-ordered).  CUDA graphs captured
        over this conditioner (BDMSampler.enable_cuda_graphs) keep pointing at valid, current data; the
        channel-last transposition happens inside the copy."""
        assert tuple(local_features.shape) == (self.feat.shape[0], self.C) + tuple(
            self.feat.shape[1:3] if self.channel_last else self.feat.shape[2:4]), "batch shape changed"
        if self.channel_last:
            if local_features.device != self.feat.device:      # host -> device first, then transpose on the device
                local_features = local_features.to(self.feat.device, non_blocking=True)
            self.feat.copy_(local_features.permute(0, 2, 3, 1))
        else:
            self.feat.copy_(local_features, non_blocking=True)
        cam = self.cameras
        cam.R.copy_(cameras.R, non_blocking=True)
        cam.T.copy_(cameras.T, non_blocking=True)
        if self.scale_factor != 1.0:
            cam.T.mul_(self.scale_factor)
        cam.focal.copy_(cameras.focal, non_blocking=True)
        cam.principal.copy_(cameras.principal, non_blocking=True)
        return self

    def surface_projection(self, points):
        """points f32[B,N,3] -> f32[B,N,C]: each visible point gets the feature vector of its pixel."""
        cam = self.cameras
        out, _ = _backend.surface_projection(points.contiguous(), cam.R, cam.T, cam.focal, cam.principal, self.feat,
                                             self.radius, feat_is_hwc=self.channel_last)
        return out

    def get_input_with_conditioning(self, x_t):
        """x_t f32[B,N,3] -> f32[B,N,3+C]   (projection_model.py:179-231 with local conditioning only)"""
        return torch.cat([x_t, self.surface_projection(x_t[:, :, :3])], dim=2)


def _channel_first_input(self, x_t):
    """x_t f32[B,N,3] -> f32[B,3+C,N], the denoiser's channel-first input, in one pass (no [B,N,C]
    intermediate, no concat, no transpose).  Same values as get_input_with_conditioning(x_t).transpose(1,2)."""
    cam = self.cameras
    out, _ = _backend.conditioning_input(x_t.contiguous(), cam.R, cam.T, cam.focal, cam.principal, self.feat, self.radius)
    return out


ProjectionConditioner.get_input_channel_first = _channel_first_input


def surface_projection(points, cameras, local_features, radius=0.0075, scale_factor=1.0):
    """Functional form with the reference's argument meaning (projection_model.py:127-157)."""
    return ProjectionConditioner(local_features, cameras, radius, scale_factor, channel_last=False) \
        .surface_projection(points)

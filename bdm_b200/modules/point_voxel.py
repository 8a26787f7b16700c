"""Voxelization and the point-voxel convolution block.

reference: modules/voxelization.py:9-28, modules/pvconv.py:65-137
"""
import os

import torch
import torch.nn as nn

from .. import functional as F
from ..functional import geometry
from ..functional import ops as _ops
from . import layers as _layers
from .layers import SE3d, Attention, FusedSequential, SharedMLP, Swish

# The first Conv3d of a PVConv block reads a freshly voxelized cloud: N points occupy at most N of the R^3
# voxels (~5 % at R=32, N=4096).  When N / R^3 <= SPARSE_MAX_FILL the block skips the dense grid and the
# dense convolution: per-occupied-voxel averages -> one GEMM against the 27 taps -> a gather that writes
# the convolution's dense output once (csrc/sparse_conv.cu).  Same sums as the dense convolution (the
# skipped terms are exact zeros); the GEMM runs in TF32 exactly when torch would run the Conv3d in TF32.
# The saving shrinks with the fill ratio (R=16 / N=1024 is break-even on its own), but this route is also
# what puts the block's voxel branch in channels-last memory, so it is on for every stage by default
# (measured on the PC^2 step: 6.44 ms with a 1/8 threshold, 6.21 ms without one).
# Inference on CUDA only; BDM_SPARSE_CONV=0 disables it.
SPARSE_FIRST_CONV = os.environ.get("BDM_SPARSE_CONV", "1") != "0"
SPARSE_MAX_FILL = float(os.environ.get("BDM_SPARSE_MAX_FILL", "1.0"))
DEFER_SE_GATE = os.environ.get("BDM_DEFER_SE_GATE", "1") != "0"
# Keep the voxel branch of a sparse-first-conv block in channels-last memory ([B,R,R,R,C]): the gather writes
# it, the fused norm kernels and the devoxelization read it, and cuDNN's Conv3d (whose tensor-core kernels are
# NDHWC inside) stops wrapping two transposes around every call.  Same values; BDM_CHANNELS_LAST=0 disables.
CHANNELS_LAST_VOXELS = os.environ.get("BDM_CHANNELS_LAST", "1") != "0"
# First convolution on the tcgen05 kernel too (csrc/conv3_tc05.cu) where the dense product is cheaper than the
# tap-product round trip of the sparse route: the occupied voxels' averages are scattered into the convolution's fp16
# operand (zeros elsewhere), and the convolution emits the first norm's statistics.  Measured at 32 shapes (B200):
# 64->64 at R=32 295 vs 363 us, 128->128 at R=16 ~125 vs 197 us, 32->32 at R=32 ~165 vs 196 us; wide inputs (390->32)
# and the 8^3 grids stay on the sparse route.  BDM_DENSE_FIRST_TC05=0 disables.
DENSE_FIRST_TC05 = os.environ.get("BDM_DENSE_FIRST_TC05", "1") != "0"
DENSE_FIRST_TC05_MAX_CIN = int(os.environ.get("BDM_DENSE_FIRST_TC05_MAX_CIN", "128"))
# ... and there the kernel skips the tap windows (258 flat rows x one (dx, dy)) that hold no occupied voxel: 31-44 % of
# them at R=32 along the sampling trajectory, 4-16 % at R=16.  Exact (the skipped products are zeros).  BDM_CONV3_SKIP_EMPTY=0.
SKIP_EMPTY_WINDOWS = os.environ.get("BDM_CONV3_SKIP_EMPTY", "1") != "0"


def normalized_voxel_coords(coords, resolution, normalize=True, eps=0):
    """Float coordinates in [0, R-1] for a cloud f32[B,3,N]: centre on the mean, scale by twice the
    largest point norm, shift to [0,1], stretch to the grid and clamp.  This is deliberately the
    reference's own sequence of torch ops (voxelization.py:17-23) -- mean, norm, max, divide, +0.5,
    *R, clamp -- because the rounded integer coordinates, and every voxel index after them, must be
    bit-identical."""
    centred = coords - coords.mean(2, keepdim=True)
    if normalize:
        extent = centred.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + eps
        unit = centred / extent + 0.5
    else:
        unit = (centred + 1) / 2.0
    return torch.clamp(unit * resolution, 0, resolution - 1)


# One kernel instead of the eight torch launches of normalized_voxel_coords + round (inference on CUDA, fp32).
# BDM_FUSED_VOXEL_COORDS=0 restores the torch sequence.
FUSED_VOXEL_COORDS = os.environ.get("BDM_FUSED_VOXEL_COORDS", "1") != "0"


def _coordinate_plan(coords, r, normalize, eps):
    """(float voxel coords, int voxel coords, voxel plan) for one (coords tensor, resolution)."""
    coords = coords.detach()
    if (FUSED_VOXEL_COORDS and coords.is_cuda and coords.dtype == torch.float32 and not _ops.REFERENCE_CALL_PATTERN
            and hasattr(_ops._B, "voxelize_coords")):
        norm_coords, vox_coords = _ops._B.voxelize_coords(coords.contiguous(), r, normalize, eps)
    else:
        norm_coords = normalized_voxel_coords(coords, r, normalize, eps)
        vox_coords = torch.round(norm_coords).to(torch.int32)  # half-to-even, like the reference
    return norm_coords, vox_coords, F.voxel_plan(vox_coords, r)


class _LastPlan:
    """The last (coords tensor, resolution) -> plan outside a geometry scope.  Consecutive PVConv blocks
    of one stage receive the SAME coords tensor object (PVConv.forward passes it through), so 2-3
    voxelizations per stage share one normalisation (~8 small torch launches) and one index/sort
    kernel.  Holding the coords tensor keeps its storage from being recycled; the version counter
    catches in-place edits.  Results are bit-identical to recomputing."""

    def __init__(self):
        self.key = self.coords = self.value = None

    def lookup(self, coords, r, normalize, eps):
        key = (coords.data_ptr(), geometry.tensor_version(coords), tuple(coords.shape), coords.device, r, normalize, eps)
        if self.coords is coords and self.key == key:
            return self.value
        self.key, self.coords = key, coords
        self.value = _coordinate_plan(coords, r, normalize, eps)
        return self.value


_last_plan = _LastPlan()


def coordinate_plan(coords, r, normalize=True, eps=0):
    if geometry.active() is not None:
        return geometry.memo("vox", (coords,), (int(r), bool(normalize), float(eps)),
                             lambda: _coordinate_plan(coords, r, normalize, eps))
    return _last_plan.lookup(coords, r, normalize, eps)


class Voxelization(nn.Module):
    def __init__(self, resolution, normalize=True, eps=0):
        super().__init__()
        self.r = int(resolution)
        self.normalize = normalize
        self.eps = eps

    def forward(self, features, coords):
        """-> (voxel grid f32[B,C,R,R,R], float voxel coordinates f32[B,3,N])"""
        norm_coords, vox_coords, plan = coordinate_plan(coords, self.r, self.normalize, self.eps)
        if plan is None:  # backend without the split entry points (e.g. the reference's own extension)
            return F.avg_voxelize(features, vox_coords, self.r), norm_coords
        return F.avg_voxelize_planned(features, plan), norm_coords

    def extra_repr(self):
        return 'resolution={}{}'.format(self.r, ', normalized eps = {}'.format(self.eps) if self.normalize else '')


def _voxel_stack(c_in, c_out, k, attention, dropout, with_se, with_se_relu, make_norm, make_act):
    """conv-norm-act-[dropout]-conv-norm-(attention|act)-[SE]; positions in the Sequential are the
    reference's (pvconv.py:75-88) so `voxel_layers.{i}` keys line up."""
    pad = k // 2
    seq = [nn.Conv3d(c_in, c_out, k, stride=1, padding=pad), make_norm(c_out), make_act()]
    if dropout is not None:
        seq.append(nn.Dropout(dropout))
    seq += [nn.Conv3d(c_out, c_out, k, stride=1, padding=pad), make_norm(c_out),
            Attention(c_out, 8) if attention else make_act()]
    if with_se:
        seq.append(SE3d(c_out, use_relu=with_se_relu))
    return FusedSequential(*seq)


POINT_BRANCH_STREAM = os.environ.get("BDM_POINT_STREAM", "1") != "0"
_POINT_STREAMS = {}


def _point_stream(device):
    dev = torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _POINT_STREAMS.get(key)
    if st is None:
        st = _POINT_STREAMS[key] = torch.cuda.Stream(device=dev)
    return st


class _PVConvBase(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, resolution, normalize, eps):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.resolution = resolution
        self.voxelization = Voxelization(resolution, normalize=normalize, eps=eps)

    def _sparse_eligible(self, features):
        conv = self.voxel_layers[0]
        vox = self.voxelization
        return (SPARSE_FIRST_CONV and features.is_cuda and features.dtype == torch.float32
                and not torch.is_grad_enabled() and not _ops.REFERENCE_CALL_PATTERN
                and hasattr(_ops._B, "sparse_conv3_gather")
                and _ops._B.sparse_conv3_supported(features.shape[2], vox.r)
                and features.shape[2] <= SPARSE_MAX_FILL * vox.r ** 3
                and isinstance(conv, nn.Conv3d) and conv.kernel_size == (3, 3, 3) and conv.stride == (1, 1, 1)
                and conv.padding == (1, 1, 1) and conv.dilation == (1, 1, 1) and conv.groups == 1
                and conv.padding_mode == 'zeros')

    def devoxelizes_channels_last(self, num_points):
        """True when an inference forward over `num_points` CUDA fp32 points keeps the voxel branch in
        channels-last memory and devoxelizes through trilinear_devoxelize_cl, which needs no x-slice plan
        (denoiser.plan_geometry_ahead then skips the binning kernel)."""
        conv = self.voxel_layers[0]
        vox = self.voxelization
        return (SPARSE_FIRST_CONV and CHANNELS_LAST_VOXELS and not _ops.REFERENCE_CALL_PATTERN
                and hasattr(_ops._B, "sparse_conv3_gather") and hasattr(_ops._B, "groupnorm_act_cl")
                and _ops._B.sparse_conv3_supported(num_points, vox.r) and num_points <= SPARSE_MAX_FILL * vox.r ** 3
                and _ops._B.groupnorm_cl_supported(self.out_channels, 8)
                and isinstance(conv, nn.Conv3d) and conv.kernel_size == (3, 3, 3) and conv.stride == (1, 1, 1)
                and conv.padding == (1, 1, 1) and conv.dilation == (1, 1, 1) and conv.groups == 1
                and conv.padding_mode == 'zeros')

    def _dense_first_eligible(self, features):
        """the sparse-eligible block's first convolution as a dense tcgen05 convolution (see DENSE_FIRST_TC05)"""
        conv, norm, vox = self.voxel_layers[0], self.voxel_layers[1], self.voxelization
        return (DENSE_FIRST_TC05 and CHANNELS_LAST_VOXELS and _layers.FUSED_NORM_ACT and _layers.CONV3_TC05
                and hasattr(_ops._B, "conv3_tc05_fill_planes") and bool(torch.backends.cudnn.allow_tf32)
                and isinstance(norm, nn.GroupNorm) and _ops._B.groupnorm_cl_supported(self.out_channels, 8)
                and vox.r >= _layers.CONV3_TC05_MIN_R and conv.in_channels <= DENSE_FIRST_TC05_MAX_CIN
                and conv.in_channels % 8 == 0 and (conv.in_channels * features.shape[2]) % 4 == 0
                and _ops._B.conv3_tc05_supported(conv.in_channels, conv.out_channels, vox.r))

    def _dense_first_conv(self, features, coords):
        """-> ((output of voxel_layers[0] INCLUDING its bias, channels-last; its GroupNorm statistics; True), coords)"""
        vox, conv = self.voxelization, self.voxel_layers[0]
        norm_coords, _, plan = coordinate_plan(coords, vox.r, vox.normalize, vox.eps)
        prepared = _layers.conv3_prepared(conv, None, 1)
        occupied = _ops._B.avg_voxelize_compact(features.contiguous(), plan, amax_into=prepared)     # [B, Cin, N] + max|.|
        planes = _layers.half_planes(features.shape[0], conv.in_channels, vox.r, features.device)
        _ops._B.conv3_tc05_fill_planes(occupied, plan, prepared, planes, amax_ready=True)
        out, stats = _ops._B.conv3_tc05(planes, prepared, conv.out_channels, bias=conv.bias, stats=_layers.CONV3_STATS,
                                        sparse=SKIP_EMPTY_WINDOWS)
        return (out.permute(0, 4, 1, 2, 3), stats, True), norm_coords

    def _tap_matrix(self, conv):
        """Conv3d weight [Cout,Cin,3,3,3] -> [Cin, 27*Cout] (column k*Cout+co), cached per weight version."""
        w = conv.weight
        key = (w.data_ptr(), geometry.tensor_version(w), w.device)
        cached = getattr(self, "_taps", None)
        if cached is None or cached[0] != key:
            cached = (key, w.detach().permute(1, 2, 3, 4, 0).reshape(w.shape[1], -1).contiguous())
            self._taps = cached
        return cached[1]

    def _sparse_first_conv(self, features, coords):
        """-> (bias-less output of voxel_layers[0] on the voxelized features, float voxel coordinates)"""
        vox = self.voxelization
        norm_coords, _, plan = coordinate_plan(coords, vox.r, vox.normalize, vox.eps)
        occupied = _ops._B.avg_voxelize_compact(features.contiguous(), plan)     # [B, Cin, N]
        with _layers.matmul_precision_of_convs():     # TF32 exactly when the Conv3d it replaces would use it
            taps = torch.matmul(occupied.transpose(1, 2), self._tap_matrix(self.voxel_layers[0]))  # [B, N, 27*Cout]
        if (CHANNELS_LAST_VOXELS and hasattr(_ops._B, "groupnorm_act_cl")
                and _ops._B.groupnorm_cl_supported(self.out_channels, 8)):
            norm = self.voxel_layers[1]
            second = next((m for m in list(self.voxel_layers)[2:] if isinstance(m, nn.Conv3d)), None)
            if (isinstance(norm, nn.GroupNorm) and _layers.FUSED_NORM_ACT
                    and ((self.out_channels // norm.num_groups) * vox.r ** 3 > 32768
                         or (second is not None and _layers.conv3_tc05_applicable(second, norm, self.out_channels, vox.r)))):
                # group too large for the one-pass norm, or the norm feeds the tcgen05 convolution (which wants the
                # statistics up front): the gather also emits the norm's statistics
                out, stats = _ops._B.sparse_conv3_gather(taps, plan, channels_last=True, stats=True)
                return (out.permute(0, 4, 1, 2, 3), stats), norm_coords
            return _ops._B.sparse_conv3_gather(taps, plan, channels_last=True).permute(0, 4, 1, 2, 3), norm_coords
        return _ops._B.sparse_conv3_gather(taps, plan), norm_coords

    def forward(self, inputs):
        features, coords, temb = inputs
        # Inference: an SE gate at the end of the voxel stack is one scalar per (shape, channel) and
        # devoxelization is linear in the grid, so the gate is applied to the devoxelized [B,C,N] features
        # instead of to the [B,C,R^3] grid (one pass over the grid less).
        defer = (DEFER_SE_GATE and features.is_cuda and not torch.is_grad_enabled()
                 and not _ops.REFERENCE_CALL_PATTERN and hasattr(_ops._B, "groupnorm_act"))
        first = first_stats = None
        first_biased = False
        # the point branch reads only `features`: issued on a second stream at the top of the block, its small GEMM and norm
        # kernels fill the SMs the voxel branch's narrow kernels (plans, SE gates, kernel tails) leave idle
        point_branch = side = None
        if defer and POINT_BRANCH_STREAM and features.dtype == torch.float32:
            side = _point_stream(features.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                point_branch = self.point_features(features).contiguous()
        if self._sparse_eligible(features):
            if self._dense_first_eligible(features):
                first, grid_coords = self._dense_first_conv(features, coords)
            else:
                first, grid_coords = self._sparse_first_conv(features, coords)
            if isinstance(first, tuple):
                first, first_stats, first_biased = (tuple(first) + (False,))[:3]
            grid = None
        else:
            grid, grid_coords = self.voxelization(features, coords)
        grid = self.voxel_layers(grid, first_output=first, defer_gate=defer, first_stats=first_stats,
                                 first_biased=first_biased, defer_norm=defer)
        gate = norm_coef = None
        if defer:
            grid, gate, norm_coef = grid
        if _layers.is_channels_last_3d(grid) and not torch.is_grad_enabled():
            # channels-last branch: (last norm + Swish,) devoxelize, SE gate and the residual add of the point branch in
            # one kernel
            fused = _ops._B.trilinear_devoxelize_cl(grid.permute(0, 2, 3, 4, 1), grid_coords.contiguous(), self.resolution,
                                                    gate=gate, residual=self._point_branch(features, point_branch, side),
                                                    norm_coef=norm_coef)
            return fused, coords, temb
        assert norm_coef is None
        from_voxels = F.trilinear_devoxelize(grid, grid_coords, self.resolution, self.training)
        if gate is not None:
            return torch.addcmul(self._point_branch(features, point_branch, side), from_voxels, gate[:, :, None]), coords, temb
        return from_voxels + self._point_branch(features, point_branch, side), coords, temb

    def _point_branch(self, features, ahead, side):
        """the point branch: computed here, or joined from the second stream forward() started it on"""
        if ahead is None:
            return self.point_features(features).contiguous()
        main = torch.cuda.current_stream()
        main.wait_stream(side)
        ahead.record_stream(main)
        return ahead


class PVConv(_PVConvBase):
    """voxelize -> 3-D convs (GroupNorm/Swish) -> trilinear devoxelize, fused with a point-wise MLP.
    Submodules: voxelization, voxel_layers, point_features."""

    def __init__(self, in_channels, out_channels, kernel_size, resolution, attention=False,
                 dropout=0.1, with_se=False, with_se_relu=False, normalize=True, eps=0):
        super().__init__(in_channels, out_channels, kernel_size, resolution, normalize, eps)
        self.voxel_layers = _voxel_stack(in_channels, out_channels, kernel_size, attention, dropout, with_se,
                                         with_se_relu, lambda c: nn.GroupNorm(num_groups=8, num_channels=c), Swish)
        self.point_features = SharedMLP(in_channels, out_channels)


class PVConvReLU(_PVConvBase):
    """BatchNorm / LeakyReLU variant of PVConv."""

    def __init__(self, in_channels, out_channels, kernel_size, resolution, attention=False, leak=0.2,
                 dropout=0.1, with_se=False, with_se_relu=False, normalize=True, eps=0):
        super().__init__(in_channels, out_channels, kernel_size, resolution, normalize, eps)
        self.voxel_layers = _voxel_stack(in_channels, out_channels, kernel_size, attention, dropout, with_se,
                                         with_se_relu, nn.BatchNorm3d, lambda: nn.LeakyReLU(leak, True))
        self.point_features = SharedMLP(in_channels, out_channels)

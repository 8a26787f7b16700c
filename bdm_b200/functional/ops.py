"""Autograd wrappers around the 12 native entry points.

Each class states which reference wrapper it mirrors (paths under
/root/reference/experiments/model/pvcnn/modules/functional/).  Forward semantics, dtype coercions
(`.int()`, `.contiguous()`), what is saved for backward and which inputs receive gradients are the
reference's; the implementation underneath is bdm_b200.backend.
"""
import torch
import torch.nn.functional as tnf
from torch.autograd import Function

from . import geometry
from .backend import _backend as _B


# ----------------------------------------------------------------------------------------------
# voxelize / devoxelize
# ----------------------------------------------------------------------------------------------
class _AvgVoxelize(Function):
    """voxelization.py:8-37.  (features f32[B,C,N], voxel coords int[B,3,N], R) -> f32[B,C,R,R,R]"""

    @staticmethod
    def forward(ctx, features, coords, resolution):
        feats = features.contiguous()
        grid, point_voxel, voxel_count = _B.avg_voxelize_forward(feats, coords.int().contiguous(), resolution)
        ctx.save_for_backward(point_voxel, voxel_count)
        return grid.view(feats.shape[0], feats.shape[1], resolution, resolution, resolution)

    @staticmethod
    def backward(ctx, grad_grid):
        point_voxel, voxel_count = ctx.saved_tensors
        flat = grad_grid.contiguous().view(grad_grid.shape[0], grad_grid.shape[1], -1)
        return _B.avg_voxelize_backward(flat, point_voxel, voxel_count), None, None


class _TrilinearDevoxelize(Function):
    """devoxelization.py:8-39.  (grid f32[B,C,R,R,R], coords f32[B,3,N] in [0,R-1], R, training)
    -> f32[B,C,N]; corner indices / weights are kept for backward only when training."""

    @staticmethod
    def forward(ctx, features, coords, resolution, is_training=True):
        nb, nc = features.shape[:2]
        flat = features.contiguous().view(nb, nc, -1)
        pts = coords.contiguous()
        plan = None if is_training else devoxelize_plan(pts, resolution)
        if plan is not None:
            outs, corner_idx, corner_w = _B.trilinear_devoxelize_forward(resolution, is_training, pts, flat, plan)
        else:
            outs, corner_idx, corner_w = _B.trilinear_devoxelize_forward(resolution, is_training, pts, flat)
        if is_training:
            ctx.save_for_backward(corner_idx, corner_w)
            ctx.r = resolution
        return outs

    @staticmethod
    def backward(ctx, grad_out):
        corner_idx, corner_w = ctx.saved_tensors
        r = ctx.r
        g = _B.trilinear_devoxelize_backward(grad_out.contiguous(), corner_idx, corner_w, r)
        return g.view(grad_out.size(0), grad_out.size(1), r, r, r), None, None, None


class _AvgVoxelizePlanned(Function):
    """avg_voxelize for callers that already hold the coordinate-only half of the op (`voxel_plan`)."""

    @staticmethod
    def forward(ctx, features, plan):
        feats = features.contiguous()
        grid = _B.avg_voxelize_fill(feats, plan)
        ctx.save_for_backward(plan.ind, plan.cnt)
        r = plan.r
        return grid.view(feats.shape[0], feats.shape[1], r, r, r)

    @staticmethod
    def backward(ctx, grad_grid):
        point_voxel, voxel_count = ctx.saved_tensors
        flat = grad_grid.contiguous().view(grad_grid.shape[0], grad_grid.shape[1], -1)
        return _B.avg_voxelize_backward(flat, point_voxel, voxel_count), None


avg_voxelize = _AvgVoxelize.apply
avg_voxelize_planned = _AvgVoxelizePlanned.apply
trilinear_devoxelize = _TrilinearDevoxelize.apply


def devoxelize_plan(coords, resolution):
    """Coordinate-only half of inference devoxelization, shared by every grid devoxelized over the same
    coords tensor (memoised by tensor identity: geometry scope, else the last call).  None when the
    active backend has no such entry point."""
    if REFERENCE_CALL_PATTERN or not hasattr(_B, "devoxelize_plan"):
        return None
    return geometry.memo("devox", (coords,), (int(resolution),), lambda: _B.devoxelize_plan(coords, resolution),
                         keep_last=True)


def voxel_plan(vox_coords, resolution):
    """Coordinate-only half of avg_voxelize (voxel index, counts, sorted lookup tables), reusable for
    every feature tensor voxelized over the same coordinates.  None when the active backend has no
    such entry point (the reference's extension does not)."""
    if REFERENCE_CALL_PATTERN or not hasattr(_B, "voxel_plan"):
        return None
    return _B.voxel_plan(vox_coords.int().contiguous(), resolution)


# ----------------------------------------------------------------------------------------------
# index gathers
# ----------------------------------------------------------------------------------------------
class _Group(Function):
    """grouping.py:8-28.  (features f32[B,C,N], neighbour indices int32[B,M,U]) -> f32[B,C,M,U]"""

    @staticmethod
    def forward(ctx, features, indices):
        feats, idx = features.contiguous(), indices.contiguous()
        ctx.save_for_backward(idx)
        ctx.num_points = feats.size(-1)
        return _B.grouping_forward(feats, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _B.grouping_backward(grad_out.contiguous(), idx, ctx.num_points), None


class _Gather(Function):
    """sampling.py:10-31.  (features f32[B,C,N], centre indices int[B,M]) -> f32[B,C,M]"""

    @staticmethod
    def forward(ctx, features, indices):
        feats, idx = features.contiguous(), indices.int().contiguous()
        ctx.save_for_backward(idx)
        ctx.num_points = feats.size(-1)
        return _B.gather_features_forward(feats, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _B.gather_features_backward(grad_out.contiguous(), idx, ctx.num_points), None


grouping = _Group.apply
gather = _Gather.apply


# When True the modules issue exactly the reference's native-call sequence (12 groupings and 8 three-NN
# calls per denoiser forward) instead of skipping the redundant ones.  Used only to time the
# reference's own kernels under the reference's own call pattern (bench.py `reference_cuda`).
REFERENCE_CALL_PATTERN = False


def group_time_embedding(temb, neighbor_indices):
    """`grouping(temb, idx)` as called at modules/ball_query.py:30.  The denoisers pass a time
    embedding that is an `expand` of one vector per shape (pvcnn.py:88): stride 0 along the point
    axis, so every gathered element is the same value and the grouped tensor is that vector
    broadcast to [B,C,M,U] -- bit-identical to the gather without moving 4*C*M*U bytes per shape.
    Any other layout takes the real gather."""
    if not REFERENCE_CALL_PATTERN and temb.dim() == 3 and temb.size(-1) > 0 and temb.stride(-1) == 0:
        m, u = neighbor_indices.shape[1], neighbor_indices.shape[2]
        return temb[:, :, :1].unsqueeze(-1).expand(-1, -1, m, u)
    return grouping(temb, neighbor_indices)


def furthest_point_sample(coords, num_samples):
    """sampling.py:37-48.  coords f32[B,3,N] -> coordinates of the M sampled centres f32[B,3,M]"""
    pts = coords.contiguous()
    return geometry.memo("fps", (pts,), (int(num_samples),),
                         lambda: gather(pts, _B.furthest_point_sampling(pts, num_samples)))


def ball_query(centers_coords, points_coords, radius, num_neighbors):
    """ball_query.py:8-19.  (centres f32[B,3,M], points f32[B,3,N]) -> int32[B,M,U]"""
    cen, pts = centers_coords.contiguous(), points_coords.contiguous()
    return geometry.memo("ball", (cen, pts), (float(radius), int(num_neighbors)),
                         lambda: _B.ball_query(cen, pts, radius, num_neighbors))


# ----------------------------------------------------------------------------------------------
# three-nearest-neighbour interpolation
# ----------------------------------------------------------------------------------------------
class _NeighborInterpolate(Function):
    """interpolatation.py:8-35 (sic).  (points f32[B,3,N], centres f32[B,3,M], feats f32[B,C,M])
    -> f32[B,C,N]; only the features receive a gradient."""

    @staticmethod
    def forward(ctx, points_coords, centers_coords, centers_features):
        cen = centers_coords.contiguous()
        out, idx, w = _B.three_nearest_neighbors_interpolate_forward(points_coords.contiguous(), cen,
                                                                    centers_features.contiguous())
        ctx.save_for_backward(idx, w)
        ctx.num_centers = cen.size(-1)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, w = ctx.saved_tensors
        g = _B.three_nearest_neighbors_interpolate_backward(grad_out.contiguous(), idx, w, ctx.num_centers)
        return None, None, g


class _InterpolateWith(Function):
    """Interpolation half alone, for callers that hold (indices, weights) from `three_nn_search`."""

    @staticmethod
    def forward(ctx, centers_features, indices, weights):
        feats = centers_features.contiguous()
        ctx.save_for_backward(indices, weights)
        ctx.num_centers = feats.size(-1)
        return _B.three_nn_interpolate(feats, indices, weights)

    @staticmethod
    def backward(ctx, grad_out):
        idx, w = ctx.saved_tensors
        return _B.three_nearest_neighbors_interpolate_backward(grad_out.contiguous(), idx, w, ctx.num_centers), None, None


nearest_neighbor_interpolate = _NeighborInterpolate.apply
three_nn_interpolate = _InterpolateWith.apply


def three_nn_search(points_coords, centers_coords):
    """Search half alone -> (indices int32[B,3,N], weights f32[B,3,N]).  No gradient flows through
    the coordinates in the reference either (its backward returns None for both)."""
    pts, cen = points_coords.detach().contiguous(), centers_coords.detach().contiguous()
    return geometry.memo("nn3", (pts, cen), (), lambda: _B.three_nn_search(pts, cen))


# ----------------------------------------------------------------------------------------------
# upstream-PVCNN leftovers kept for import compatibility (plain torch; loss.py:7-17, sampling.py:51-88)
# ----------------------------------------------------------------------------------------------
def kl_loss(x, y):
    p = tnf.softmax(x.detach(), dim=1)
    return (p * (p.log() - tnf.log_softmax(y, dim=1))).sum(dim=1).mean()


def huber_loss(error, delta):
    a = error.abs()
    q = torch.clamp(a, max=delta)
    return (0.5 * q * q + delta * (a - q)).mean()


def logits_mask(coords, logits, num_points_per_object):
    """Frustum-PointNet leftover of upstream PVCNN (reference functional/sampling.py:51-88).  No BDM driver
    reaches it: the name stays importable, the routine is out of scope."""
    raise NotImplementedError("logits_mask is not part of the BDM hot path (reference functional/sampling.py:51-88)")

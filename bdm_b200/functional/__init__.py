"""Drop-in mirror of the reference's `modules.functional` package
(/root/reference/experiments/model/pvcnn/modules/functional/__init__.py:1-7 and the PVD copy):
the same ten public names with the same call signatures and autograd behaviour, served by
libbdm_b200.so through `backend._backend`.  Three extra names (`three_nn_search`,
`three_nn_interpolate`, `group_time_embedding`) let the modules skip work the reference repeats."""
from .backend import _backend
from .ops import (avg_voxelize, avg_voxelize_planned, voxel_plan, devoxelize_plan, ball_query, furthest_point_sample, gather, group_time_embedding, grouping,
                  huber_loss, kl_loss, logits_mask, nearest_neighbor_interpolate, three_nn_interpolate,
                  three_nn_search, trilinear_devoxelize)

__all__ = ['ball_query', 'trilinear_devoxelize', 'grouping', 'nearest_neighbor_interpolate', 'kl_loss',
           'huber_loss', 'gather', 'furthest_point_sample', 'logits_mask', 'avg_voxelize',
           'three_nn_search', 'three_nn_interpolate', 'group_time_embedding', 'avg_voxelize_planned',
           'voxel_plan', 'devoxelize_plan', '_backend']

"""A `_backend`-shaped object over the CPU oracle, for CPU-only tests of the host-side mirror
(modules / denoiser / drop-in).  TEST INFRASTRUCTURE: takes torch CPU tensors, calls oracle/."""
import numpy as np
import torch

import oracle as O


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _n(t):
    return t.detach().cpu().numpy()


class OracleBackend:
    calls = None

    def __init__(self):
        self.calls = []

    def _log(self, name, *shapes):
        self.calls.append((name,) + tuple(tuple(s) for s in shapes))

    def avg_voxelize_forward(self, features, coords, resolution):
        self._log("avg_voxelize_forward", features.shape)
        return [_t(x) for x in O.avg_voxelize_forward(_n(features), _n(coords), int(resolution))]

    def avg_voxelize_backward(self, grad_y, indices, cnt):
        return _t(O.avg_voxelize_backward(_n(grad_y), _n(indices), _n(cnt)))

    def trilinear_devoxelize_forward(self, r, is_training, coords, features):
        self._log("trilinear_devoxelize_forward", features.shape)
        return [_t(x) for x in O.trilinear_devoxelize_forward(int(r), bool(is_training), _n(coords), _n(features))]

    def trilinear_devoxelize_backward(self, grad_y, indices, weights, r):
        return _t(O.trilinear_devoxelize_backward(_n(grad_y), _n(indices), _n(weights), int(r)))

    def furthest_point_sampling(self, coords, num_samples):
        self._log("furthest_point_sampling", coords.shape)
        return _t(O.furthest_point_sampling(_n(coords), int(num_samples)))

    def gather_features_forward(self, features, indices):
        self._log("gather_features_forward", features.shape)
        return _t(O.gather_features_forward(_n(features), _n(indices)))

    def gather_features_backward(self, grad_y, indices, n):
        return _t(O.gather_features_backward(_n(grad_y), _n(indices), int(n)))

    def ball_query(self, centers, points, radius, num_neighbors):
        self._log("ball_query", centers.shape, points.shape)
        return _t(O.ball_query(_n(centers), _n(points), float(radius), int(num_neighbors)))

    def grouping_forward(self, features, indices):
        self._log("grouping_forward", features.shape, indices.shape)
        return _t(O.grouping_forward(_n(features), _n(indices)))

    def grouping_backward(self, grad_y, indices, n):
        return _t(O.grouping_backward(_n(grad_y), _n(indices), int(n)))

    def three_nn_search(self, points, centers):
        self._log("three_nn_search", points.shape, centers.shape)
        idx, w = O.three_nn(_n(points), _n(centers))
        return _t(idx), _t(w)

    def three_nn_interpolate(self, features, indices, weights):
        self._log("three_nn_interpolate", features.shape)
        return _t(O.three_interpolate(_n(features), _n(indices), _n(weights)))

    def three_nearest_neighbors_interpolate_forward(self, points, centers, features):
        self._log("three_nearest_neighbors_interpolate_forward", features.shape)
        return [_t(x) for x in O.three_nearest_neighbors_interpolate_forward(_n(points), _n(centers), _n(features))]

    def three_nearest_neighbors_interpolate_backward(self, grad_y, indices, weights, m):
        return _t(O.three_nearest_neighbors_interpolate_backward(_n(grad_y), _n(indices), _n(weights), int(m)))

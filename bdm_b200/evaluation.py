"""Chamfer distance and F-score on the B200 nearest-neighbour kernel.

reference: experiments/evaluation/evaluation_cd.py:111-132 (pytorch3d chamfer_distance on mean-centred
fp64 clouds, x1000) and experiments/evaluation/evaluation_f1.py:90-110 (dense squared-distance matrix in
the expansion form, clamp 1e-12, threshold 0.01 on the SQUARED distance).  The reference evaluates one
pair per call and moves the distances to Python lists; here a batch of pairs is two kernel launches
per metric and the reductions stay on the device."""
import torch

from . import backend as _backend


def _as_batch(x):
    x = x if x.dim() == 3 else x.unsqueeze(0)
    return x.to(torch.float64).contiguous()


def center(x):
    """evaluation_cd.py:115,123: subtract the per-cloud mean"""
    return x - x.mean(dim=1, keepdim=True)


def chamfer_distance(pred, gt):
    """pred f64[B,N,3], gt f64[B,M,3] -> per-pair CD f64[B] = mean_i min_j |p_i-g_j|^2 + mean_j min_i |.|^2
    (pytorch3d defaults: squared L2, point_reduction='mean'; the caller averages over pairs)."""
    pred, gt = _as_batch(pred), _as_batch(gt)
    s_pg, _ = _backend.nn_f64_reduce(pred, gt, expanded=False)
    s_gp, _ = _backend.nn_f64_reduce(gt, pred, expanded=False)
    return s_pg / pred.shape[1] + s_gp / gt.shape[1]


def compute_pc_to_pc_dist(src, tgt):
    """evaluation_f1.py:90-98, batched: min over tgt of the clamped expansion-form squared distance."""
    d, _ = _backend.nn_f64(_as_batch(src), _as_batch(tgt), expanded=True, return_index=False)
    return d


def fscore(gt, pred, thr=0.01):
    """evaluation_f1.py:101-110 -> per-pair F f64[B]"""
    gt, pred = _as_batch(gt), _as_batch(pred)
    _, c1 = _backend.nn_f64_reduce(gt, pred, expanded=True, thr=thr)      # d1 = compute_pc_to_pc_dist(gt, pred)
    _, c2 = _backend.nn_f64_reduce(pred, gt, expanded=True, thr=thr)      # d2 = compute_pc_to_pc_dist(pred, gt)
    precision = c1.double() / gt.shape[1]
    recall = c2.double() / pred.shape[1]
    return 2 * recall * precision / (recall + precision + 1e-12)


def evaluate(pred, gt):
    """Mean-centre both clouds like the reference scripts, return (CD*1000 [B], F-score@0.01 [B])."""
    pred, gt = center(_as_batch(pred)), center(_as_batch(gt))
    return chamfer_distance(pred, gt) * 1000.0, fscore(gt, pred)

"""Drop-in for the reference's ``model/point_utils.py`` (lines cited per function).  Index
construction (FPS, ball query, sorted k-NN, 3-NN) runs in single CUDA launches; the gathers and
the interpolation weights are ordinary differentiable tensor ops."""
from __future__ import annotations

import torch

from . import ops


_FPS_START_FEED = None  # set by step.GraphedTrainStep: start indices come from a static device buffer


def set_fps_start_feed(feed):
    """``feed()`` must return a device int32 tensor [B] of FPS start indices (drawn by the caller
    with ``torch.randint`` exactly like the reference).  None restores the in-place draw."""
    global _FPS_START_FEED
    _FPS_START_FEED = feed


def farthest_point_sample(xyz, npoint):
    """point_utils.py:5-26.  xyz [B,3,N] -> int64 [B,npoint].  The start index is drawn with
    ``torch.randint`` on the CPU generator exactly like line 17, so seeded runs consume the RNG
    identically to the reference."""
    B, _, N = xyz.shape
    if _FPS_START_FEED is not None:
        start = _FPS_START_FEED()
    else:
        start = torch.randint(0, N, (B,), dtype=torch.long)
    return ops.fps(xyz, npoint, start).long()


def index_points(points, idx):
    """point_utils.py:60-83.  points [B,C,N] (or [B,C,N,1]), idx [B,S] / [B,S,K] ->
    [B,C,S] / [B,C,S,K]."""
    if points.dim() == 4:
        points = points.squeeze(3)
    B, C, _ = points.shape
    flat = idx.reshape(B, 1, -1).long().expand(B, C, -1)
    return torch.gather(points, 2, flat).reshape(B, C, *idx.shape[1:])


def square_distance(src, dst):
    """point_utils.py:112-131.  [B,C,N],[B,C,M] -> [B,N,M]."""
    B, _, N = src.shape
    M = dst.shape[2]
    dist = -2 * torch.matmul(src.permute(0, 2, 1), dst)
    dist = dist + torch.sum(src ** 2, 1).view(B, N, 1)
    dist = dist + torch.sum(dst ** 2, 1).view(B, 1, M)
    return dist


def query_ball_point(radius, nsample, xyz, new_xyz):
    """point_utils.py:86-109.  radius given: lowest-index ``nsample`` points within the radius,
    padded with the first hit.  radius None: the ``nsample`` nearest points, ascending."""
    if radius is not None:
        return ops.ball_query(xyz, new_xyz, radius, nsample).long()
    return ops.knn_query(xyz, new_xyz, nsample).long()


def upsample_inter(xyz1, xyz2, points1, points2, k):
    """point_utils.py:134-165: inverse-squared-distance interpolation of points2 (at xyz2) onto
    xyz1 from the k nearest nodes; gradients flow to xyz2 and points2 as in the reference."""
    if points1 is not None and points1.dim() == 4:
        points1 = points1.squeeze(3)
    if points2.dim() == 4:
        points2 = points2.squeeze(3)
    B, _, N = xyz1.shape
    idx = ops.three_nn(xyz1, xyz2, k).long()  # [B,N,k]
    nb = index_points(xyz2, idx)  # [B,3,N,k]
    dots = (xyz1.unsqueeze(3) * nb).sum(1)
    dists = -2 * dots + torch.sum(xyz1 ** 2, 1).view(B, N, 1) + torch.sum(nb ** 2, 1)
    dists = torch.where(dists < 1e-10, torch.full_like(dists, 1e-10), dists)
    weight = 1.0 / dists
    weight = weight / torch.sum(weight, dim=-1).view(B, N, 1)
    interpolated = torch.sum(index_points(points2, idx) * weight.view(B, 1, N, k), dim=3)
    if points1 is not None:
        return torch.cat([points1, interpolated], dim=1)
    return interpolated

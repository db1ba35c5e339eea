"""Query-point extraction from masks -- drop-in for the reference's `sam_pt/utils/query_points.py`.

Same functions, same argument meaning, same use of the GLOBAL torch RNG (`torch.randperm` on the host), so that with the same
`torch.manual_seed` the same pixels are drawn as in the reference:

* `extract_random_mask_points`   (reference :29-61)
* `extract_kmedoid_points`       (:64-104)  k-medoids of scikit-learn-extra re-built on the GPU: the float32 distance matrix, its
                                            numpy-order row sums, and the whole alternate loop run in libsampt_b200
                                            (csrc/query_points.cu); the heuristic initialisation's `np.argpartition` is numpy's
                                            own call on the n row sums (its tie order is implementation-defined)
* `extract_corner_points`        (:107-163) Shi-Tomasi corners: OpenCV on the host exactly as the reference (one frame + one mask
                                            per call, pre-processing that runs once per clip, not per frame)
* `erode_mask_proportional_to_its_furthest_points_distance` (:165-194)
* `extract_mixed_points`         (:197-237)

Masks may live on the GPU (the VOS path keeps them there); results are float32 (n, 2) = (x, y) on the mask's device.
"""
from __future__ import annotations

from ctypes import c_int
from typing import List

import numpy as np
import torch

from sampt_b200 import native


def _mask_pixels(mask: torch.Tensor) -> torch.Tensor:
    return mask.nonzero().float()            # (n, 2) = (y, x), row-major order on CPU and CUDA alike


def extract_random_mask_points(mask: torch.Tensor, n_points_to_select: int) -> torch.Tensor:
    if mask.sum() == 0:
        print("Warning: mask.sum() == 0 in extract_random_mask_points")
        return torch.zeros((n_points_to_select, 2), device=mask.device)
    px = _mask_pixels(mask)
    if len(px) < n_points_to_select:
        sel = px.repeat(n_points_to_select // len(px) + 1, 1)[:n_points_to_select]
    else:
        sel = px[torch.randperm(len(px))[:n_points_to_select].to(px.device)]
    sel = sel.flip(1)
    assert sel.shape == (n_points_to_select, 2)
    return sel


def kmedoids_gpu(points_yx: torch.Tensor, n_clusters: int, max_iter: int = 300, info: dict | None = None) -> torch.Tensor:
    """`KMedoids(n_clusters).fit(points).cluster_centers_` for float32 points (n, 2) on a CUDA device, n <= 2048."""
    if not points_yx.is_cuda:
        raise RuntimeError("kmedoids_gpu runs in libsampt_b200 on a CUDA device; there is no CPU fallback")
    pts = points_yx.contiguous().float()
    n = pts.shape[0]
    if n_clusters > n:
        raise ValueError("The number of medoids must be less than the number of samples")
    dev = pts.device
    ctx = native.get_context(dev)
    L = native.lib()
    D = torch.empty((n, n), device=dev, dtype=torch.float32)
    rowsum = torch.empty((n,), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        native.check(L.sampt_kmedoids_distances(ctx.handle, native.ptr(pts), c_int(n), native.ptr(D), native.ptr(rowsum),
                                                native.stream_ptr()), "kmedoids_distances")
        # "heuristic" init of the package: the k points with the smallest distance sums, in numpy's argpartition order
        med0 = np.argpartition(rowsum.cpu().numpy(), n_clusters - 1)[:n_clusters]
        med = torch.as_tensor(med0.astype(np.int32), device=dev)
        scratch_i = torch.empty((2 * n,), device=dev, dtype=torch.int32)
        scratch_f = torch.empty((n,), device=dev, dtype=torch.float32)
        n_iter = torch.zeros((1,), device=dev, dtype=torch.int32)
        native.check(L.sampt_kmedoids_iterate(ctx.handle, native.ptr(D), c_int(n), c_int(n_clusters), c_int(max_iter), native.ptr(med),
                                              native.ptr(scratch_i), native.ptr(scratch_f), native.ptr(n_iter), native.stream_ptr()),
                     "kmedoids_iterate")
    if info is not None:
        info["n_iter"] = int(n_iter.item())
        info["medoid_idxs"] = med.cpu().numpy()
    return pts[med.long()]


def extract_kmedoid_points(mask: torch.Tensor, n_points_to_select: int, subsample_size: int = 1800, info: dict | None = None):
    if mask.sum() == 0:
        print("Warning: mask.sum() == 0 in extract_kmedoid_points")
        return torch.zeros((n_points_to_select, 2), device=mask.device)
    px = _mask_pixels(mask)
    if len(px) < n_points_to_select:
        sel = px.repeat(n_points_to_select // len(px) + 1, 1)[:n_points_to_select]
    else:
        px = px[torch.randperm(len(px))[:subsample_size].to(px.device)]
        sel = kmedoids_gpu(px, n_points_to_select, info=info)
    sel = sel.flip(1)                      # (y, x) -> (x, y)
    assert sel.shape == (n_points_to_select, 2)
    return sel


def erode_mask_proportional_to_its_furthest_points_distance(mask: torch.Tensor, erosion_percentage: float) -> torch.Tensor:
    import cv2
    px = _mask_pixels(mask)
    diameter = torch.norm(px.max(0)[0] - px.min(0)[0]).item()
    size = int(diameter * erosion_percentage)
    er = cv2.erode(mask.cpu().numpy().astype(np.uint8), np.ones((size, size), np.uint8), iterations=1)
    return torch.from_numpy(er).type(mask.dtype).to(mask.device)


def extract_corner_points(image: torch.Tensor, mask: torch.Tensor, n_points_to_select: int, kmedoid_subsample_size: int = 2000):
    import cv2
    if mask.sum() == 0:
        print("Warning: mask.sum() == 0 in extract_corner_points")
        return torch.zeros((n_points_to_select, 2), device=mask.device)
    img = image.permute(1, 2, 0).cpu().numpy()
    er = erode_mask_proportional_to_its_furthest_points_distance(mask, erosion_percentage=0.06)
    if er.sum() < 10:
        er = erode_mask_proportional_to_its_furthest_points_distance(mask, erosion_percentage=0.02)
    if er.sum() < 10:
        er = erode_mask_proportional_to_its_furthest_points_distance(mask, erosion_percentage=0.01)
    if er.sum() < 10:
        er = mask
    px = _mask_pixels(er)
    diameter = torch.norm(px.max(0)[0] - px.min(0)[0]).item()
    gray = cv2.cvtColor(np.ascontiguousarray(img), cv2.COLOR_RGB2GRAY)
    pts = cv2.goodFeaturesToTrack(image=gray, maxCorners=n_points_to_select, qualityLevel=0.001,
                                  minDistance=diameter / n_points_to_select, mask=er.cpu().numpy().astype(np.uint8), blockSize=3,
                                  gradientSize=3)
    if pts is None:
        pts = np.empty((0, 1, 2))
    pts = torch.from_numpy(pts).type(torch.float32).squeeze(1).to(mask.device)
    if len(pts) < n_points_to_select:
        fill = extract_kmedoid_points(mask, n_points_to_select - pts.shape[0], subsample_size=kmedoid_subsample_size)
        pts = torch.cat((pts, fill), dim=0)
    assert pts.shape == (n_points_to_select, 2)
    return pts


def extract_mixed_points(query_masks: List[torch.Tensor], query_points_timestep: torch.Tensor, images: torch.Tensor,
                         n_points: int) -> List[torch.Tensor]:
    n_kmedoid, n_shi_tomasi = n_points // 4, n_points // 3
    n_random = n_points - n_kmedoid - n_shi_tomasi
    lists = []
    if n_kmedoid > 0:
        lists += [[extract_kmedoid_points(qm, n_kmedoid) for qm in query_masks]]
    if n_shi_tomasi > 0:
        lists += [[extract_corner_points(images[int(t.item()), :, :, :], qm, n_shi_tomasi)
                   for qm, t in zip(query_masks, query_points_timestep)]]
    if n_random > 0:
        lists += [[extract_random_mask_points(qm, n_random) for qm in query_masks]]
    if len(lists) == 1:
        return lists[0]
    return [torch.cat(x, dim=0) for x in zip(*lists)]

"""ORACLE (test infrastructure, NOT product code): CPU restatement of the reference's query-point extraction from masks.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU leg may import this.

Restates /root/reference/sam_pt/utils/query_points.py:
* ``extract_random_mask_points``  :29-61
* ``extract_kmedoid_points``      :64-104  (random subsample of <= 1800 mask pixels with the global torch RNG, then
                                            ``sklearn_extra.cluster.KMedoids(n_clusters=N).fit(pixels).cluster_centers_``)
* ``extract_corner_points``       :107-163 (erosion + cv2.goodFeaturesToTrack + k-medoid fill-up)
* ``erode_mask_proportional_to_its_furthest_points_distance`` :165-194
* ``extract_mixed_points``        :197-237
and ``SamPt.extract_query_points`` / ``_extract_query_points_xy`` (sam_pt/modeling/sam_pt.py:238-306).

``sklearn_extra`` (scikit-learn-extra, reference requirements.txt) is NOT installed in this image and is not vendored in the
reference, so ``KMedoids`` is restated here from its published algorithm (scikit-learn-extra 0.3.0, `_k_medoids.py`), with
the constructor defaults the reference relies on: metric="euclidean", method="alternate", init="heuristic", max_iter=300:

    D = pairwise_distances(X)                                   (sklearn: float32 in -> float32 out, sqrt of the exact
                                                                 integer squared distance for pixel coordinates)
    medoids = argpartition(D.sum(axis=1), k-1)[:k]              ("heuristic" init: the k points with the smallest row sums)
    repeat <= 300 times:
        labels = argmin(D[medoids, :], axis=0)
        for each cluster c (in order): costs = D[members][:, members].sum(axis=1); move the medoid to the member with
                                       the smallest cost if that cost is strictly smaller than the current medoid's
        stop when no medoid moved
    cluster_centers_ = X[medoids]

PARITY STATUS: the k-medoids restatement is "parity unpinned" against scikit-learn-extra itself (package absent, no network);
`pairwise_distances` IS checked against the scikit-learn in this image (tests/test_oracle_query_points.py), and the float32
`np.sum` / `np.argpartition` / `np.argmin` calls are numpy's own, as in the package.  Everything else in this file follows the
in-tree reference line by line (cv2 is present here, so `goodFeaturesToTrack` / `erode` are the real calls).
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch


# ----------------------------------------------------------------------------- KMedoids (scikit-learn-extra, restated)
def pairwise_euclidean_f32(X: np.ndarray) -> np.ndarray:
    """sklearn.metrics.pairwise_distances(X) for float32 X: float64 accumulation of the squared distance, cast to float32,
    clamp at 0, zero diagonal, sqrt in float32."""
    X64 = X.astype(np.float64)
    xx = (X64 * X64).sum(axis=1)
    d2 = (xx[:, None] + xx[None, :] - 2.0 * (X64 @ X64.T)).astype(np.float32)
    np.maximum(d2, 0, out=d2)
    np.fill_diagonal(d2, 0)
    return np.sqrt(d2, out=d2)


def kmedoids_alternate(X: np.ndarray, n_clusters: int, max_iter: int = 300, info: dict | None = None) -> np.ndarray:
    """-> cluster_centers_ (n_clusters, d), float32 rows of X."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    D = pairwise_euclidean_f32(X)
    if n_clusters > D.shape[0]:
        raise ValueError("The number of medoids must be less than the number of samples")
    medoid_idxs = np.argpartition(np.sum(D, axis=1), n_clusters - 1)[:n_clusters]
    n_iter = 0
    for n_iter in range(0, max_iter):
        old = np.copy(medoid_idxs)
        labels = np.argmin(D[medoid_idxs, :], axis=0)
        for k in range(n_clusters):
            members = np.where(labels == k)[0]
            if len(members) == 0:
                continue
            costs = np.sum(D[members, members[:, np.newaxis]], axis=1)
            best = np.argmin(costs)
            cur = costs[np.argmax(members == medoid_idxs[k])]
            if costs[best] < cur:
                medoid_idxs[k] = members[best]
        if np.all(old == medoid_idxs):
            break
    if info is not None:
        info["n_iter"] = n_iter + 1
        info["medoid_idxs"] = medoid_idxs.copy()
    return X[medoid_idxs]


# ----------------------------------------------------------------------------- query_points.py
def extract_random_mask_points(mask: torch.Tensor, n: int) -> torch.Tensor:
    """query_points.py:29-61"""
    if mask.sum() == 0:
        return torch.zeros((n, 2))
    px = mask.nonzero().float()
    if len(px) < n:
        sel = px.repeat(n // len(px) + 1, 1)[:n]
    else:
        sel = px[torch.randperm(len(px))[:n]]
    return sel.flip(1)


def extract_kmedoid_points(mask: torch.Tensor, n: int, subsample_size: int = 1800, info: dict | None = None) -> torch.Tensor:
    """query_points.py:64-104"""
    if mask.sum() == 0:
        return torch.zeros((n, 2))
    px = mask.nonzero().float()
    if len(px) < n:
        sel = px.repeat(n // len(px) + 1, 1)[:n]
    else:
        px = px[torch.randperm(len(px))[:subsample_size]]
        sel = torch.from_numpy(kmedoids_alternate(px.numpy(), n, info=info)).type(torch.float32)
    return sel.flip(1)


def erode_mask(mask: torch.Tensor, erosion_percentage: float) -> torch.Tensor:
    """query_points.py:165-194"""
    import cv2
    px = mask.nonzero().float()
    diameter = torch.norm(px.max(0)[0] - px.min(0)[0]).item()
    size = int(diameter * erosion_percentage)
    er = cv2.erode(mask.cpu().numpy().astype(np.uint8), np.ones((size, size), np.uint8), iterations=1)
    return torch.from_numpy(er).type(mask.dtype)


def extract_corner_points(image: torch.Tensor, mask: torch.Tensor, n: int, kmedoid_subsample_size: int = 2000) -> torch.Tensor:
    """query_points.py:107-163"""
    import cv2
    if mask.sum() == 0:
        return torch.zeros((n, 2))
    img = image.permute(1, 2, 0).cpu().numpy()
    er = erode_mask(mask, 0.06)
    if er.sum() < 10:
        er = erode_mask(mask, 0.02)
    if er.sum() < 10:
        er = erode_mask(mask, 0.01)
    if er.sum() < 10:
        er = mask
    px = er.nonzero().float()
    diameter = torch.norm(px.max(0)[0] - px.min(0)[0]).item()
    gray = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
    pts = cv2.goodFeaturesToTrack(image=gray, maxCorners=n, qualityLevel=0.001, minDistance=diameter / n,
                                  mask=er.cpu().numpy().astype(np.uint8), blockSize=3, gradientSize=3)
    if pts is None:
        pts = np.empty((0, 1, 2))
    pts = torch.from_numpy(pts).type(torch.float32).squeeze(1)
    if len(pts) < n:
        pts = torch.cat((pts, extract_kmedoid_points(mask, n - pts.shape[0], subsample_size=kmedoid_subsample_size)), dim=0)
    return pts


def extract_mixed_points(query_masks: List[torch.Tensor], timesteps: torch.Tensor, images: torch.Tensor, n: int):
    """query_points.py:197-237"""
    nk, ns = n // 4, n // 3
    nr = n - nk - ns
    lists = []
    if nk > 0:
        lists.append([extract_kmedoid_points(qm, nk) for qm in query_masks])
    if ns > 0:
        lists.append([extract_corner_points(images[int(t.item())], qm, ns) for qm, t in zip(query_masks, timesteps)])
    if nr > 0:
        lists.append([extract_random_mask_points(qm, nr) for qm in query_masks])
    if len(lists) == 1:
        return lists[0]
    return [torch.cat(x, dim=0) for x in zip(*lists)]


# ----------------------------------------------------------------------------- SamPt.extract_query_points (sam_pt.py:238-306)
def _extract_xy(images, query_masks, timesteps, method: str, n: int):
    if method == "kmedoids":
        return [extract_kmedoid_points(qm, n) for qm in query_masks]
    if method == "shi-tomasi":
        return [extract_corner_points(images[int(t.item())], qm, n) for qm, t in zip(query_masks, timesteps)]
    if method == "random":
        return [extract_random_mask_points(qm, n) for qm in query_masks]
    if method == "mixed":
        return extract_mixed_points(query_masks, timesteps, images, n)
    raise NotImplementedError(f"Point selection method {method} not implemented")


def extract_query_points(images, query_masks, timesteps, *, positive_method="kmedoids", negative_method="mixed",
                         positive_points_per_mask=16, negative_points_per_mask=1) -> torch.Tensor:
    query_masks = query_masks.cpu()
    timesteps = timesteps.cpu()
    xy = _extract_xy(images, query_masks, timesteps, positive_method, positive_points_per_mask)
    if negative_points_per_mask > 0:
        neg = _extract_xy(images, [1 - qm for qm in query_masks], timesteps, negative_method, negative_points_per_mask)
        xy = [torch.cat(x, dim=0) for x in zip(xy, neg)]
    xy = torch.stack(xy, dim=0)
    t = timesteps[:, None, None].repeat(1, xy.shape[1], 1)
    return torch.concat([t, xy], dim=2)

"""ORACLE (test infrastructure, NOT product code): the reference's point re-initialisation control flow and patch-similarity
filter, restated.  Only ``tests/`` may import this.

* ``forward_w_reinit`` / ``forward_w_reinit_inner``: /root/reference/sam_pt/modeling/sam_pt.py:355-543, line by line, with the three
  primitives it calls injected as callables (``track_points(images, query_points)``, ``apply_sam(images, traj, vis)``,
  ``extract_query_points(images, masks, timesteps)``) -- so the control flow (horizons, the four re-init variants, invalid-mask
  handling, left/right stitching) can be checked in isolation by driving it with the product's own primitives, whose parity is
  established separately.  (Driving it with the CPU oracle primitives is not meaningful as an end-to-end check: re-initialisation
  samples points from predicted masks, and a single differing mask pixel changes the sampled set.)
* ``patch_similarity``: sam_pt.py:597-682 (`extract_patches_from_points`, `compute_patch_similarity`, the visibility rewrite),
  with ``skimage.color.rgb2lab`` (absent here) restated from its published arithmetic (sRGB -> linear -> XYZ (D65, 2 deg) -> Lab).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

REINIT_FAILED, PATCH_NON_SIMILAR, REJECTED_AFTER = -1, -3, -4


# ----------------------------------------------------------------------------- re-initialisation (sam_pt.py:355-543)
def forward_w_reinit_inner(images, query_points, *, track_points, apply_sam, extract_query_points, reinit_horizon,
                           reinit_point_tracker_horizon, reinit_variant):
    n_frames, channels, height, width = images.shape
    n_masks, points_per_mask, _ = query_points.shape
    assert reinit_point_tracker_horizon >= reinit_horizon
    dev = query_points.device
    trajectories = torch.full((n_frames, n_masks, points_per_mask, 2), torch.nan, dtype=torch.float32, device=dev)
    visibilities = torch.full((n_frames, n_masks, points_per_mask), False, dtype=torch.float32, device=dev)
    scores_per_frame = torch.full((n_frames, n_masks), torch.nan, dtype=torch.float32, device=dev)
    logits = torch.full((n_masks, n_frames, height, width), torch.nan, dtype=torch.float32, device=dev)
    query_points_timestep = query_points[:, 0, 0].int()
    current_query_points = query_points.clone()
    for start_frame in range(int(query_points_timestep.min()), n_frames):
        end_frame = min(start_frame + reinit_horizon, n_frames)
        end_frame_tracker = min(start_frame + reinit_point_tracker_horizon, n_frames)
        current_timesteps = current_query_points[:, 0, 0].int()
        tracked_masks_indices = current_timesteps == start_frame
        if tracked_masks_indices.sum() == 0:
            continue
        query_points_i = current_query_points[tracked_masks_indices].clone()
        query_points_i[:, :, 0] -= start_frame
        assert (query_points_i[:, :, 0] == 0).all()
        trajectories_i, visibilities_i = track_points(images[start_frame:end_frame_tracker], query_points_i)
        trajectories_i = trajectories_i[:reinit_horizon, :, :, :]
        visibilities_i = visibilities_i[:reinit_horizon, :, :]
        _, logits_i, scores_per_frame_i = apply_sam(images[start_frame:end_frame], trajectories_i, visibilities_i)
        logits_i = logits_i.type(torch.float32)
        logits[tracked_masks_indices, start_frame:end_frame] = logits_i
        pred_masks_sam_i = logits_i > 0
        trajectories[start_frame:end_frame, tracked_masks_indices] = trajectories_i
        visibilities[start_frame:end_frame, tracked_masks_indices] = visibilities_i
        scores_per_frame[start_frame:end_frame, tracked_masks_indices] = scores_per_frame_i
        if end_frame == n_frames:
            continue
        area_per_frame = pred_masks_sam_i[:, 1:, :, :].sum([2, 3]).float()
        area_per_frame[area_per_frame <= 25] = torch.nan
        if reinit_horizon // 4 < area_per_frame.shape[1]:
            area_per_frame[:, :reinit_horizon // 4] = torch.nan
        if reinit_variant == "reinit-on-horizon-and-sync-masks":
            next_timestep = reinit_horizon - 1 - 1
            other_timesteps = current_timesteps[current_timesteps > start_frame]
            if len(other_timesteps) > 0:
                next_timestep = min(next_timestep, int(other_timesteps.min()) - start_frame - 1)
            qts = torch.full((pred_masks_sam_i.shape[0],), next_timestep, dtype=torch.int64, device=dev)
        elif reinit_variant == "reinit-at-median-of-area-diff":
            qts = area_per_frame.cpu().nanmedian(dim=1).indices.to(dev)
        elif reinit_variant == "reinit-on-similar-mask-area":
            target_mask_area = pred_masks_sam_i[:, 0, :, :].sum([1, 2])
            area_diff = torch.abs(area_per_frame - target_mask_area[:, None])
            area_diff[area_diff.isnan()] = torch.inf
            qts = area_diff.argmin(dim=1)
        elif reinit_variant == "reinit-on-similar-mask-area-and-sync-masks":
            target_mask_area = pred_masks_sam_i[:, 0, :, :].sum([1, 2])
            area_diff = torch.abs(area_per_frame - target_mask_area[:, None])
            area_diff = area_diff / target_mask_area[:, None]
            area_diff[area_diff.isnan()] = 720
            area_diff_per_frame = area_diff.sum(dim=0)
            other_timesteps = current_timesteps[current_timesteps > start_frame]
            if len(other_timesteps) > 0:
                area_diff_per_frame[int(other_timesteps.min()) - start_frame - 1] -= 36
            qts = torch.full((pred_masks_sam_i.shape[0],), int(area_diff_per_frame.argmin(dim=0)), dtype=torch.int64, device=dev)
        else:
            raise ValueError(f"Unknown reinit variant: {reinit_variant}")
        ar = torch.arange(len(qts), device=dev)
        invalid_masks = area_per_frame[ar, qts] <= 0
        if (~invalid_masks).sum() > 0:
            query_masks = pred_masks_sam_i[:, 1:, :, :][ar, qts].type(torch.float32)
            update = extract_query_points(images[start_frame + 1:end_frame], query_masks[~invalid_masks], qts[~invalid_masks].float())
            valid_tracked_masks = tracked_masks_indices.clone()
            valid_tracked_masks[tracked_masks_indices] = ~invalid_masks
            current_query_points[valid_tracked_masks] = update.to(current_query_points.device)
            current_query_points[valid_tracked_masks, :, 0] += start_frame + 1
        if invalid_masks.sum() > 0:
            invalid_tracked_masks = tracked_masks_indices.clone()
            invalid_tracked_masks[tracked_masks_indices] = invalid_masks
            current_query_points[invalid_tracked_masks, :, 0] = n_frames
            current_query_points[invalid_tracked_masks, :, 1:] = 0
            trajectories[end_frame:, invalid_tracked_masks] = -72
            visibilities[end_frame:, tracked_masks_indices] = REINIT_FAILED
            logits[invalid_tracked_masks, end_frame:] = -float("inf")
    scores = scores_per_frame.nanmean(dim=1)
    return trajectories, visibilities, logits, scores, scores_per_frame


def forward_w_reinit(images, query_points, **kw):
    n_frames = images.shape[0]
    tr, vr, lr, _, sr = forward_w_reinit_inner(images, query_points, **kw)
    qf = query_points.clone()
    qf[:, :, 0] = n_frames - query_points[:, :, 0] - 1
    tl, vl, ll, _, sl = forward_w_reinit_inner(images.flip(0), qf, **kw)
    tl, vl, ll = tl.flip(0), vl.flip(0), ll.flip(1)
    ts = query_points[:, 0, 0].int()
    trajectories = torch.full_like(tr, torch.nan)
    visibilities = torch.full_like(vr, False)
    logits = torch.full_like(lr, torch.nan)
    scores_per_frame = torch.full_like(sr, torch.nan)
    for m, t in enumerate(ts.tolist()):
        trajectories[t:, m], trajectories[:t, m] = tr[t:, m], tl[:t, m]
        visibilities[t:, m], visibilities[:t, m] = vr[t:, m], vl[:t, m]
        logits[m, t:], logits[m, :t] = lr[m, t:], ll[m, :t]
        scores_per_frame[t:, m], scores_per_frame[:t, m] = sr[t:, m], sl[:t, m]
    assert not torch.isnan(trajectories).any() and not torch.isnan(logits).any()
    return trajectories, visibilities, logits, scores_per_frame.nanmean(dim=0), scores_per_frame


# ----------------------------------------------------------------------------- patch-similarity filter (sam_pt.py:597-682)
def rgb2lab(img_u8_hwc: np.ndarray) -> np.ndarray:
    """skimage.color.rgb2lab(uint8 image) (illuminant D65, observer 2), float64."""
    arr = img_u8_hwc.astype(np.float64) / 255.0
    mask = arr > 0.04045
    arr = np.where(mask, np.power((arr + 0.055) / 1.055, 2.4), arr / 12.92)
    m = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])
    xyz = arr @ m.T
    xyz = xyz / np.array([0.95047, 1.0, 1.08883])
    mask = xyz > 0.008856
    f = np.where(mask, np.cbrt(xyz), 7.787 * xyz + 16.0 / 116.0)
    x, y, z = f[..., 0], f[..., 1], f[..., 2]
    return np.stack([116.0 * y - 16.0, 500.0 * (x - y), 200.0 * (y - z)], axis=-1)


def patch_similarity(rgbs_u8, query_points, trajectories, visibilities, patch_size=3, threshold=0.01):
    """rgbs (T,3,H,W) uint8, query_points (N,3), trajectories (T,N,2), visibilities (T,N) float -> (similarities (T,N), vis (T,N))."""
    T, _, h, w = rgbs_u8.shape
    N = query_points.shape[0]
    lab = np.stack([rgb2lab(f) for f in rgbs_u8[:, [2, 1, 0], :, :].permute(0, 2, 3, 1).cpu().numpy()])   # channel swap as the reference
    lab = torch.as_tensor(lab, dtype=torch.float32).permute(0, 3, 1, 2)

    def patches(imgs, pts):
        tmpl = torch.arange(-(patch_size // 2), patch_size // 2 + 1)
        tmpl = torch.stack(torch.meshgrid(tmpl, tmpl, indexing="ij"), dim=-1).reshape(-1, 2)
        xy = pts[:, :, None, :] + tmpl[None, None, :, :] + 0.5
        grid = (xy / torch.tensor([w, h])[None, None, :]) * 2 - 1
        return F.grid_sample(imgs, grid, align_corners=False, mode="bilinear").permute(0, 2, 3, 1)

    tq = query_points[:, 0].long()
    qp = patches(lab[tq], query_points[:, 1:].cpu()[:, None, :]).squeeze(1)           # (N, ps^2, 3)
    tp = patches(lab, trajectories.cpu())                                             # (T, N, ps^2, 3)
    diff = tp.flatten(2, 3) - qp[None].flatten(2, 3)
    sim = torch.exp(-torch.norm(diff, dim=-1) / (2 * patch_size ** 2))
    vis = visibilities.clone().cpu()
    vis[(vis == 1) & ~(sim > threshold)] = PATCH_NON_SIMILAR
    for n in range(N):
        q = int(query_points[n, 0])
        for f in range(q + 1, T):
            if vis[f, n] != PATCH_NON_SIMILAR:
                continue
            vis[f + 1:, n] = REJECTED_AFTER
            break
        for f in range(q - 1, -1, -1):
            if vis[f, n] != PATCH_NON_SIMILAR:
                continue
            vis[:f:, n] = REJECTED_AFTER
            break
    return sim, vis

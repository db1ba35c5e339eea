"""ORACLE (test infrastructure, NOT product code): CPU restatement of the SamPt orchestrator's hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import this.

Restates /root/reference/sam_pt/modeling/sam_pt.py:
* ``SamPt.forward``                    :122-236 (query_points branch only; extract_query_masks is elided because its
                                        output is only asserted on)
* ``SamPt._forward``                   :337-353
* ``SamPt._track_points`` / inner      :545-596, :684-692 (patch-matching branch off by default)
* ``SamPt._apply_sam_to_trajectories`` :694-866 (prepare_points :726-758, predict_mask :760-837)

The reference module itself cannot be imported here (needs segment_anything, skimage, sklearn_extra, SuperGlue), hence
the restatement; the tracker it calls is pinned to the real reference (oracle/pips_ref.py) and the SAM predictor to the
HF cross-check (oracle/sam_ref.py).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import pips_ref, sam_ref

OUTSIDE_FRAME = -2  # PointVisibilityType.OUTSIDE_FRAME, sam_pt/utils/util.py:267-282


def track_points(pips_sd, images_u8, query_points, mask_batch_size: int = 5, tracker=None):
    """_track_points + __track_points_inner (sam_pt.py:545-596, 684-692). images (T,3,H,W) u8; query (M,P,3).
    `tracker(images (1,T,3,H,W) u8, queries (1,N,3)) -> (traj (1,T,N,2), vis (1,T,N) bool)` replaces the PIPS tracker (e.g. the
    CoTracker restatement, configs/model/point_tracker/cotracker.yaml); `pips_sd` is then unused."""
    M, P, _ = query_points.shape
    trajs, viss = [], []
    for i in range(0, M, mask_batch_size):
        q = query_points[i:i + mask_batch_size]
        m = q.shape[0]
        if tracker is not None:
            traj, vis = tracker(images_u8[None], q.reshape(1, m * P, 3))
        else:
            traj, vis = pips_ref.pips_tracker_forward(pips_sd, images_u8[None], q.reshape(1, m * P, 3))
        traj, vis = traj[0], vis[0].float()
        traj = traj.reshape(-1, m, P, 2)
        vis = vis.reshape(-1, m, P)
        h, w = images_u8.shape[-2:]
        vis[traj[:, :, :, 0] / w < 0.01] = OUTSIDE_FRAME
        vis[traj[:, :, :, 1] / h < 0.01] = OUTSIDE_FRAME
        vis[traj[:, :, :, 0] / w > 0.99] = OUTSIDE_FRAME
        vis[traj[:, :, :, 1] / h > 0.99] = OUTSIDE_FRAME
        trajs.append(traj)
        viss.append(vis)
    return torch.cat(trajs, dim=1), torch.cat(viss, dim=1)


@torch.no_grad()
def apply_sam_to_trajectories(predictor: sam_ref.RefSamPredictor, images_u8, trajectories, visibilities, *,
                              positive_points_per_mask: int, negative_points_per_mask: int = 0,
                              iterative_refinement_iterations: int = 12, sam_iou_threshold: float = 0.7,
                              add_other_objects_positive_points_as_negative_points: bool = True,
                              features_cache: Optional[dict] = None, taps: Optional[dict] = None):
    """sam_pt.py:694-866. Returns (pred_scores (M,), logits (M,T,H,W) f32, scores_per_frame (T,M))."""
    T, _, H, W = images_u8.shape
    _, M, Pm, _ = trajectories.shape
    logits = np.full((M, T, H, W), -float("inf"))
    spf = np.full((T, M), -float("inf"))
    ssum, scnt = np.zeros(M), np.zeros(M)

    def prepare_points(f, m):
        pc = trajectories[f, m]
        pl = np.ones((len(pc)), dtype=int)
        if negative_points_per_mask > 0:
            pl[positive_points_per_mask:] = 0
        visible = visibilities[f, m] == 1
        vc = pc[visible].cpu().numpy()
        vl = pl[visible.cpu().numpy()]
        if M > 1 and add_other_objects_positive_points_as_negative_points:
            other = torch.cat([
                trajectories[f, o, :positive_points_per_mask][visibilities[f, o, :positive_points_per_mask] == 1]
                for o in range(M) if o != m], dim=0).cpu().numpy()
            vc = np.concatenate([vc, other], axis=0)
            vl = np.concatenate([vl, np.zeros((len(other)), dtype=int)], axis=0)
        return vc, vl

    def predict_mask(vc, vl, f, m):
        if len(vc) == 0:
            return np.full((H, W), -float("inf"), dtype=np.float64), None
        vc = torch.as_tensor(predictor.transform.apply_coords(vc, predictor.original_size), dtype=torch.float)
        vl = torch.as_tensor(vl, dtype=torch.int)
        if negative_points_per_mask == 0:
            ml, iou, low = predictor.predict_torch(vc[None], vl[None], None, None, False, True)
        else:
            _, _, low = predictor.predict_torch(vc[vl == 1][None], vl[vl == 1][None], None, None, False, True)
            ml, iou, low = predictor.predict_torch(vc[None], vl[None], None, low, False, True)
        n_ref = 0
        box_margin = float("inf")
        for _ in range(iterative_refinement_iterations):
            mm = ml[0, 0] > 0
            if mm.sum() < 2:
                break
            box_margin = min(box_margin, _box_edge_margin(ml[0, 0]))
            yx = mm.nonzero()
            box = torch.tensor([yx[:, 1].min(), yx[:, 0].min(), yx[:, 1].max(), yx[:, 0].max()], dtype=torch.float)
            ml, iou, low = predictor.predict_torch(vc[None], vl[None], box[None, None, :], low, False, True)
            n_ref += 1
        if taps is not None:
            taps.setdefault("low_res", {})[(f, m)] = low[0, 0].clone()
            taps.setdefault("n_refine", {})[(f, m)] = n_ref
            taps.setdefault("box_margin", {})[(f, m)] = box_margin
        s = iou[0, 0].cpu().numpy()
        if s < sam_iou_threshold:
            return np.full((H, W), -float("inf"), dtype=np.float64), s
        return ml[0, 0].cpu().numpy(), s

    for f in range(T):
        if features_cache is not None and f in features_cache:
            predictor.features = features_cache[f]["features"]
            predictor.interm = features_cache[f].get("interm")
            predictor.original_size = (H, W)
            predictor.input_size = sam_ref.get_preprocess_shape(H, W, predictor.cfg.img_size)
        else:
            predictor.set_image(images_u8[f].permute(1, 2, 0).cpu().numpy())
            if features_cache is not None:
                features_cache[f] = {"features": predictor.features, "interm": predictor.interm}
        for m in range(M):
            vc, vl = prepare_points(f, m)
            ml, s = predict_mask(vc, vl, f, m)
            logits[m, f] = ml
            if s is not None:
                spf[f, m] = s
                ssum[m] += s
                scnt[m] += 1
    pred_scores = ssum / np.where(scnt != 0, scnt, 1)
    return pred_scores, torch.from_numpy(logits).float(), torch.from_numpy(spf).float()


def _box_edge_margin(logit: torch.Tensor) -> float:
    """How far the refinement box of `logit > 0` (sam_pt.py:793-800: min / max of the positive pixels' coordinates) is from
    moving by one pixel: for every edge, the smaller of (largest logit of the edge row / column: it must stay > 0) and (minus the
    largest logit of the row / column just outside: it must stay <= 0).  A parity test may only hold a frame to the mask-IoU bar
    when this margin, minimised over the 12 refinement iterations, exceeds the decoder's own parity tolerance: a box that moves by
    one pixel changes every later logit of the chain (test infrastructure, not part of the reference)."""
    rows, cols = logit.max(dim=1).values, logit.max(dim=0).values
    out = float("inf")
    for v in (rows, cols):
        pos = (v > 0).nonzero()[:, 0]
        lo, hi = int(pos.min()), int(pos.max())
        out = min(out, float(v[lo]), float(v[hi]))
        if lo > 0:
            out = min(out, float(-v[lo - 1]))
        if hi < v.numel() - 1:
            out = min(out, float(-v[hi + 1]))
    return out


@torch.no_grad()
def sampt_forward(pips_sd, predictor: sam_ref.RefSamPredictor, video: Dict, *, positive_points_per_mask: int,
                  negative_points_per_mask: int = 0, iterative_refinement_iterations: int = 12,
                  sam_iou_threshold: float = 0.7, point_tracker_mask_batch_size: int = 5,
                  features_cache: Optional[dict] = None, taps: Optional[dict] = None, tracker=None):
    """SamPt.forward with the query_points branch (sam_pt.py:122-236)."""
    images = torch.stack(video["image"], dim=0)
    assert images.dtype == torch.uint8
    if video.get("query_points") is None:
        raise ValueError("No query points or masks provided")
    qp = video["query_points"]
    traj, vis = track_points(pips_sd, images, qp, point_tracker_mask_batch_size, tracker=tracker)
    _, logits, spf = apply_sam_to_trajectories(
        predictor, images, traj, vis, positive_points_per_mask=positive_points_per_mask,
        negative_points_per_mask=negative_points_per_mask,
        iterative_refinement_iterations=iterative_refinement_iterations, sam_iou_threshold=sam_iou_threshold,
        features_cache=features_cache, taps=taps)
    scores = spf.mean(dim=0)
    target_hw = tuple(video["target_hw"])
    rf = torch.tensor(target_hw) / torch.tensor(logits.shape[-2:])
    if tuple(logits.shape[-2:]) != target_hw:
        logits = F.interpolate(logits, size=target_hw, mode="bilinear", align_corners=False)
    traj = traj * rf
    return {"logits": [m for m in logits], "scores": scores.tolist(), "scores_per_frame": spf.tolist(),
            "trajectories": traj, "visibilities": vis}

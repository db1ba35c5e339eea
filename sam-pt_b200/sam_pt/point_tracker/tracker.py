"""`PointTracker` interface of the reference (sam_pt/point_tracker/tracker.py:8-118): same method names, argument
meaning and result keys, so harness code written against the reference runs unchanged."""
from abc import ABC, abstractmethod
from typing import Tuple

import torch
from torch import nn


class PointTracker(ABC, nn.Module):
    @abstractmethod
    def forward(self, rgbs, query_points) -> Tuple[torch.Tensor, torch.Tensor]:
        """rgbs (B,T,3,H,W) uint8 0..255, query_points (B,N,3)=(t,x,y) -> trajectories (B,T,N,2), visibilities (B,T,N)."""

    def evaluate_batch(self, rgbs, query_points, trajectories_gt=None, visibilities_gt=None):
        """reference tracker.py:47-83: forward + shape check + results moved to the CPU."""
        traj, vis = self.forward(rgbs, query_points)
        B, T, N = rgbs.shape[0], rgbs.shape[1], query_points.shape[1]
        assert traj.shape == (B, T, N, 2)

        def cpu(t):
            return t.detach().clone().cpu() if t is not None else None

        return {"trajectories_pred": cpu(traj), "visibilities_pred": cpu(vis), "query_points": cpu(query_points),
                "trajectories_gt": cpu(trajectories_gt), "visibilities_gt": cpu(visibilities_gt)}

    @classmethod
    def unpack_results(cls, packed_results, batch_idx):
        """reference tracker.py:85-118: one dict per (video, point)."""
        out = []
        tp = packed_results["trajectories_pred"]
        for b in range(tp.shape[0]):
            for n in range(tp.shape[2]):
                r = {"idx": f"{batch_idx}_{b}_{n}", "iter": batch_idx, "video_idx": b, "point_idx_in_video": n,
                     "query_point": packed_results["query_points"][b, n, :], "trajectory_pred": tp[b, :, n, :],
                     "visibility_pred": packed_results["visibilities_pred"][b, :, n]}
                if packed_results["trajectories_gt"] is not None:
                    r["trajectory_gt"] = packed_results["trajectories_gt"][b, :, n, :]
                    r["visibility_gt"] = packed_results["visibilities_gt"][b, :, n]
                out.append(r)
        return out

"""Demo wire formats of the reference: the query-point text file and the frame-folder loader, plus the
post-processing `run_inference` applies to the model output.  Mirrors /root/reference/demo/demo.py:

* `load_query_points`  demo.py:225-252, format data/demo_data/README.md:18-38
    number_of_positive_points
    mask_timestep ; x,y x,y ...            (one line per mask; positives first, then negatives)
* `load_demo_data`     demo.py:191-222 (sorted *.jpg then *.png, frame_stride, max_frames, optional longest-side resize)
* `run_inference`      demo.py:114-158 (background channel of zero logits, -1e8 before each mask's query frame)

The GUI parts (cv2.imshow loops, interactive point picking) are outside the hot path and not built.
"""
from __future__ import annotations

import glob
import os

import torch


def load_query_points(query_points_path, frame_stride, resize_factor):
    """-> (query_points (n_masks, n_points, 3) float32 = (t, x, y), num_positive_points)."""
    timesteps, points = [], []
    with open(query_points_path, "r") as f:
        lines = f.readlines()
    num_positive_points = int(lines[0].strip())
    for line in lines[1:]:
        line = line.strip()
        if not line:
            continue
        timestep, queries_xy = line.split(";")
        xy = torch.tensor([[float(v) for v in tok.split(",")] for tok in queries_xy.split()])
        assert xy.ndim == 2 and xy.shape[1] == 2, f"malformed query point line: {line!r}"
        xy = xy * resize_factor
        timestep = int(timestep)
        assert timestep % frame_stride == 0
        timesteps.append(timestep // frame_stride)
        points.append(xy)
    xy = torch.stack(points)                       # every mask must carry the same number of points (torch.stack, demo.py:247)
    t = torch.tensor(timesteps, dtype=torch.float32)[:, None, None]
    return torch.cat([t.repeat(1, xy.shape[1], 1), xy], dim=2), num_positive_points


def save_query_points(query_points_path, query_points: torch.Tensor, num_positive_points: int):
    """Inverse of `load_query_points` (the format the interactive demo writes, demo.py:300-313)."""
    with open(query_points_path, "w") as f:
        f.write(f"{int(num_positive_points)}\n")
        for mask in query_points:
            t = int(mask[0, 0].item())
            f.write(f"{t} ; " + " ".join(f"{x:.6f},{y:.6f}" for x, y in mask[:, 1:].tolist()) + "\n")


def load_demo_data(frames_path, query_points_path, frame_stride=1, longest_side_length=None, annot_size=8, annot_line_width=4,
                   max_frames=None):
    """-> (rgbs (T,3,H,W) uint8, num_positive_points, query_points)."""
    import cv2
    assert query_points_path is not None
    frames = sorted(glob.glob(os.path.join(frames_path, "*.jpg")))
    frames += sorted(glob.glob(os.path.join(frames_path, "*.png")))
    assert len(frames) > 0, f"No frames found in {frames_path}"
    frames = frames[::frame_stride]
    if max_frames is not None:
        frames = frames[:max_frames]
    rgbs = []
    resize_factor = 1.0
    for frame in frames:
        img = cv2.cvtColor(cv2.imread(frame), cv2.COLOR_BGR2RGB)
        img = torch.from_numpy(img).permute(2, 0, 1)
        if longest_side_length is not None:
            resize_factor = longest_side_length / max(img.shape[1], img.shape[2])
            img = torch.nn.functional.interpolate(img[None], scale_factor=resize_factor)[0]   # nearest, as the reference
        rgbs.append(img)
    rgbs = torch.stack(rgbs)
    query_points, num_positive_points = load_query_points(query_points_path, frame_stride, resize_factor)
    return rgbs, num_positive_points, query_points


def run_inference(model, rgbs, query_points, target_hw):
    """demo.py:114-158: forward pass + background channel + suppression before the query frame.
    -> (logits (T, 1 + n_masks, H, W), trajectories, visibilities, scores)."""
    n_masks, n_points_per_mask, _ = query_points.shape
    n_frames = rgbs.shape[0]
    video = {"video_id": 0, "image": [rgb for rgb in rgbs], "target_hw": target_hw, "query_points": query_points}
    outputs = model(video)
    logits_list = outputs["logits"]
    trajectories, visibilities, scores = outputs["trajectories"], outputs["visibilities"], outputs["scores"]
    logits = torch.stack([torch.zeros_like(logits_list[0])] + logits_list, dim=1)
    assert logits.shape[0] == n_frames
    assert trajectories.shape == (n_frames, n_masks, n_points_per_mask, 2)
    for i, timestep in enumerate(query_points[:, 0, 0].tolist()):
        logits[:int(timestep), i + 1] = -1e8
    return logits, trajectories, visibilities, scores

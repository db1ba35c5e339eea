"""`CoTrackerPointTracker` drop-in (reference sam_pt/point_tracker/cotracker/tracker.py:27-170): same constructor kwargs
(configs/model/point_tracker/cotracker.yaml), same forward contract.  The CoTracker model itself (un-vendored upstream
package in the reference) is `sam_pt.point_tracker.cotracker.cotracker.CoTracker`, which runs in libsampt_b200."""
from ctypes import c_int

import torch

from sam_pt.point_tracker.cotracker.cotracker import CoTracker
from sam_pt.point_tracker.tracker import PointTracker
from sampt_b200 import native


def get_points_on_a_grid(grid_size, interp_shape, device="cpu"):
    """upstream cotracker.get_points_on_a_grid as used at reference tracker.py:100: (1, g*g, 2) xy, margin W//64."""
    if grid_size == 1:
        return torch.tensor([interp_shape[1] / 2, interp_shape[0] / 2], device=device)[None, None]
    gy, gx = torch.meshgrid(torch.linspace(0, grid_size - 1, grid_size, device=device),
                            torch.linspace(0, grid_size - 1, grid_size, device=device), indexing="ij")
    step = interp_shape[1] // 64
    gy = step + gy.reshape(1, -1) / float(grid_size - 1) * (interp_shape[0] - step * 2)
    gx = step + gx.reshape(1, -1) / float(grid_size - 1) * (interp_shape[1] - step * 2)
    return torch.stack([gx, gy], dim=-1)


def build_cotracker(checkpoint_path=None):
    """upstream cotracker.models.build_cotracker.build_cotracker: stride-4 / window-8 model + optional checkpoint
    ({"model": state_dict} or a bare state_dict).  `None` keeps the seeded random initialisation (synthetic benches)."""
    model = CoTracker(S=8, stride=4, space_depth=6, time_depth=6)
    if checkpoint_path is not None:
        name = str(checkpoint_path).split("/")[-1].split(".")[0]
        if name != "cotracker_stride_4_wind_8":
            raise NotImplementedError(f"only cotracker_stride_4_wind_8 is built (configs/model/point_tracker/cotracker.yaml:2), got {name}")
        with open(checkpoint_path, "rb") as f:
            sd = torch.load(f, map_location="cpu")
        if "model" in sd:
            sd = sd["model"]
        model.load_state_dict(sd)
    return model


class CoTrackerPointTracker(PointTracker):
    def __init__(self, checkpoint_path, interp_shape, visibility_threshold, support_grid_size, support_grid_every_n_frames,
                 add_debug_visualisations):
        super().__init__()
        self.checkpoint_path = checkpoint_path
        self.interp_shape = tuple(int(v) for v in interp_shape)
        self.visibility_threshold = visibility_threshold
        self.support_grid_size = support_grid_size
        self.support_grid_every_n_frames = support_grid_every_n_frames
        self.add_debug_visualisations = add_debug_visualisations
        if add_debug_visualisations:
            raise NotImplementedError("debug visualisations (reference tracker.py:119-150) are outside the hot path")
        print(f"Loading CoTracker model from {self.checkpoint_path}")
        self.model = build_cotracker(self.checkpoint_path)
        if torch.cuda.is_available():
            self.model.to("cuda")
        self.model.eval()

    @property
    def device(self):
        return self.model.norm.weight.device

    def resize_clip(self, frames_u8: torch.Tensor) -> torch.Tensor:
        """(T,3,H,W) uint8 -> (T,3,*interp_shape) float32: F.interpolate(..., mode="bilinear") of reference tracker.py:79-81."""
        ctx = self.model.native_context()
        T, C, H, W = frames_u8.shape
        Ho, Wo = self.interp_shape
        out = torch.empty((T, C, Ho, Wo), device=frames_u8.device, dtype=torch.float32)
        native.check(native.lib().sampt_resize_bilinear_u8_f32(ctx.handle, native.ptr(frames_u8.contiguous()), c_int(T * C), c_int(H),
                                                               c_int(W), c_int(Ho), c_int(Wo), native.ptr(out),
                                                               native.stream_ptr()), "resize_bilinear")
        return out

    def forward(self, rgbs, query_points):
        """rgbs (1,T,3,H,W) uint8 (or float 0..255), query_points (1,N,3)=(t,x,y) -> (1,T,N,2) float32, (1,T,N) bool.
        Resize to interp_shape, add the support grid every n frames, forward pass, time-reversed pass filling the frames
        before each query (`traj == 0`), rescale to the input resolution (reference tracker.py:72-170)."""
        if rgbs.shape[0] != 1:
            raise NotImplementedError("batch size 1 (the SAM-PT hot path tracks one clip at a time)")
        dev = self.device
        frames = rgbs[0].to(dev)
        if frames.dtype != torch.uint8:
            frames = frames.round().clamp(0, 255).to(torch.uint8)
        T, _, H, W = frames.shape
        pyr = self.model.encode_frames(self.resize_clip(frames))
        return self._track_pyramid(pyr, query_points, T, H, W)

    # ---- frame-sharded multi-GPU path (SamPt.forward_clips_sharded): encoder on the owned frames, windows on gathered features
    def shard_features(self, frames_u8):
        """(n,3,H,W) uint8 frames this rank owns -> (n,ih/4,iw/4,128) fp32 features at the interp resolution (6.3 MB/frame @384x512)."""
        frames_u8 = frames_u8.to(self.device)
        if frames_u8.shape[0] == 0:
            ih, iw = self.interp_shape
            return torch.empty((0, ih // 4, iw // 4, 128), device=self.device, dtype=torch.float32)
        return self.model.fnet_frames(self.resize_clip(frames_u8))

    def track_on_features(self, fmaps, query_points, frame_hw):
        """fmaps (T,ih/4,iw/4,128) of every frame in frame order (after the all-gather); frame_hw = input resolution (H, W)."""
        return self._track_pyramid(self.model.build_pyramid(fmaps), query_points, fmaps.shape[0], int(frame_hw[0]), int(frame_hw[1]))

    def _track_pyramid(self, pyr, query_points, T, H, W):
        dev = self.device
        n_points = query_points.shape[1]
        ih, iw = self.interp_shape
        q = query_points[0].float().to(dev).clone()
        q[:, 1] *= iw / W
        q[:, 2] *= ih / H
        if self.support_grid_size > 0:
            grids = []
            for i in range(0, T, self.support_grid_every_n_frames):
                g = get_points_on_a_grid(self.support_grid_size, self.interp_shape, device=dev)[0]
                grids.append(torch.cat([torch.full_like(g[:, :1], float(i)), g], dim=1))
            q = torch.cat([q] + grids, dim=0)
        # CoTrackerForShortVideosWrapper (tracker.py:12-24): clips shorter than the window repeat their last frame
        pad = max(self.model.S - T, 0)
        order_fwd = list(range(T)) + [T - 1] * pad
        order_bwd = list(range(T - 1, -1, -1)) + [0] * pad
        traj, vis = self.model.track(pyr, q, order_fwd, iters=6)
        traj, vis = traj[:T], vis[:T]
        # _compute_backward_tracks (tracker.py:154-170)
        qb = q.clone()
        qb[:, 0] = T - qb[:, 0] - 1
        traj_b, vis_b = self.model.track(pyr, qb, order_bwd, iters=6)
        traj_b, vis_b = traj_b[:T].flip(0), vis_b[:T].flip(0)
        mask = traj == 0
        traj = torch.where(mask, traj_b, traj)
        vis = torch.where(mask[:, :, 0], vis_b, vis)
        traj = traj[:, :n_points].clone()
        vis = vis[:, :n_points] > self.visibility_threshold
        traj[:, :, 0] *= W / float(iw)
        traj[:, :, 1] *= H / float(ih)
        return traj[None], vis[None]

"""`SamPt` drop-in (reference sam_pt/modeling/sam_pt.py:21-866): same 18 constructor kwargs, `.device`, `forward(video)`
contract and output dict.  The per-frame work is re-organised for the GPU:

* the clip is uploaded once (uint8), the tracker consumes it on the device;
* SAM's encoder runs on batches of frames (tcgen05 GEMMs need M = B*4096 rows to fill 148 SMs);
* `predict_mask`'s 1-2 + <=12 `predict_torch` calls per (frame, mask) are ONE native call with the `area < 2` break test
  evaluated on the device (the reference synchronises ~6x per refinement iteration, sam_pt.py:811-820);
* one device->host copy per clip (trajectories + visibilities, a few KB) replaces the per-frame copies; the host then
  prepares every frame's prompt exactly as `prepare_points` does (sam_pt.py:726-758).
"""
import os
from typing import Optional

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from sam_pt.point_tracker import PointTracker
from sam_pt.utils.util import PointVisibilityType


class SamPt(nn.Module):
    def __init__(self, point_tracker: PointTracker, sam_predictor, sam_iou_threshold: float,
                 positive_point_selection_method: str, negative_point_selection_method: str, positive_points_per_mask: int,
                 negative_points_per_mask: int, add_other_objects_positive_points_as_negative_points: bool,
                 max_other_objects_positive_points: Optional[int], point_tracker_mask_batch_size: int,
                 iterative_refinement_iterations: bool, use_patch_matching_filtering: bool, patch_size: int,
                 patch_similarity_threshold: float, use_point_reinit: bool, reinit_point_tracker_horizon: int,
                 reinit_horizon: int, reinit_variant: str):
        super().__init__()
        self.point_tracker = point_tracker
        self.sam_predictor = sam_predictor
        self.sam_iou_threshold = sam_iou_threshold
        self._sam = sam_predictor.model  # makes .to(device) reach SAM (reference sam_pt.py:95-96)
        self.iterative_refinement_iterations = iterative_refinement_iterations
        self.positive_point_selection_method = positive_point_selection_method
        self.negative_point_selection_method = negative_point_selection_method
        self.positive_points_per_mask = positive_points_per_mask
        self.negative_points_per_mask = negative_points_per_mask
        self.add_other_objects_positive_points_as_negative_points = add_other_objects_positive_points_as_negative_points
        self.max_other_objects_positive_points = max_other_objects_positive_points
        self.point_tracker_mask_batch_size = point_tracker_mask_batch_size
        self.use_patch_matching_filtering = use_patch_matching_filtering
        self.patch_size = patch_size
        self.patch_similarity_threshold = patch_similarity_threshold
        self.use_point_reinit = use_point_reinit
        self.reinit_point_tracker_horizon = reinit_point_tracker_horizon
        self.reinit_horizon = reinit_horizon
        self.reinit_variant = reinit_variant
        self.encoder_batch = 8          # frames per ViT launch
        # run the ViT encoder (tensor pipe) on its own stream, concurrently with the PIPS tracker and the mask decoder
        # (fp32 CUDA-core pipes, latency bound): the two halves of the path do not depend on each other until decode
        self.overlap_streams = os.environ.get("SAMPT_OVERLAP", "1") != "0"
        self._enc_stream = None
        # the per-frame decode chains (13 predict_torch calls = ~500 tiny kernels each) are latency bound and independent:
        # replay them round-robin on several streams, each with its own CUDA-graph instance / buffers
        self.decode_streams = int(os.environ.get("SAMPT_DECODE_STREAMS", "8"))
        self._dec_streams = None
        self.outputs_on_cpu = False     # reference returns CPU tensors; keeping them on the device avoids a 82 MB copy
        self.frame_annotations = []

    @property
    def device(self):
        return self._sam.device

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, video):
        """reference sam_pt.py:122-236 (query_points branch; query_masks needs the k-medoids sampler, §8f 'next')."""
        if self.training:
            raise NotImplementedError(f"{self._get_name()} does not support training...")
        frames = video["image"]
        assert frames[0].dtype == torch.uint8, "Input images must be in uint8 format (0-255)"
        # upload frame by frame (async from pinned memory) and stack ON the device: no 60 MB host-side torch.stack copy
        images_dev = torch.stack([f.to(self.device, non_blocking=True) for f in frames], dim=0)
        n_frames, channels, height, width = images_dev.shape
        if video.get("query_masks") is not None:
            raise NotImplementedError("query_masks (k-medoids / Shi-Tomasi point sampling, sam_pt/utils/query_points.py) is "
                                      "outside the built hot path; pass query_points (SURVEY §8f item 1)")
        if video.get("query_points") is None:
            raise ValueError("No query points or masks provided")
        if self.use_point_reinit or self.use_patch_matching_filtering:
            raise NotImplementedError("point re-initialisation / patch-similarity filtering are off in the reference "
                                      "defaults (configs/model/sam_pt.yaml:23,27) and not built (SURVEY §8f item 3)")
        print("SAM-PT: Using query points")
        query_points = video["query_points"]
        n_masks, n_points_per_mask, _ = query_points.shape
        self.frame_annotations = [[] for _ in range(n_frames)]

        trajectories, visibilities, logits, scores, scores_per_frame = self._forward(images_dev, query_points)

        target_hw = tuple(int(v) for v in video["target_hw"])
        resize_factor = torch.tensor(target_hw) / torch.tensor(logits.shape[-2:])
        assert (resize_factor[0] - resize_factor[1]).abs().item() < 0.01, "The resizing should have been isotropic"
        if tuple(logits.shape[-2:]) != target_hw:
            logits = F.interpolate(logits, size=target_hw, mode="bilinear", align_corners=False)
        trajectories = trajectories * resize_factor.to(trajectories.device)

        assert logits.shape == (n_masks, n_frames, target_hw[0], target_hw[1])
        assert scores.shape == (n_masks,)
        assert scores_per_frame.shape == (n_frames, n_masks)
        assert trajectories.shape == (n_frames, n_masks, n_points_per_mask, 2)
        assert visibilities.shape == (n_frames, n_masks, n_points_per_mask)
        if self.outputs_on_cpu:
            logits, trajectories, visibilities = logits.cpu(), trajectories.cpu(), visibilities.cpu()
        return {"logits": [m for m in logits], "scores": scores.tolist(), "scores_per_frame": scores_per_frame.tolist(),
                "trajectories": trajectories, "visibilities": visibilities}

    def _forward(self, images, query_points):
        pre = self._start_encoder(images) if self.overlap_streams else None
        trajectories, visibilities = self._track_points(images, query_points)
        _, logits, scores_per_frame = self._apply_sam_to_trajectories(images, trajectories, visibilities, pre=pre)
        scores = scores_per_frame.mean(dim=0)
        return trajectories, visibilities, logits, scores, scores_per_frame

    @torch.no_grad()
    def _start_encoder(self, images):
        """Enqueue the SAM image encoder for every chunk of `images` on the encoder stream; returns per-chunk
        (features, event) so the decode loop on the main stream can wait for exactly the chunk it needs."""
        pred = self.sam_predictor
        main = torch.cuda.current_stream()
        if self._enc_stream is None:
            pred.model.native_context()  # weight registration happens on the main stream, once
            self._enc_stream = torch.cuda.Stream(device=self.device)
        es = self._enc_stream
        es.wait_stream(main)  # the uploaded frames
        B = max(1, int(self.encoder_batch))
        want_interm = pred._uses_interm()
        out = []
        with torch.cuda.stream(es):
            for f0 in range(0, images.shape[0], B):
                enc = pred.encode_frames(images[f0:f0 + B], want_interm=want_interm)
                ev = torch.cuda.Event()
                ev.record(es)
                for t in (enc if isinstance(enc, tuple) else (enc,)):
                    t.record_stream(main)
                out.append((enc, ev))
        return out

    # ------------------------------------------------------------------------------------------------ tracking
    def _track_points(self, rgbs, query_points):
        """reference sam_pt.py:545-596,684-692: batches of `point_tracker_mask_batch_size` masks, flatten (M,P,3)->(1,N,3),
        bool->float visibilities, out-of-frame relabel.  Stays on the device."""
        num_masks, points_per_mask, _ = query_points.shape
        self.point_tracker.eval()
        trajs, viss = [], []
        h, w = rgbs.shape[-2:]
        for i in range(0, num_masks, self.point_tracker_mask_batch_size):
            q = query_points[i:i + self.point_tracker_mask_batch_size].to(self.device)
            m = q.shape[0]
            with torch.no_grad():
                traj, vis = self.point_tracker.to(self.device)(rgbs.unsqueeze(0), q.reshape(1, m * points_per_mask, 3))
            traj = traj[0].reshape(-1, m, points_per_mask, 2)
            vis = vis[0].float().reshape(-1, m, points_per_mask)
            out = float(PointVisibilityType.OUTSIDE_FRAME.value)
            oob = (traj[..., 0] / w < 0.01) | (traj[..., 1] / h < 0.01) | (traj[..., 0] / w > 0.99) | (traj[..., 1] / h > 0.99)
            vis = torch.where(oob, torch.full_like(vis, out), vis)
            trajs.append(traj)
            viss.append(vis)
        return torch.cat(trajs, dim=1), torch.cat(viss, dim=1)

    # ------------------------------------------------------------------------------------------------ SAM
    @torch.no_grad()
    def _apply_sam_to_trajectories(self, images, trajectories, visibilities, pre=None):
        """reference sam_pt.py:694-866.  images (T,3,H,W) uint8 on the device."""
        n_frames = images.shape[0]
        logits, scores_pf, counted = self._apply_sam_to_frames(images, list(range(n_frames)), trajectories, visibilities, pre=pre)
        counted_d = counted.to(self.device)
        cnt = counted_d.sum(dim=0).clamp(min=1)
        pred_scores = torch.where(counted_d, scores_pf, torch.zeros_like(scores_pf)).sum(dim=0) / cnt
        return pred_scores, logits, scores_pf

    @torch.no_grad()
    def _apply_sam_to_frames(self, images, frame_ids, trajectories, visibilities, pre=None):
        """SAM on a subset of frames: images (n,3,H,W) uint8 on the device are the frames `frame_ids` of the clip whose
        full-clip trajectories (T,M,P,2) / visibilities (T,M,P) are given.  Returns logits (M,n,H,W), scores (n,M) on the
        device and `counted` (n,M) bool on the host (frames with at least one visible point)."""
        n_sub, _, height, width = images.shape
        n_frames, n_masks, points_per_mask, _ = trajectories.shape
        assert visibilities.shape == (n_frames, n_masks, points_per_mask)
        dev = self.device
        pred = self.sam_predictor
        # the only device->host copy before the results: a few KB of trajectories / visibility codes
        traj_h = trajectories.detach().cpu()
        vis_h = visibilities.detach().cpu()

        def prepare_points(f, m):  # sam_pt.py:726-758
            pc = traj_h[f, m]
            labels = np.ones((len(pc)), dtype=int)
            if self.negative_points_per_mask > 0:
                labels[self.positive_points_per_mask:] = 0
            vmask = (vis_h[f, m] == 1)
            coords = pc[vmask].numpy()
            labels = labels[vmask.numpy()]
            if n_masks > 1 and self.add_other_objects_positive_points_as_negative_points:
                other = torch.cat([traj_h[f, o, :self.positive_points_per_mask][vis_h[f, o, :self.positive_points_per_mask] == 1]
                                   for o in range(n_masks) if o != m], dim=0).numpy()
                if self.max_other_objects_positive_points is not None and len(other) > self.max_other_objects_positive_points:
                    idx = np.random.choice(len(other), self.max_other_objects_positive_points, replace=False)
                    other = other[idx]
                coords = np.concatenate([coords, other], axis=0)
                labels = np.concatenate([labels, np.zeros((len(other)), dtype=int)], axis=0)
            return coords, labels

        logits = torch.full((n_masks, n_sub, height, width), -float("inf"), device=dev, dtype=torch.float32)
        scores_pf = torch.full((n_sub, n_masks), -float("inf"), device=dev, dtype=torch.float32)
        counted = torch.zeros((n_sub, n_masks), dtype=torch.bool)
        n_ref = int(self.iterative_refinement_iterations) if self.iterative_refinement_iterations else 0
        B = max(1, int(self.encoder_batch))
        want_interm = pred._uses_interm()
        main = torch.cuda.current_stream()
        nslot = max(1, int(self.decode_streams))
        if nslot > 1 and (self._dec_streams is None or len(self._dec_streams) != nslot):
            self._dec_streams = [torch.cuda.Stream(device=dev) for _ in range(nslot)]
        used = set()
        thr = float(self.sam_iou_threshold)
        for ci, f0 in enumerate(range(0, n_sub, B)):
            chunk = images[f0:f0 + B]
            ev = None
            if pre is not None:  # encoder output produced on the encoder stream (see _start_encoder)
                enc, ev = pre[ci]
            else:
                enc = pred.encode_frames(chunk, want_interm=want_interm)
                if nslot > 1:
                    ev = torch.cuda.Event()
                    ev.record(main)
            feats, interm = enc if want_interm else (enc, None)
            for j in range(chunk.shape[0]):
                i = f0 + j
                f = frame_ids[i]
                slot = i % nslot
                stream = self._dec_streams[slot] if nslot > 1 else main
                if nslot > 1 and slot not in used:
                    stream.wait_stream(main)  # logits / scores buffers were created on the main stream
                    used.add(slot)
                with torch.cuda.stream(stream):
                    if ev is not None:
                        stream.wait_event(ev)
                    pred.set_frames_features((height, width), (feats[j:j + 1], interm[j:j + 1]) if want_interm else feats[j:j + 1])
                    for m in range(n_masks):
                        coords, labels = prepare_points(f, m)
                        if len(coords) == 0:
                            continue  # all points invisible -> mask stays -inf, score -inf (sam_pt.py:766-767,855)
                        c1024 = torch.as_tensor(pred.transform.apply_coords(coords, pred.original_size), dtype=torch.float, device=dev)
                        lab = torch.as_tensor(labels, dtype=torch.int, device=dev)
                        iou, _, _ = pred.predict_refine(c1024, lab, 1 if self.negative_points_per_mask > 0 else 0, n_ref, logits[m, i],
                                                        slot=slot)
                        # "Mask is empty if SAM's IoU score is too low" (sam_pt.py:833-835), without a host round trip
                        logits[m, i] = torch.where(iou[0] < thr, torch.full_like(logits[m, i], -float("inf")), logits[m, i])
                        scores_pf[i, m] = iou[0]
                        counted[i, m] = True
            if nslot > 1:
                for t in (enc if isinstance(enc, tuple) else (enc,)):
                    for sl in used:
                        t.record_stream(self._dec_streams[sl])
        if nslot > 1:
            for sl in used:
                main.wait_stream(self._dec_streams[sl])
        return logits, scores_pf, counted

    # ------------------------------------------------------------------------------------------------ multi-GPU
    @torch.no_grad()
    def forward_clips_sharded(self, videos, gather_logits: bool = False):
        """Frame-sharded processing of `len(videos)` clips across the ranks of the default process group:
        rank r owns frames {f : f mod G == r} of EVERY clip.
          A. local : PIPS encoder (fnet) on the owned frames of every clip
          B. NCCL  : ONE all-gather of the fp32 feature maps (13 MB/frame @480x854) -> every rank holds all features
          C. local : linked tracker chain of clip c on rank c mod G (pyramid built locally after the gather)
          D. NCCL  : all-gather of the (T,N,3) trajectories/visibilities (a few KB)
          E. local : SAM encode + prompt/mask decode on the owned frames of every clip
        Returns, per clip, {"trajectories","visibilities","logits" (M, n_owned, H, W), "frame_ids", "scores_per_frame"}
        (logits stay sharded unless gather_logits)."""
        import torch.distributed as dist
        from sampt_b200 import sharding
        world, rank = dist.get_world_size(), dist.get_rank()
        dev = self.device
        trk = self.point_tracker.to(dev)
        if not hasattr(trk, "track_on_features"):
            raise NotImplementedError("frame-sharded tracking is built for the PIPS tracker (BASELINE configs[3]); CoTracker clips "
                                      "run one clip per GPU through SamPt.forward")
        C = len(videos)
        T = len(videos[0]["image"])
        own = sharding.owned_frames(T, rank, world)
        # A. upload only the owned frames, encode them
        own_frames, local_fm = [], []
        for v in videos:
            fr = torch.stack([v["image"][f].to(dev, non_blocking=True) for f in own], dim=0)
            own_frames.append(fr)
        pres = [self._start_encoder(fr) if self.overlap_streams else None for fr in own_frames]
        for fr in own_frames:
            local_fm.append(trk.model.fnet_frames(fr))
        local = torch.stack(local_fm, dim=1)  # (n_own, C, H4, W4, 128): frame-major so one collective serves all clips
        # B. the exchange step
        full = sharding.allgather_frames(local, T)  # (T, C, H4, W4, 128)
        # C. chains: clip c on rank c % world
        results = [None] * C
        tv_local = []
        shapes = []
        for c, v in enumerate(videos):
            q = v["query_points"]
            M, P, _ = q.shape
            shapes.append((M, P))
            if c % world == rank:
                pyr = trk.model.build_pyramid(full[:, c].contiguous())
                traj, vis = trk.track_on_features(pyr, q.reshape(1, M * P, 3).to(dev))
                tv_local.append(torch.cat([traj[0], vis[0].float()[..., None]], dim=-1))  # (T, N, 3)
        del full
        # D. share the trajectories (tiny)
        N_max = max(m * p for m, p in shapes)
        mine = torch.zeros((len(range(rank, C, world)), T, N_max, 3), device=dev)
        for i, t in enumerate(tv_local):
            mine[i, :, : t.shape[1]] = t
        n_slots = (C + world - 1) // world
        slab = torch.zeros((n_slots, T, N_max, 3), device=dev)
        slab[: mine.shape[0]] = mine
        gathered = torch.empty((world * n_slots, T, N_max, 3), device=dev)
        dist.all_gather_into_tensor(gathered, slab)
        # E. SAM on the owned frames
        h, w = own_frames[0].shape[-2:]
        for c, v in enumerate(videos):
            M, P = shapes[c]
            tv = gathered[(c % world) * n_slots + c // world, :, : M * P]
            traj = tv[..., :2].reshape(T, M, P, 2)
            vis = tv[..., 2].reshape(T, M, P)
            out_code = float(PointVisibilityType.OUTSIDE_FRAME.value)
            oob = (traj[..., 0] / w < 0.01) | (traj[..., 1] / h < 0.01) | (traj[..., 0] / w > 0.99) | (traj[..., 1] / h > 0.99)
            vis = torch.where(oob, torch.full_like(vis, out_code), vis)
            logits, spf, _ = self._apply_sam_to_frames(own_frames[c], own, traj, vis, pre=pres[c])
            res = {"trajectories": traj, "visibilities": vis, "logits": logits, "frame_ids": own, "scores_per_frame": spf}
            if gather_logits:
                res["logits"] = sharding.allgather_frames(logits.transpose(0, 1).contiguous(), T).transpose(0, 1)
                res["frame_ids"] = list(range(T))
            results[c] = res
        return results

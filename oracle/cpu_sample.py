"""ORACLE-SIDE TIMING AID (test infrastructure, NOT product code): the bounded CPU calibration sample.

Only ``tests/`` (incl. ``tests/golden/make_golden_full.py``) and ``bench.py``'s cpu_baseline / ``--impl reference`` leg import this.

The CPU baseline of ``bench.py`` is the MEASURED full-clip run of the reference path recorded by
``tests/golden/make_golden_full.py`` (``tests/golden/<config>_full_cpu.json``: seconds per stage on the build container's cores).
A full clip is ~10 minutes of CPU, too long for a bench run, so on the GPU box ``bench.py`` re-times this small sample -- the
same three stage units that were timed next to the full run in the build container -- and rescales each cached stage by
``unit seconds here / unit seconds there``.  The sample is never the baseline by itself (no composition model, no window-count
guess): it only converts a measured full run from one host's cores to another's.

Units (each is the reference's own arithmetic for one work item of the stage):
* ``vit``     : `set_image` of one frame -- PIL-exact resize + preprocess + ViT (reference sam_pt/modeling/sam_pt.py:849)
* ``decode``  : `predict_mask` of one frame and one mask -- 1 + up to 12 `predict_torch` calls (sam_pt.py:760-837)
* ``tracker`` : the tracker's encoder on 2 frames at its working resolution (pips.py:191-287; 480x854 for PIPS, the interp
                resolution 384x512 for CoTracker) + for PIPS one 6-iteration `Pips.forward` update on an 8-frame window
                (pips.py:507-546), for CoTracker one 6-iteration UpdateFormer window (N = P + 4 support points)
"""
from __future__ import annotations

import time

import torch

SAM_SEED, PIPS_SEED = 7202, 7201


@torch.no_grad()
def bounded_sample(H: int, W: int, P: int, tracker: str = "pips", hq: bool = False, threads: int | None = None, vit: str = "vit_h"):
    """Returns {"vit": s, "decode": s, "tracker": s, "threads": n} (wall seconds of one unit of each stage)."""
    from sampt_b200 import synth
    from . import cotracker_ref, pips_ref, sam_ref, sampt_ref
    if threads:
        torch.set_num_threads(threads)
    cfg = {"vit_b": sam_ref.VIT_B, "vit_h": sam_ref.VIT_H, "vit_l": sam_ref.VIT_L}[vit]
    sam_sd = synth.condition_sam(synth.make_state_dict(sam_ref.sam_state_dict_shapes(cfg, hq=hq), SAM_SEED))
    clip = synth.make_clip(2, H, W)
    frames = clip["frames"]
    q = synth.make_query_points(clip, P)
    pred = sam_ref.RefSamPredictor(sam_sd, cfg, hq=hq)
    t0 = time.time()
    pred.set_image(frames[0].permute(1, 2, 0).numpy())
    t_vit = time.time() - t0
    t0 = time.time()
    traj = q[:, :, 1:][None]                                                  # (1, M=1, P, 2): frame 0's prompts
    sampt_ref.apply_sam_to_trajectories(pred, frames[:1], traj, torch.ones((1, 1, P)), positive_points_per_mask=P,
                                        sam_iou_threshold=-1e9, features_cache={0: {"features": pred.features, "interm": pred.interm}})
    t_dec = time.time() - t0
    if tracker == "pips":
        sd = synth.condition_pips(synth.make_state_dict(pips_ref.pips_state_dict_shapes(), PIPS_SEED))
        x = 2 * (frames.float() / 255.0) - 1.0
        t0 = time.time()
        fm = torch.cat([pips_ref.fnet(sd, x[i:i + 1]) for i in range(2)], dim=0)
        pips_ref.pips_forward(sd, q[0, :, 1:][None], None, None, 6, fmaps=fm[[0, 1, 1, 1, 1, 1, 1, 1]][None])
        t_trk = time.time() - t0
    else:
        sd = synth.condition_cotracker(synth.make_state_dict(cotracker_ref.cotracker_state_dict_shapes(), PIPS_SEED + 1), vis_bias=0.6, coord_scale=0.001)
        ih, iw = 384, 512
        r = torch.nn.functional.interpolate(frames.float(), (ih, iw), mode="bilinear")
        x = 2 * (r / 255.0) - 1.0
        t0 = time.time()
        fm = torch.cat([pips_ref.fnet(sd, x[i:i + 1], 4) for i in range(2)], dim=0)
        qq = q.clone()
        qq[:, :, 1] *= iw / W
        qq[:, :, 2] *= ih / H
        g = cotracker_ref.get_points_on_a_grid(2, (ih, iw))
        qq = torch.cat([qq, torch.cat([torch.zeros_like(g[:, :, :1]), g], dim=2)], dim=1)
        cotracker_ref.cotracker_forward(sd, r[[0, 1, 1, 1, 1, 1, 1, 1]][None], qq, iters=6, fmaps_all=fm[[0, 1, 1, 1, 1, 1, 1, 1]])
        t_trk = time.time() - t0
    return {"vit": t_vit, "decode": t_dec, "tracker": t_trk, "threads": torch.get_num_threads()}


def rescale_full_run(cached: dict, live_sample: dict) -> dict:
    """cached = tests/golden/<config>_full_cpu.json; live_sample = bounded_sample() on this host.
    Returns the full-clip seconds per stage rescaled to this host, and frames/s."""
    there = cached["calibration_sample"]
    s = cached["seconds"]
    est = {"tracker": s["tracker"] * live_sample["tracker"] / there["tracker"],
           "sam_set_image": s["sam_set_image"] * live_sample["vit"] / there["vit"],
           "sam_decode": s["sam_decode"] * live_sample["decode"] / there["decode"]}
    est["total"] = est["tracker"] + est["sam_set_image"] + est["sam_decode"]
    return {"seconds": est, "frames_per_s": cached["frames_run"] / est["total"]}

"""Full-clip golden fixtures + MEASURED CPU seconds for the BASELINE configs (build container only; minutes of CPU).

    python tests/golden/make_golden_full.py --config C2      # 50 x 480x854, SAM ViT-H + PIPS, 8 points   (BASELINE configs[1])
    python tests/golden/make_golden_full.py --config C3      # 50 x 480x854, SAM ViT-H + CoTracker, 64 points (configs[2])
    python tests/golden/make_golden_full.py --config C5s     # 8-frame slice of configs[4]: 1080x1920, HQ-SAM ViT-H + CoTracker, 256 pts

What runs (no subset, no extrapolation): the whole `SamPt.forward` path on every frame of the clip --
* PIPS tracker  = the UNMODIFIED reference `PipsPointTracker` imported from /root/reference (tests/golden/_refimport.py), i.e. the
  reference itself, with its per-window encoder recompute; the restatement `oracle/pips_ref.py` is run beside it and the
  difference recorded (`pips_restatement_max_diff_px`);
* CoTracker     = `oracle/cotracker_ref.py` (parity unpinned, see its header);
* SAM           = `oracle/sam_ref.py` (pinned to HF transformers), `SamPt` control flow = `oracle/sampt_ref.py`
  (reference sam_pt/modeling/sam_pt.py:694-866).
Wall seconds of every stage are recorded: they are the measured full-clip CPU baseline `bench.py` reports
(`cpu_baseline.sample = "full <config>, cached"`), together with the seconds of the bounded calibration sample
(`oracle/cpu_sample.py`) on the same cores, which `bench.py` re-times on the GPU box to rescale the cached full run to that host.

Writes tests/golden/<config>_full.npz (trajectories, visibilities, scores, bit-packed `logits > 0` per frame, refinement counts,
per-frame logit statistics) and tests/golden/<config>_full_cpu.json (seconds).  Inputs are NOT stored: they are re-generated from
the seeds (`sampt_b200.synth`).  ViT features are cached under /tmp (C2 and C3 share clip + SAM weights) -- not committed.
"""
import argparse
import hashlib
import json
import os
import platform
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "sam-pt_b200"))
sys.path.insert(0, ROOT)

from sampt_b200 import synth  # noqa: E402
from oracle import cotracker_ref, cpu_sample, pips_ref, sam_ref, sampt_ref  # noqa: E402

SAM_SEED, PIPS_SEED = 7202, 7201          # = bench.py / tests
COT_COORD_SCALE = 0.001   # synth.condition_cotracker: contractive over the 12 chained windows of a 50-frame clip
COT_VIS_BIAS = 0.6                        # synth.condition_cotracker: ~90 % of the C3 / C5 query points visible
CONFIGS = {
    # name: T, H, W, P, tracker, hq, clip seed, frames kept (slice of the clip the oracle runs on)
    "C2": dict(T=50, H=480, W=854, P=8, tracker="pips", hq=False, seed=72, keep=50),
    "C3": dict(T=50, H=480, W=854, P=64, tracker="cotracker", hq=False, seed=72, keep=50),
    "C5s": dict(T=100, H=1080, W=1920, P=256, tracker="cotracker", hq=True, seed=72, keep=8),
}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return platform.processor()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--feat-cache", default="/tmp/sampt_feat_cache")
    args = ap.parse_args()
    c = CONFIGS[args.config]
    torch.set_num_threads(args.threads)
    torch.manual_seed(0)
    cfg = sam_ref.VIT_H
    sam_sd = synth.condition_sam(synth.make_state_dict(sam_ref.sam_state_dict_shapes(cfg, hq=c["hq"]), SAM_SEED))
    video = synth.make_video_dict(c["T"], c["H"], c["W"], c["P"], seed=c["seed"])
    video["image"] = video["image"][:c["keep"]]
    images = torch.stack(video["image"])
    T = images.shape[0]
    secs = {}

    # ------------------------------------------------------------------ tracker
    extra = {}
    if c["tracker"] == "pips":
        from _refimport import import_reference_pips
        pips_sd = synth.condition_pips(synth.make_state_dict(pips_ref.pips_state_dict_shapes(), PIPS_SEED))
        R = import_reference_pips()
        with tempfile.TemporaryDirectory() as d:
            synth.write_pips_checkpoint_dir(pips_sd, d)
            trk = R["PipsPointTracker"](checkpoint_path=d, stride=4, s=8).eval()
            t0 = time.time()
            with torch.no_grad():
                traj_ref, vis_ref = trk(images[None], video["query_points"].reshape(1, -1, 3))
            secs["tracker"] = time.time() - t0
        t0 = time.time()
        traj_r, vis_r = pips_ref.pips_tracker_forward(pips_sd, images[None], video["query_points"].reshape(1, -1, 3))
        secs["tracker_restatement_not_counted"] = time.time() - t0
        extra["pips_restatement_max_diff_px"] = float((traj_r - traj_ref).abs().max())
        extra["pips_restatement_vis_equal"] = bool(torch.equal(vis_r, vis_ref))
        print("PIPS reference vs restatement:", extra, flush=True)
        tracker = lambda im, q: (traj_ref, vis_ref)          # noqa: E731  (the reference's own output feeds SAM)
        pips_sd_arg = None
    else:
        cot_sd = synth.condition_cotracker(synth.make_state_dict(cotracker_ref.cotracker_state_dict_shapes(), PIPS_SEED + 1), vis_bias=COT_VIS_BIAS, coord_scale=COT_COORD_SCALE)
        box = {}

        def tracker(im, q):
            t0 = time.time()
            raw = {}
            out = cotracker_ref.cotracker_point_tracker_forward(cot_sd, im, q, raw=raw)
            box["t"] = box.get("t", 0.0) + time.time() - t0
            box.setdefault("vis_sigmoid", []).append(raw["vis_sigmoid"][0])      # (T, m*P) per mask batch
            return out
        pips_sd_arg = None

    # ------------------------------------------------------------------ SAM (timed: set_image = resize + ViT; the rest = decode)
    pred = sam_ref.RefSamPredictor(sam_sd, cfg, hq=c["hq"])
    key = hashlib.sha1(f"{c['H']}x{c['W']}-{c['seed']}-{SAM_SEED}-{int(c['hq'])}-{c['T']}".encode()).hexdigest()[:12]
    cdir = os.path.join(args.feat_cache, key)
    os.makedirs(cdir, exist_ok=True)
    fcache = {}
    vit_secs_path = os.path.join(cdir, "vit_secs.json")
    vit_secs = json.load(open(vit_secs_path)) if os.path.exists(vit_secs_path) else {}
    for f in range(T):
        p = os.path.join(cdir, f"f{f:03d}.pt")
        if os.path.exists(p) and str(f) in vit_secs:
            fcache[f] = torch.load(p)
        else:
            t0 = time.time()
            pred.set_image(images[f].permute(1, 2, 0).numpy())
            vit_secs[str(f)] = time.time() - t0
            fcache[f] = {"features": pred.features, "interm": pred.interm}
            torch.save(fcache[f], p)
            json.dump(vit_secs, open(vit_secs_path, "w"))
            print(f"ViT frame {f}: {vit_secs[str(f)]:.1f}s", flush=True)
    secs["sam_set_image"] = float(sum(vit_secs[str(f)] for f in range(T)))
    taps = {}
    t0 = time.time()
    ref = sampt_ref.sampt_forward(pips_sd_arg, pred, video, positive_points_per_mask=c["P"], sam_iou_threshold=-1e9,
                                  features_cache=fcache, taps=taps, tracker=tracker)
    t_all = time.time() - t0
    if c["tracker"] == "cotracker":
        secs["tracker"] = box["t"]
        secs["sam_decode"] = t_all - box["t"]
    else:
        secs["sam_decode"] = t_all
    secs["total"] = secs["tracker"] + secs["sam_set_image"] + secs["sam_decode"]

    # ------------------------------------------------------------------ calibration sample on the same cores
    sample = cpu_sample.bounded_sample(c["H"], c["W"], c["P"], tracker=c["tracker"], hq=c["hq"], threads=args.threads)

    lg = torch.stack(ref["logits"])                                      # (M,T,H,W)
    bits = np.packbits((lg > 0).numpy().reshape(lg.shape[0], T, -1), axis=-1)
    a = lg.abs().reshape(lg.shape[0], T, -1)
    stats = torch.stack([a.median(dim=-1).values, a.amax(dim=-1), (a < 1e-2).float().mean(dim=-1), (a < 1e-3).float().mean(dim=-1),
                         (lg > 0).reshape(lg.shape[0], T, -1).float().mean(dim=-1)], dim=-1).numpy()
    n_ref = np.array([[taps["n_refine"].get((f, m), -1) for m in range(lg.shape[0])] for f in range(T)], dtype=np.int32)
    out = os.path.join(HERE, f"{args.config}_full.npz")
    np.savez_compressed(out, trajectories=ref["trajectories"].numpy(), visibilities=ref["visibilities"].numpy(),
                        scores_per_frame=np.array(ref["scores_per_frame"], dtype=np.float32), mask_bits=bits,
                        logit_stats=stats.astype(np.float32), n_refine=n_ref, query_points=video["query_points"].numpy(),
                        hw=np.array([c["H"], c["W"]]),
                        vis_sigmoid=(torch.cat(box["vis_sigmoid"], dim=1).reshape(T, -1, c["P"]).numpy()
                                     if c["tracker"] == "cotracker" else np.zeros((0,), dtype=np.float32)))
    meta = {"config": args.config, **{k: c[k] for k in ("T", "H", "W", "P", "tracker", "hq", "seed", "keep")},
            "frames_run": T, "seconds": secs, "frames_per_s": T / secs["total"], "threads": args.threads, "cpu": cpu_model(),
            "calibration_sample": sample, "sam_seed": SAM_SEED, "tracker_seed": PIPS_SEED if c["tracker"] == "pips" else PIPS_SEED + 1,
            "logit_stats_columns": ["median|l|", "max|l|", "frac |l|<1e-2", "frac |l|<1e-3", "frac l>0"], **extra,
            "what": "whole SamPt.forward on every frame; tracker = unmodified reference PipsPointTracker (C2) or the CoTracker "
                    "restatement; SAM = oracle restatement; seconds are wall clock on `threads` torch threads"}
    json.dump(meta, open(os.path.join(HERE, f"{args.config}_full_cpu.json"), "w"), indent=1)
    print("wrote", out, os.path.getsize(out), "bytes;", json.dumps(meta["seconds"]), f"{meta['frames_per_s']:.4f} frames/s")


if __name__ == "__main__":
    main()

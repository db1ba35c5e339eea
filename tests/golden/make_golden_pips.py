"""Generate golden vectors for the PIPS path by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_pips.py

Imports /root/reference's Pips / PipsPointTracker (see _refimport.py), runs them on CPU on seeded synthetic
inputs (sampt_b200.synth) and writes small fixtures next to this file.  /root/reference does not exist on the
GPU box, so tests only ever read the fixtures.  Inputs are NOT stored: they are re-generated from the seeds.
"""
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "sam-pt_b200"))
sys.path.insert(0, ROOT)

from _refimport import import_reference_pips  # noqa: E402
from sampt_b200 import synth  # noqa: E402
from oracle import pips_ref  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)

PIPS_SEED = 7201


def pips_sd():
    return synth.condition_pips(synth.make_state_dict(pips_ref.pips_state_dict_shapes(), PIPS_SEED))


def main():
    R = import_reference_pips()
    sd = pips_sd()
    model = R["Pips"](S=8, stride=4).eval()
    missing = model.load_state_dict(sd, strict=True)
    out = {}

    # ---- (1) one Pips.forward window: 8 frames 96x128, 5 points -------------------------------------------------
    clip = synth.make_clip(8, 96, 128, seed=72)
    rgbs = clip["frames"][None].float()
    xys = synth.make_query_points(clip, 5, seed=72)[:, :, 1:]
    with torch.no_grad():
        preds, _, vis_e, ffeat, _ = model(xys, rgbs, iters=6, return_feat=True)
        fmaps = model.fnet(2 * (rgbs[0] / 255.0) - 1.0)
        preds2, _, vis_e2, _, _ = model(xys + 3.0, rgbs, feat_init=ffeat, iters=6, return_feat=True)
    out["window"] = {
        "cfg": {"T": 8, "H": 96, "W": 128, "P": 5, "seed": 72},
        "coords_per_iter": torch.stack(preds)[:, 0].clone(),  # (6,8,N,2)
        "vis_e": vis_e[0].clone(), "ffeat": ffeat[0].clone(),
        "fmaps_sub": fmaps[:, ::16, ::3, ::4].clone(),  # subsample keeps the fixture small
        "fmaps_mean": fmaps.mean(dim=(2, 3)).clone(), "fmaps_abs_mean": fmaps.abs().mean().clone(),
        "coords_feat_init": preds2[-1][0].clone(), "vis_e_feat_init": vis_e2[0].clone(),
    }

    # ---- (2) full tracker: 14 frames 96x128, points born at t=0, 5, 13 (exercises both directions + padding) ----
    T, H, W = 14, 96, 128
    clip = synth.make_clip(T, H, W, seed=73)
    q = synth.make_query_points(clip, 6, seed=73)
    q[0, 2:4, 0] = 5.0
    q[0, 2:4, 1:] = synth.make_query_points(clip, 2, seed=74, t=5)[0, :, 1:]
    q[0, 5, 0] = 13.0
    with tempfile.TemporaryDirectory() as d:
        synth.write_pips_checkpoint_dir(sd, d)
        trk = R["PipsPointTracker"](checkpoint_path=d, stride=4, s=8).eval()
        with torch.no_grad():
            traj, vis = trk(clip["frames"][None], q)
    out["tracker"] = {"cfg": {"T": T, "H": H, "W": W, "seed": 73}, "query_points": q.clone(),
                      "trajectories": traj.clone(), "visibilities": vis.clone()}

    # ---- (3) C1-shaped tracker run: 2 frames 240x320, 4 points -------------------------------------------------
    clip = synth.make_clip(2, 240, 320, seed=72)
    q = synth.make_query_points(clip, 4, seed=72)
    with tempfile.TemporaryDirectory() as d:
        synth.write_pips_checkpoint_dir(sd, d)
        trk = R["PipsPointTracker"](checkpoint_path=d, stride=4, s=8).eval()
        with torch.no_grad():
            traj, vis = trk(clip["frames"][None], q)
    out["c1_tracker"] = {"cfg": {"T": 2, "H": 240, "W": 320, "P": 4, "seed": 72}, "query_points": q.clone(),
                         "trajectories": traj.clone(), "visibilities": vis.clone()}

    path = os.path.join(HERE, "pips_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

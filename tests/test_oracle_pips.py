"""Pins oracle/pips_ref.py (the CPU restatement) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden_pips.py, run in the build container where /root/reference exists)."""
import os

import pytest
import torch

from oracle import pips_ref
from sampt_b200 import synth

PIPS_SEED = 7201


@pytest.fixture(scope="module")
def gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "pips_golden.pt"))


@pytest.fixture(scope="module")
def sd():
    return synth.condition_pips(synth.make_state_dict(pips_ref.pips_state_dict_shapes(), PIPS_SEED))


def test_pips_window_matches_reference(gold, sd):
    g = gold["window"]
    c = g["cfg"]
    clip = synth.make_clip(c["T"], c["H"], c["W"], seed=c["seed"])
    rgbs = clip["frames"][None].float()
    xys = synth.make_query_points(clip, c["P"], seed=c["seed"])[:, :, 1:]
    x = 2 * (rgbs[0] / 255.0) - 1.0
    fm = pips_ref.fnet(sd, x)
    assert torch.allclose(fm[:, ::16, ::3, ::4], g["fmaps_sub"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(fm.mean(dim=(2, 3)), g["fmaps_mean"], atol=1e-5)
    preds, vis_e, ffeat = pips_ref.pips_forward(sd, xys, rgbs, None, 6)
    got = torch.stack(preds)[:, 0]
    assert (got - g["coords_per_iter"]).abs().max() < 2e-4  # px
    assert torch.allclose(vis_e[0], g["vis_e"], atol=1e-4)
    assert torch.allclose(ffeat[0], g["ffeat"], atol=1e-5)
    preds2, vis_e2, _ = pips_ref.pips_forward(sd, xys + 3.0, rgbs, ffeat, 6)
    assert (preds2[-1][0] - g["coords_feat_init"]).abs().max() < 2e-4
    assert torch.allclose(vis_e2[0], g["vis_e_feat_init"], atol=1e-4)


@pytest.mark.parametrize("share_fnet", [False, True])
def test_pips_tracker_matches_reference(gold, sd, share_fnet):
    g = gold["tracker"]
    c = g["cfg"]
    clip = synth.make_clip(c["T"], c["H"], c["W"], seed=c["seed"])
    traj, vis = pips_ref.pips_tracker_forward(sd, clip["frames"][None], g["query_points"], share_fnet=share_fnet)
    assert traj.shape == g["trajectories"].shape
    assert (traj - g["trajectories"]).abs().max() < 1e-3  # px (north-star tolerance)
    assert torch.equal(vis, g["visibilities"])


def test_pips_tracker_c1(gold, sd):
    g = gold["c1_tracker"]
    c = g["cfg"]
    clip = synth.make_clip(c["T"], c["H"], c["W"], seed=c["seed"])
    traj, vis = pips_ref.pips_tracker_forward(sd, clip["frames"][None], g["query_points"])
    assert (traj - g["trajectories"]).abs().max() < 1e-3
    assert torch.equal(vis, g["visibilities"])

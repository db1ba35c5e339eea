"""CPU test of the HOST side of the product CoTracker (`CoTracker.track`: sorting by birth frame, growing active set, window
carry-over, tail padding, frame-order maps for the time-reversed pass and short clips).  The two native entry points it calls
are replaced by oracle-backed fakes, so this checks the bookkeeping — the part the GPU parity tests cover with only a few
configurations — over many randomised clips.  (The product itself never imports the oracle; the substitution happens here.)"""
import random
from types import SimpleNamespace

import pytest
import torch

from oracle import cotracker_ref as R
from sampt_b200 import synth


class _FakeLib:
    """oracle-backed stand-ins for sampt_cotracker_sample_features / sampt_cotracker_window (same argument order)."""

    def __init__(self, sd):
        self.sd = sd
        self.windows = 0

    def sampt_cotracker_sample_features(self, _h, fmaps, H4, W4, frame, xy, n, S, out, _st):
        from oracle import pips_ref
        for i in range(n.value):
            fm = fmaps[int(frame[i])].permute(2, 0, 1)[None]                      # (1,128,H4,W4)
            f = pips_ref.bilinear_sample2d(fm, xy[i:i + 1, 0][None], xy[i:i + 1, 1][None])[0, :, 0]
            out[i] = f[None, :].repeat(S.value, 1)
        return 0

    def sampt_cotracker_window(self, _h, p0, p1, p2, p3, H4, W4, fidx, coords, ffeats, tm, vi, _temb, n, iters, td, sdp, vis, _st):
        N = n.value
        assert coords.shape == (N, 8, 2) and ffeats.shape == (N, 8, 128) and tm.shape == (N, 8) and vi.shape == (N, 8)
        assert int(fidx[0]) == 0 and int(fidx[1]) == 0
        fm = p0[fidx[2:].long()].permute(0, 3, 1, 2)[None]                        # (1,8,128,H4,W4)
        preds, v = R.forward_iteration(self.sd, fm, coords.permute(1, 0, 2)[None].clone(), ffeats.permute(1, 0, 2)[None].clone(),
                                       vi.t()[None, :, :, None].clone(), (tm.t()[None, :, :, None] > 0.5), iters.value)
        coords.copy_(preds[-1][0].permute(1, 0, 2) / 4.0)
        vis.copy_(v[0].t())
        self.windows += 1
        return 0


@pytest.fixture()
def harness(monkeypatch):
    from sam_pt.point_tracker.cotracker import cotracker as C
    sd = synth.condition_cotracker(synth.make_state_dict(R.cotracker_state_dict_shapes(), seed=31))
    fake = _FakeLib(sd)
    monkeypatch.setattr(C.native, "lib", lambda: fake)
    monkeypatch.setattr(C.native, "ptr", lambda t: t)
    monkeypatch.setattr(C.native, "stream_ptr", lambda: None)
    monkeypatch.setattr(C.native, "check", lambda rc, what="": None)
    model = C.CoTracker()
    model._time_emb = None
    monkeypatch.setattr(model, "native_context", lambda: SimpleNamespace(handle=None))
    return model, fake, sd


def _features(T, H4=16, W4=24, seed=0):  # >= 16 so that the coarsest pyramid level is at least 2x2
    g = torch.Generator().manual_seed(seed)
    return torch.randn((T, H4, W4, 128), generator=g)


def _run_product(model, fm, q, order, iters):
    pyr = [fm, fm[:, ::2, ::2], fm[:, ::4, ::4], fm[:, ::8, ::8]]                 # only level 0 is read by the fakes
    return model.track(pyr, q, order, iters=iters)


def _run_oracle(sd, fm, q, order, iters):
    fm_o = fm[list(order)].permute(0, 3, 1, 2)                                     # the clip as the reference would see it
    T = len(order)
    traj, vis = R.cotracker_forward(sd, torch.zeros((1, T, 3, fm.shape[1] * 4, fm.shape[2] * 4)), q[None], iters=iters, fmaps_all=fm_o)
    return traj[0], vis[0]


@pytest.mark.parametrize("case", range(6))
def test_window_bookkeeping_matches_oracle(harness, case):
    model, fake, sd = harness
    rng = random.Random(100 + case)
    T = rng.choice([8, 9, 12, 13, 16, 21])
    N = rng.randint(1, 6)
    fm = _features(T, seed=case)
    births = [rng.randrange(0, T) for _ in range(N)]
    if case % 2 == 0:
        births[0] = 0
    q = torch.tensor([[float(b), rng.uniform(4, 90), rng.uniform(4, 60)] for b in births])
    order = list(range(T)) if case % 3 else list(range(T - 1, -1, -1))             # every third case: the time-reversed pass
    traj, vis = _run_product(model, fm, q, order, iters=2)
    traj_o, vis_o = _run_oracle(sd, fm, q, order, iters=2)
    assert traj.shape == (T, N, 2) and vis.shape == (T, N)
    assert (traj - traj_o).abs().max().item() < 1e-4, (T, births)
    assert (vis - vis_o).abs().max().item() < 1e-5
    assert fake.windows >= 1


def test_short_clip_order_map_equals_padded_clip(harness):
    """CoTrackerForShortVideosWrapper repeats the last frame up to 8 frames; the product expresses that (and the flipped
    variant) as a frame-order map over the un-padded features."""
    model, fake, sd = harness
    T = 5
    fm = _features(T, seed=9)
    q = torch.tensor([[0.0, 20.0, 20.0], [3.0, 40.0, 30.0], [4.0, 50.0, 12.0]])
    for order in (list(range(T)) + [T - 1] * 3, list(range(T - 1, -1, -1)) + [0] * 3):
        traj, vis = _run_product(model, fm, q if order[0] == 0 else torch.cat([T - 1 - q[:, :1], q[:, 1:]], 1), order, iters=2)
        traj_o, vis_o = _run_oracle(sd, fm, q if order[0] == 0 else torch.cat([T - 1 - q[:, :1], q[:, 1:]], 1), order, iters=2)
        assert traj.shape == (8, 3, 2)
        assert (traj - traj_o).abs().max().item() < 1e-4
        assert (vis - vis_o).abs().max().item() < 1e-5


def test_no_point_in_early_windows_is_skipped(harness):
    model, fake, sd = harness
    T = 20
    fm = _features(T, seed=3)
    q = torch.tensor([[14.0, 30.0, 20.0], [17.0, 12.0, 33.0]])
    traj, vis = _run_product(model, fm, q, list(range(T)), iters=1)
    traj_o, vis_o = _run_oracle(sd, fm, q, list(range(T)), iters=1)
    assert (traj[:8] == 0).all()                       # windows starting at 0 and 4 contain no point yet
    assert (traj - traj_o).abs().max().item() < 1e-4 and (vis - vis_o).abs().max().item() < 1e-5

"""CPU checks of the CoTracker restatement (oracle/cotracker_ref.py, PARITY UNPINNED) and of the host-side tables the product
shares with it.  With no upstream goldens available these are structural: shapes, parameter count (SURVEY Appendix B.4:
about 24 M parameters), the semantics the reference wrapper relies on, and permutation equivariance of the joint tracker."""
import math

import torch

from oracle import cotracker_ref as R
from sampt_b200 import synth


def _sd(seed=31):
    return synth.condition_cotracker(synth.make_state_dict(R.cotracker_state_dict_shapes(), seed=seed))


def test_state_dict_table_matches_product_and_param_count():
    from sam_pt.point_tracker.cotracker.cotracker import cotracker_shapes
    a, b = cotracker_shapes(), R.cotracker_state_dict_shapes()
    assert a == b
    n = sum(math.prod(v) for v in b.values())
    assert 23.5e6 < n < 25e6, n
    # LayerNorms inside the UpdateFormer carry no parameters (elementwise_affine=False)
    assert not any(".norm1." in k or ".norm2." in k for k in b)


def test_host_tables_match():
    from sam_pt.point_tracker.cotracker.cotracker import _time_embed_table
    from sam_pt.point_tracker.cotracker.tracker import get_points_on_a_grid
    assert torch.equal(_time_embed_table(456, 8), R.time_embed(456, 8))
    for g in (1, 2, 5):
        assert torch.allclose(get_points_on_a_grid(g, (384, 512)), R.get_points_on_a_grid(g, (384, 512)))
    g2 = R.get_points_on_a_grid(2, (384, 512))[0]
    assert g2.min().item() == 8.0 and g2[:, 0].max().item() == 504.0 and g2[:, 1].max().item() == 376.0


def test_pos_embed_sampling_hits_table_at_integer_coords():
    tab = torch.from_numpy(R.get_2d_sincos_pos_embed(456, (6, 9))).float().reshape(6, 9, 456)
    c = torch.tensor([[[4.0, 2.0], [0.0, 5.0], [8.0, 0.0]]])
    s = R.sample_pos_embed((6, 9), 456, c)
    assert torch.allclose(s[0, 0], tab[2, 4]) and torch.allclose(s[0, 1], tab[5, 0]) and torch.allclose(s[0, 2], tab[0, 8])
    # first half of the channels depends on x only, second half on y only
    assert torch.allclose(tab[1, 3, :228], tab[4, 3, :228]) and torch.allclose(tab[2, 1, 228:], tab[2, 7, 228:])


def test_windows_birth_frames_and_zero_before_first_window():
    sd = _sd()
    T, H, W = 14, 64, 96
    frames = synth.make_clip(T, H, W, seed=5)["frames"].float()
    q = torch.tensor([[[0.0, 30.0, 20.0], [9.0, 60.0, 40.0], [0.0, 50.0, 30.0]]])
    traj, vis = R.cotracker_forward(sd, frames[None], q, iters=2)
    assert traj.shape == (1, T, 3, 2) and vis.shape == (1, T, 3)
    # the point born at t=9 joins with the window starting at frame 4: exact zeros (vis = sigmoid(0)) before it
    assert (traj[0, :4, 1] == 0).all() and (vis[0, :4, 1] == 0.5).all() and (traj[0, 4:, 1] != 0).all()
    assert torch.isfinite(traj).all()


def test_joint_tracker_is_permutation_equivariant():
    sd = _sd()
    T, H, W = 8, 64, 96
    frames = synth.make_clip(T, H, W, seed=7)["frames"].float()
    q = torch.tensor([[[0.0, 30.0, 20.0], [0.0, 60.0, 40.0], [0.0, 50.0, 30.0], [0.0, 12.0, 50.0]]])
    perm = torch.tensor([2, 0, 3, 1])
    t1, v1 = R.cotracker_forward(sd, frames[None], q, iters=3)
    t2, v2 = R.cotracker_forward(sd, frames[None], q[:, perm], iters=3)
    assert (t1[:, :, perm] - t2).abs().max().item() < 1e-3
    assert (v1[:, :, perm] - v2).abs().max().item() < 1e-4
    # ... and NOT independent: removing a point changes the others (space attention), unlike PIPS
    t3, _ = R.cotracker_forward(sd, frames[None], q[:, :3], iters=3)
    assert (t3 - t1[:, :, :3]).abs().max().item() > 1e-6


def test_wrapper_short_clip_and_backward_merge():
    sd = _sd()
    T, H, W = 5, 48, 64
    frames = synth.make_clip(T, H, W, seed=3)["frames"]
    q = torch.tensor([[[0.0, 20.0, 20.0], [3.0, 40.0, 30.0]]])
    traj, vis = R.cotracker_point_tracker_forward(sd, frames[None], q, interp_shape=(48, 64))
    assert traj.shape == (1, T, 2, 2) and vis.shape == (1, T, 2) and vis.dtype == torch.bool
    assert (traj != 0).all()  # the time-reversed pass fills every frame

"""GPU parity of the SURVEY §8 f3 row: patch-similarity filtering against the restated reference arithmetic, and the point
re-initialisation control flow against the restated reference loop driven by the product's own primitives (see
oracle/reinit_ref.py for why that is the meaningful comparison)."""
import pytest
import torch

from oracle import pips_ref, reinit_ref, sam_ref
from sampt_b200 import factory, synth

pytestmark = pytest.mark.gpu


def _model(tmp_path, **kw):
    cfg = sam_ref.VIT_TEST
    sam_sd = synth.condition_sam(synth.make_state_dict(sam_ref.sam_state_dict_shapes(cfg), 31))
    pips_sd = synth.condition_pips(synth.make_state_dict(pips_ref.pips_state_dict_shapes(), 7201))
    ckpt = synth.write_pips_checkpoint_dir(pips_sd, str(tmp_path / "pips"))
    m = factory.build_sam_pt("vit_test", sam_sd, ckpt, positive_points_per_mask=4, sam_iou_threshold=-1e9)
    for k, v in kw.items():
        setattr(m, k, v)
    return m


@pytest.mark.parametrize("ps,thr", [(3, 0.01), (5, 0.05), (3, 0.6)])
def test_patch_similarity_matches_reference_arithmetic(tmp_path, ps, thr):
    m = _model(tmp_path, patch_size=ps, patch_similarity_threshold=thr)
    clip = synth.make_clip(9, 96, 128, seed=21)
    g = torch.Generator().manual_seed(4)
    N = 7
    q = torch.cat([torch.randint(0, 9, (N, 1), generator=g).float(), torch.rand((N, 2), generator=g) * torch.tensor([127.0, 95.0])], dim=1)
    q[0, 1:] = torch.tensor([0.2, 0.4])                     # patch partly outside the frame (zero padding)
    traj = torch.rand((9, N, 2), generator=g) * torch.tensor([127.0, 95.0])
    traj[3, 1] = torch.tensor([126.9, 94.8])
    vis = (torch.rand((9, N), generator=g) > 0.2).float()
    sim_ref, vis_ref = reinit_ref.patch_similarity(clip["frames"], q, traj, vis, ps, thr)
    for _ in range(8):                                      # a similarity within float noise of the threshold is not a comparable
        if not ((sim_ref - thr).abs() < 1e-4).any():        # decision (and everything after it cascades): nudge the threshold off it
            break
        thr *= 1.03
        sim_ref, vis_ref = reinit_ref.patch_similarity(clip["frames"], q, traj, vis, ps, thr)
    m.patch_similarity_threshold = thr
    vis_gpu = m._patch_filter(clip["frames"].cuda(), q.cuda(), traj.cuda(), vis.cuda())
    sim_gpu = m._last_patch_similarities.cpu()
    assert (sim_gpu - sim_ref).abs().max() < 2e-5, (sim_gpu - sim_ref).abs().max()
    assert not ((sim_ref - thr).abs() < 1e-4).any()
    assert torch.equal(vis_gpu.cpu(), vis_ref)
    assert (vis_ref == reinit_ref.PATCH_NON_SIMILAR).any() or thr < 0.05


@pytest.mark.parametrize("variant", ["reinit-at-median-of-area-diff", "reinit-on-horizon-and-sync-masks", "reinit-on-similar-mask-area",
                                     "reinit-on-similar-mask-area-and-sync-masks"])
def test_reinit_control_flow_matches_reference_loop(tmp_path, variant):
    m = _model(tmp_path, use_point_reinit=True, reinit_horizon=4, reinit_point_tracker_horizon=6, reinit_variant=variant)
    video = synth.make_video_dict(11, 96, 128, 4, seed=33)
    q0 = video["query_points"]
    q1 = q0.clone()
    q1[..., 1] = (q1[..., 1] + 31.0) % 120.0 + 4.0
    q1[..., 0] = 2.0                                        # second mask annotated on frame 2: exercises the flipped pass + syncing
    qp = torch.cat([q0, q1], dim=0).cuda()
    images = torch.stack(video["image"]).cuda()
    torch.manual_seed(5)
    got = m._forward_w_reinit(images, qp)
    torch.manual_seed(5)
    ref = reinit_ref.forward_w_reinit(
        images, qp, track_points=m._track_points, apply_sam=m._apply_sam_to_trajectories,
        extract_query_points=m.extract_query_points, reinit_horizon=4, reinit_point_tracker_horizon=6, reinit_variant=variant)
    for a, b, name in zip(got, ref, ("trajectories", "visibilities", "logits", "scores", "scores_per_frame")):
        assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)), name
    # and through the public entry point
    torch.manual_seed(5)
    video["query_points"] = qp.cpu()
    out = m(video)
    assert torch.equal(out["trajectories"], got[0].cpu())
    assert len(out["logits"]) == 2 and out["logits"][0].shape == (11, 96, 128)

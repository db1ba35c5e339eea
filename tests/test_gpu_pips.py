"""GPU parity: libsampt_b200's PIPS path (through the C ABI / drop-in classes) against the pinned CPU oracle and the
golden vectors the unmodified reference produced.  Tolerances: coords 1e-3 px (north star), visibilities exact."""
import os

import pytest
import torch

from oracle import pips_ref
from sampt_b200 import synth

pytestmark = pytest.mark.gpu
PIPS_SEED = 7201


@pytest.fixture(scope="module")
def sd():
    return synth.condition_pips(synth.make_state_dict(pips_ref.pips_state_dict_shapes(), PIPS_SEED))


@pytest.fixture(scope="module")
def model(sd):
    from sam_pt.point_tracker.pips import Pips
    m = Pips(S=8, stride=4)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "pips_golden.pt"))


def test_linear_f32_matches_torch():
    from ctypes import c_int
    from sampt_b200 import native
    ctx = native.get_context("cuda")
    g = torch.Generator().manual_seed(0)
    # (33..64 rows: the clustered split-K kernel incl. K / N tails and every cluster size; others: small-M and pipelined kernels)
    for (M, N, K, act) in [(64, 512, 520, 0), (64, 2048, 512, 1), (64, 512, 2048, 0), (5, 1040, 512, 0), (4096, 128, 256, 2),
                           (300, 70, 36, 1), (33, 1040, 512, 3), (48, 70, 36, 1), (64, 4736, 64, 0), (40, 96, 2048, 2),
                           (64, 512, 4096, 0), (263, 2048, 256, 2), (500, 130, 384, 0), (129, 256, 2048, 1)]:
        x = torch.randn((M, K), generator=g).cuda()
        w = (torch.randn((N, K), generator=g) / K ** 0.5).cuda()
        b = torch.randn((N,), generator=g).cuda()
        r = torch.randn((M, N), generator=g).cuda()
        y = torch.empty((M, N), device="cuda")
        native.check(native.lib().sampt_linear_f32(ctx.handle, native.ptr(x), c_int(K), native.ptr(w), c_int(K), native.ptr(b),
                                                   native.ptr(r), c_int(N), native.ptr(y), c_int(N), c_int(M), c_int(N),
                                                   c_int(K), c_int(act), native.stream_ptr()))
        ref = torch.nn.functional.linear(x.cpu().double(), w.cpu().double(), b.cpu().double())
        if act == 1:
            ref = torch.nn.functional.gelu(ref)
        elif act == 2:
            ref = torch.relu(ref)
        elif act == 3:
            ref = torch.nn.functional.gelu(ref, approximate="tanh")
        ref = ref + r.cpu().double()
        assert (y.cpu().double() - ref).abs().max() < 2e-5, (M, N, K, act)


def test_fnet_and_pyramid_match_oracle(sd, model):
    clip = synth.make_clip(3, 96, 128, seed=5)
    frames = clip["frames"]
    x = 2 * (frames.float() / 255.0) - 1.0
    ref = pips_ref.fnet(sd, x)  # (T,128,h,w)
    pyr = model.encode_frames(frames.cuda())
    got = pyr[0].permute(0, 3, 1, 2).cpu()
    scale = ref.abs().max().item()
    assert (got - ref).abs().max() < 2e-5 * max(scale, 1.0)
    ref_pyr = pips_ref.build_pyramid(ref[None])
    for l in range(1, 4):
        assert torch.allclose(pyr[l].permute(0, 3, 1, 2).cpu(), ref_pyr[l][0], atol=2e-5)


def test_fnet_odd_size(sd, model):
    """480x854-like odd intermediate sizes (427, 214, 107, 54 columns) at a smaller scale: 60x106."""
    clip = synth.make_clip(2, 60, 106, seed=6)
    ref = pips_ref.fnet(sd, 2 * (clip["frames"].float() / 255.0) - 1.0)
    got = model.encode_frames(clip["frames"].cuda())[0].permute(0, 3, 1, 2).cpu()
    assert (got - ref).abs().max() < 2e-5 * max(ref.abs().max().item(), 1.0)


def test_corr_lookup_matches_oracle(model):
    """The fused gather-dot against CorrBlock.corr+sample (dense formulation), incl. out-of-map points."""
    from ctypes import c_int
    from sampt_b200 import native
    g = torch.Generator().manual_seed(1)
    S, N, H, W = 8, 6, 24, 32
    fm = torch.randn((1, S, 128, H, W), generator=g)
    pyr = pips_ref.build_pyramid(fm)
    ffeats = torch.randn((1, S, N, 128), generator=g)
    coords = torch.rand((1, S, N, 2), generator=g) * torch.tensor([W - 1.0, H - 1.0])
    coords[0, :, 0] = torch.tensor([-5.5, 0.25])  # outside top-left
    coords[0, :, 1] = torch.tensor([W + 2.0, H - 0.5])  # outside right
    coords[0, :, 2] = torch.tensor([7.0, 9.0])  # exactly integral
    ref = pips_ref.corr_lookup(pyr, ffeats, coords)[0]  # (S,N,196)
    ctx = native.get_context("cuda")
    lv = [p[0].permute(0, 2, 3, 1).contiguous().cuda() for p in pyr]
    ff = ffeats[0].permute(1, 0, 2).contiguous().cuda()
    cc = coords[0].permute(1, 0, 2).contiguous().cuda()
    out = torch.empty((N, S, 196), device="cuda")
    native.check(native.lib().sampt_pips_corr_lookup(ctx.handle, native.ptr(lv[0]), native.ptr(lv[1]), native.ptr(lv[2]),
                                                     native.ptr(lv[3]), c_int(S), c_int(H), c_int(W), native.ptr(ff),
                                                     native.ptr(cc), c_int(N), native.ptr(out), native.stream_ptr()))
    got = out.permute(1, 0, 2).cpu()
    assert (got - ref).abs().max() < 1e-4 * max(1.0, ref.abs().max().item())


def test_single_window_matches_reference_golden(gold, model):
    g = gold["window"]
    c = g["cfg"]
    clip = synth.make_clip(c["T"], c["H"], c["W"], seed=c["seed"])
    xys = synth.make_query_points(clip, c["P"], seed=c["seed"])[:, :, 1:]
    preds, preds2, vis_e, ffeat, losses = model(xys.cuda(), clip["frames"][None].cuda(), iters=6, return_feat=True)
    # the full reference contract (pips.py:617-620): one coordinate estimate PER ITERATION, raw visibility logits, the initial
    # feature, losses=None; coord_predictions2 brackets the list with two copies of the initial and of the final estimate
    assert losses is None and len(preds) == 6 and len(preds2) == 10
    for it in range(6):
        assert (preds[it][0].cpu() - g["coords_per_iter"][it]).abs().max() < 1e-3, it
    assert torch.equal(preds2[0], preds2[1]) and torch.equal(preds2[-1], preds[-1]) and torch.equal(preds2[2], preds[0])
    assert (preds2[0][0, :, :, :].cpu() - xys[0][None]).abs().max() == 0          # zero-velocity initialisation
    assert (vis_e[0].cpu() - g["vis_e"]).abs().max() < 1e-3 * max(1.0, g["vis_e"].abs().max().item())
    assert (ffeat[0].cpu() - g["ffeat"]).abs().max() < 1e-4 * max(1.0, g["ffeat"].abs().max().item())
    # feat_init hand-over (the tracker's update pass, pips/tracker.py:93-101): golden run with xys + 3 px and the stored feature
    preds_f, _, vis_f, _ = model(xys.cuda() + 3.0, clip["frames"][None].cuda(), feat_init=g["ffeat"][None].cuda(), iters=6)
    assert (preds_f[-1][0].cpu() - g["coords_feat_init"]).abs().max() < 1e-3
    assert (vis_f[0].cpu() - g["vis_e_feat_init"]).abs().max() < 1e-3 * max(1.0, g["vis_e_feat_init"].abs().max().item())
    # coords_init: starting every slot at the query position IS the zero-velocity default
    ci = xys[0][None].repeat(8, 1, 1)[None]
    preds_c, _, _, _ = model(xys.cuda(), clip["frames"][None].cuda(), coords_init=ci.cuda(), iters=6)
    assert torch.equal(preds_c[-1], preds[-1])


def test_evaluate_batch_contract(sd, tmp_path):
    """PointTracker.evaluate_batch / unpack_results (reference sam_pt/point_tracker/tracker.py:47-118): forward + shape check +
    results detached, cloned and moved to the CPU under the reference's keys."""
    trk = _tracker(sd, tmp_path)
    clip = synth.make_clip(9, 64, 96, seed=3)
    q = synth.make_query_points(clip, 3, seed=3)
    traj, vis = trk(clip["frames"][None].cuda(), q.cuda())
    out = trk.evaluate_batch(clip["frames"][None].cuda(), q.cuda(), trajectories_gt=traj, visibilities_gt=vis)
    assert set(out) == {"trajectories_pred", "visibilities_pred", "query_points", "trajectories_gt", "visibilities_gt"}
    assert all(not v.is_cuda for v in out.values())
    assert out["trajectories_pred"].shape == (1, 9, 3, 2) and out["visibilities_pred"].shape == (1, 9, 3)
    assert torch.equal(out["trajectories_pred"], traj.cpu()) and torch.equal(out["visibilities_pred"], vis.cpu())
    rows = trk.unpack_results(out, batch_idx=7)
    assert len(rows) == 3 and rows[1]["idx"] == "7_0_1" and rows[1]["trajectory_pred"].shape == (9, 2)
    assert torch.equal(rows[2]["trajectory_gt"], traj[0, :, 2].cpu())


def _tracker(sd, tmp_path):
    from sam_pt.point_tracker.pips import PipsPointTracker
    d = synth.write_pips_checkpoint_dir(sd, str(tmp_path / "pips_ckpt"))
    return PipsPointTracker(checkpoint_path=d, stride=4, s=8).eval()


def test_tracker_matches_reference_golden(gold, sd, tmp_path):
    trk = _tracker(sd, tmp_path)
    for key in ("tracker", "c1_tracker"):
        g = gold[key]
        c = g["cfg"]
        clip = synth.make_clip(c["T"], c["H"], c["W"], seed=c["seed"])
        traj, vis = trk(clip["frames"][None].cuda(), g["query_points"].cuda())
        assert traj.shape == g["trajectories"].shape and vis.dtype == torch.bool
        assert (traj.cpu() - g["trajectories"]).abs().max() < 1e-3, key
        assert torch.equal(vis.cpu(), g["visibilities"]), key


def test_tracker_long_clip_matches_oracle(sd, tmp_path):
    """20 frames -> several linked windows, tail padding, points born mid-clip (both directions)."""
    trk = _tracker(sd, tmp_path)
    T, H, W = 20, 64, 96
    clip = synth.make_clip(T, H, W, seed=9)
    q = synth.make_query_points(clip, 5, seed=9)
    q[0, 3, 0] = 11.0
    q[0, 4, 0] = 19.0
    ref_t, ref_v = pips_ref.pips_tracker_forward(sd, clip["frames"][None], q)
    traj, vis = trk(clip["frames"][None].cuda(), q.cuda())
    assert (traj.cpu() - ref_t).abs().max() < 1e-3
    assert torch.equal(vis.cpu(), ref_v)
    # reference invariants (pips/tracker.py:188-194): query frame is a fixed point and visible
    for n in range(5):
        t0 = int(q[0, n, 0])
        assert torch.allclose(traj[0, t0, n].cpu(), q[0, n, 1:])
        assert bool(vis[0, t0, n])


def test_batch_size_gt1_raises(sd, tmp_path):
    trk = _tracker(sd, tmp_path)
    with pytest.raises(NotImplementedError):
        trk(torch.zeros((2, 4, 3, 32, 32), dtype=torch.uint8).cuda(), torch.zeros((2, 1, 3)).cuda())


def test_fnet_fp32_cuda_core_path_matches_oracle(sd):
    """The strict-fp32 CUDA-core convolution path (SAMPT_PIPS_TC=0) stays available and parity-checked."""
    from sam_pt.point_tracker.pips import Pips
    m = Pips(S=8, stride=4)
    m.fnet_on_tensor_cores = False
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    clip = synth.make_clip(2, 60, 106, seed=6)
    ref = pips_ref.fnet(sd, 2 * (clip["frames"].float() / 255.0) - 1.0)
    got = m.encode_frames(clip["frames"].cuda())[0].permute(0, 3, 1, 2).cpu()
    assert (got - ref).abs().max() < 2e-5 * max(ref.abs().max().item(), 1.0)

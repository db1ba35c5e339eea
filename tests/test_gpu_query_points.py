"""GPU parity of the query-point extraction: the k-medoids kernels pick EXACTLY the medoids of the restated
scikit-learn-extra algorithm (index work: bit-exact), and `SamPt.extract_query_points` draws the same points as the oracle
under the same torch seed (the RNG contract: host `torch.randperm`, as the reference)."""
import numpy as np
import pytest
import torch

from oracle import query_points_ref as R
from sam_pt.utils import query_points as Q

pytestmark = pytest.mark.gpu


def _blob(h, w, seed):
    g = torch.Generator().manual_seed(seed)
    lo = torch.rand((1, 1, 6, 8), generator=g)
    m = torch.nn.functional.interpolate(lo, size=(h, w), mode="bicubic", align_corners=False)[0, 0]
    return (m > m.median()).float()


@pytest.mark.parametrize("n,k,seed", [(40, 3, 0), (127, 8, 1), (129, 8, 2), (700, 16, 3), (1800, 16, 4), (2000, 5, 5), (1800, 72, 6)])
def test_kmedoids_picks_the_oracles_medoids(n, k, seed):
    px = _blob(120, 160, seed).nonzero().float()
    g = torch.Generator().manual_seed(seed)
    X = px[torch.randperm(len(px), generator=g)[:n]]
    info_ref, info = {}, {}
    C_ref = R.kmedoids_alternate(X.numpy(), k, info=info_ref)
    C = Q.kmedoids_gpu(X.cuda(), k, info=info)
    assert np.array_equal(info["medoid_idxs"], info_ref["medoid_idxs"]), (info["n_iter"], info_ref["n_iter"])
    assert info["n_iter"] == info_ref["n_iter"]
    assert np.array_equal(C.cpu().numpy(), C_ref)


def test_kmedoids_float_coordinates_and_duplicates():
    g = torch.Generator().manual_seed(9)
    X = torch.rand((300, 2), generator=g) * 50
    X[10] = X[11]                                   # duplicate points: zero off-diagonal distance
    info_ref, info = {}, {}
    R.kmedoids_alternate(X.numpy(), 6, info=info_ref)
    Q.kmedoids_gpu(X.cuda(), 6, info=info)
    assert np.array_equal(info["medoid_idxs"], info_ref["medoid_idxs"])


@pytest.mark.parametrize("pos,neg,method_neg", [(8, 0, "mixed"), (16, 1, "mixed"), (5, 12, "mixed"), (4, 3, "random"), (6, 2, "kmedoids")])
def test_extract_query_points_matches_oracle_under_the_same_seed(pos, neg, method_neg, tmp_path):
    from oracle import pips_ref
    from sampt_b200 import factory, synth
    clip = synth.make_clip(3, 120, 160, seed=5)
    masks = torch.stack([_blob(120, 160, 11), _blob(120, 160, 12)])
    ts = torch.tensor([0.0, 2.0])
    torch.manual_seed(72)
    ref = R.extract_query_points(clip["frames"], masks, ts, positive_method="kmedoids", negative_method=method_neg,
                                 positive_points_per_mask=pos, negative_points_per_mask=neg)
    ckpt = synth.write_pips_checkpoint_dir(synth.make_state_dict(pips_ref.pips_state_dict_shapes(), 7201), str(tmp_path / "pips"))
    model = factory.build_sam_pt("vit_test", None, ckpt, positive_points_per_mask=pos, negative_points_per_mask=neg)
    model.negative_point_selection_method = method_neg
    torch.manual_seed(72)
    got = model.extract_query_points(clip["frames"].cuda(), masks.cuda(), ts)
    assert got.is_cuda and got.shape == ref.shape
    assert torch.equal(got.cpu(), ref)


def test_degenerate_masks():
    z = torch.zeros((20, 30)).cuda()
    assert torch.equal(Q.extract_kmedoid_points(z, 4).cpu(), torch.zeros((4, 2)))
    t = torch.zeros((20, 30)); t[3, 4] = 1; t[5, 6] = 1
    assert Q.extract_kmedoid_points(t.cuda(), 5).cpu().tolist() == R.extract_kmedoid_points(t, 5).tolist()

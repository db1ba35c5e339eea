"""CPU checks of the query-point oracle (oracle/query_points_ref.py): its distance matrix against scikit-learn's own
`pairwise_distances` (the call scikit-learn-extra's KMedoids makes), and the k-medoids restatement's defining properties."""
import numpy as np
import torch

from oracle import query_points_ref as R


def _ellipse_mask(h=96, w=128, cy=40, cx=70, ay=22, ax=35):
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    return ((((ys - cy) / ay) ** 2 + ((xs - cx) / ax) ** 2) <= 1).float()


def test_distance_matrix_is_sklearns():
    from sklearn.metrics import pairwise_distances
    px = _ellipse_mask().nonzero().float().numpy()
    rng = np.random.default_rng(0)
    X = px[rng.permutation(len(px))[:700]]
    D = R.pairwise_euclidean_f32(X)
    assert D.dtype == np.float32
    assert np.array_equal(D, pairwise_distances(X, metric="euclidean"))


def test_kmedoids_restated_properties():
    px = _ellipse_mask().nonzero().float().numpy()
    rng = np.random.default_rng(1)
    X = px[rng.permutation(len(px))[:600]]
    info = {}
    C = R.kmedoids_alternate(X, 8, info=info)
    assert C.shape == (8, 2) and info["n_iter"] < 300
    med = info["medoid_idxs"]
    assert len(set(med.tolist())) == 8 and np.array_equal(C, X[med])
    # fixed point of the alternate update: every medoid minimises the in-cluster distance sum of its own cluster
    D = R.pairwise_euclidean_f32(X)
    labels = np.argmin(D[med, :], axis=0)
    for k in range(8):
        mem = np.where(labels == k)[0]
        costs = np.sum(D[mem, mem[:, None]], axis=1)
        assert costs[np.argmax(mem == med[k])] <= costs.min()
    # deterministic given X
    assert np.array_equal(C, R.kmedoids_alternate(X, 8))


def test_extract_points_contracts():
    m = _ellipse_mask()
    torch.manual_seed(72)
    p = R.extract_kmedoid_points(m, 8)
    assert p.shape == (8, 2) and p.dtype == torch.float32
    assert all(m[int(y), int(x)] == 1 for x, y in p.tolist())          # (x, y) order, on the mask
    tiny = torch.zeros((16, 16)); tiny[3, 4] = 1; tiny[5, 6] = 1
    q = R.extract_kmedoid_points(tiny, 5)                                 # fewer pixels than points: tiled
    assert q.tolist() == [[4, 3], [6, 5], [4, 3], [6, 5], [4, 3]]
    assert torch.equal(R.extract_kmedoid_points(torch.zeros((8, 8)), 3), torch.zeros((3, 2)))
    torch.manual_seed(1)
    r = R.extract_random_mask_points(m, 6)
    assert r.shape == (6, 2) and all(m[int(y), int(x)] == 1 for x, y in r.tolist())
    img = (torch.rand((3, 96, 128)) * 255).to(torch.uint8)
    torch.manual_seed(2)
    qp = R.extract_query_points(img[None], m[None], torch.tensor([0.0]), positive_points_per_mask=8, negative_points_per_mask=4)
    assert qp.shape == (1, 12, 3) and (qp[..., 0] == 0).all()
    assert all(m[int(y), int(x)] == 1 for x, y in qp[0, :8, 1:].tolist())
    assert all(m[int(y), int(x)] == 0 for x, y in qp[0, 8:, 1:].tolist())

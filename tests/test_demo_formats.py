"""Demo wire formats: the query-point text file and the frame-folder loader of demo/demo.py.
The expected values of the README example were produced by the reference's own `load_query_points`
(tests/golden/make_golden_demo.py executes that function's source from /root/reference/demo/demo.py)."""
import json
import os

import numpy as np
import torch

from demo.demo import load_demo_data, load_query_points, run_inference, save_query_points

HERE = os.path.dirname(os.path.abspath(__file__))


def test_query_points_file_matches_reference_parser(tmp_path):
    gold = json.load(open(os.path.join(HERE, "golden", "demo_query_points_golden.json")))
    for case in gold["cases"]:
        p = tmp_path / "q.txt"
        p.write_text(case["text"])
        qp, npos = load_query_points(str(p), case["frame_stride"], case["resize_factor"])
        assert npos == case["num_positive_points"]
        assert qp.dtype == torch.float32 and list(qp.shape) == case["shape"]
        assert torch.equal(qp, torch.tensor(case["query_points"], dtype=torch.float32))


def test_query_points_round_trip(tmp_path):
    q = torch.tensor([[[0.0, 10.5, 20.25], [0.0, 30.0, 30.0]], [[4.0, 123.123, 456.456], [4.0, 72.0, 72.0]]])
    save_query_points(str(tmp_path / "q.txt"), q, 1)
    q2, npos = load_query_points(str(tmp_path / "q.txt"), 1, 1.0)
    assert npos == 1 and torch.allclose(q, q2, atol=1e-5)


def test_frame_folder_loader(tmp_path):
    import cv2
    rng = np.random.default_rng(0)
    frames = []
    for i in range(5):
        img = rng.integers(0, 256, size=(24, 32, 3), dtype=np.uint8)
        frames.append(img)
        cv2.imwrite(str(tmp_path / f"{i:05d}.png"), img[:, :, ::-1])          # BGR on disk
    (tmp_path / "q.txt").write_text("1\n0 ; 4,5 6,7\n4 ; 8,9 10,11\n")
    rgbs, npos, qp = load_demo_data(str(tmp_path), str(tmp_path / "q.txt"), frame_stride=2, max_frames=3)
    assert rgbs.shape == (3, 3, 24, 32) and rgbs.dtype == torch.uint8 and npos == 1
    for k, i in enumerate((0, 2, 4)):
        assert np.array_equal(rgbs[k].permute(1, 2, 0).numpy(), frames[i])
    assert qp[:, 0, 0].tolist() == [0.0, 2.0]                                 # timesteps divided by the stride
    rgbs2, _, qp2 = load_demo_data(str(tmp_path), str(tmp_path / "q.txt"), frame_stride=2, longest_side_length=16)
    assert rgbs2.shape[-2:] == (12, 16) and torch.allclose(qp2[..., 1:], qp[..., 1:] * 0.5)


def test_run_inference_postprocess():
    class Fake:
        def __call__(self, video):
            T = len(video["image"]); M, P, _ = video["query_points"].shape
            return {"logits": [torch.full((T, 4, 6), float(m + 1)) for m in range(M)], "trajectories": torch.zeros((T, M, P, 2)),
                    "visibilities": torch.ones((T, M, P)), "scores": [1.0] * M}
    qp = torch.tensor([[[0.0, 1, 1]], [[2.0, 2, 2]]])
    logits, traj, vis, scores = run_inference(Fake(), torch.zeros((4, 3, 4, 6), dtype=torch.uint8), qp, (4, 6))
    assert logits.shape == (4, 3, 4, 6) and (logits[:, 0] == 0).all()
    assert (logits[:2, 2] == -1e8).all() and (logits[2:, 2] == 2).all() and (logits[:, 1] == 1).all()

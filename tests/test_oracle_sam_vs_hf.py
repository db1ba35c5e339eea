"""Pins oracle/sam_ref.py against the only independent SAM implementation available in this image
(`transformers.models.sam`, NOT the pinned upstream package) through a state-dict key remap.
The pinned `segment-anything @ aac76a1` is absent (no network) -> parity is otherwise unpinned."""
import re

import pytest
import torch

from oracle import sam_ref
from sampt_b200 import synth

transformers = pytest.importorskip("transformers")


def _remap(sd):
    out = {}
    for k, v in sd.items():
        nk = k
        if k.startswith("image_encoder."):
            nk = k.replace("image_encoder.", "vision_encoder.")
            nk = nk.replace("patch_embed.proj.", "patch_embed.projection.")
            nk = re.sub(r"blocks\.(\d+)\.norm(\d)", r"layers.\1.layer_norm\2", nk)
            nk = re.sub(r"blocks\.(\d+)\.", r"layers.\1.", nk)
            nk = nk.replace("neck.0.", "neck.conv1.").replace("neck.1.", "neck.layer_norm1.")
            nk = nk.replace("neck.2.", "neck.conv2.").replace("neck.3.", "neck.layer_norm2.")
        elif k.startswith("prompt_encoder."):
            if "positional_encoding_gaussian_matrix" in k:
                out["prompt_encoder.shared_embedding.positional_embedding"] = v
                out["shared_image_embedding.positional_embedding"] = v
                continue
            nk = nk.replace("point_embeddings.", "point_embed.")
            for a, b in (("0", "conv1"), ("1", "layer_norm1"), ("3", "conv2"), ("4", "layer_norm2"), ("6", "conv3")):
                nk = nk.replace(f"mask_downscaling.{a}.", f"mask_embed.{b}.")
        elif k.startswith("mask_decoder."):
            nk = re.sub(r"layers\.(\d)\.norm(\d)", r"layers.\1.layer_norm\2", nk)
            nk = nk.replace("norm_final_attn", "layer_norm_final_attn")
            nk = nk.replace("output_upscaling.0.", "upscale_conv1.").replace("output_upscaling.1.", "upscale_layer_norm.")
            nk = nk.replace("output_upscaling.3.", "upscale_conv2.")
            if "hypernetworks_mlps" in nk or "iou_prediction_head" in nk:
                nk = nk.replace("layers.0.", "proj_in.").replace("layers.2.", "proj_out.").replace("layers.1.", "layers.0.")
        out[nk] = v
    return out


@pytest.fixture(scope="module")
def setup():
    from transformers import SamConfig, SamModel, SamVisionConfig

    cfg = sam_ref.VIT_TEST
    sd = synth.condition_sam(synth.make_state_dict(sam_ref.sam_state_dict_shapes(cfg), 11))
    vc = SamVisionConfig(hidden_size=cfg.embed_dim, num_hidden_layers=cfg.depth, num_attention_heads=cfg.num_heads,
                         global_attn_indexes=list(cfg.global_attn_indexes), mlp_dim=cfg.embed_dim * 4,
                         output_channels=256, window_size=14)
    m = SamModel(SamConfig(vision_config=vc)).eval()
    missing, unexpected = m.load_state_dict(_remap(sd), strict=False)
    assert not unexpected, unexpected
    assert not missing, missing
    g = torch.Generator().manual_seed(3)
    x = torch.randn((1, 3, 1024, 1024), generator=g)
    return cfg, sd, m, x


def test_image_encoder_matches_hf(setup):
    cfg, sd, m, x = setup
    with torch.no_grad():
        ours = sam_ref.vit_encode(sd, x, cfg)
        theirs = m.get_image_embeddings(x)
    assert ours.shape == theirs.shape == (1, 256, 64, 64)
    assert (ours - theirs).abs().max() < 2e-4 * max(1.0, theirs.abs().max().item())


@pytest.mark.parametrize("with_mask,with_box", [(False, False), (True, False), (True, True)])
def test_prompt_and_decoder_match_hf(setup, with_mask, with_box):
    cfg, sd, m, x = setup
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        feats = sam_ref.vit_encode(sd, x, cfg)
        pts = torch.rand((1, 5, 2), generator=g) * 1000
        labels = torch.tensor([[1, 1, 0, 1, 1]])
        mask_in = torch.randn((1, 1, 256, 256), generator=g) if with_mask else None
        box = torch.tensor([[100.0, 150.0, 700.0, 640.0]]) if with_box else None
        sparse, dense = sam_ref.prompt_encode(sd, (pts, labels), box, mask_in)
        low, iou = sam_ref.mask_decode(sd, feats, sam_ref.get_dense_pe(sd), sparse, dense, multimask_output=False)
        kw = dict(image_embeddings=feats, input_points=pts[:, None], input_labels=labels[:, None], multimask_output=False)
        if with_mask:
            kw["input_masks"] = mask_in
        if with_box:
            kw["input_boxes"] = box[:, None]
        out = m(**kw)
    theirs = out.pred_masks[0, 0]
    assert (low[0] - theirs).abs().max() < 2e-4 * max(1.0, theirs.abs().max().item())
    assert (iou[0] - out.iou_scores[0, 0]).abs().max() < 1e-4

"""Pins the HQ-SAM decoder restatement in oracle/sam_ref.py (MaskDecoderHQ of m43/sam-hq @ 75c73fa, un-vendored) against
the independent `transformers.models.sam_hq` via a key remap (single-mask output, hq_token_only False/True).

NB transformers' port up-scales the PRE-transformer image embedding (spatially transposed) where upstream sam-hq up-scales
the transformer output (`src`); the oracle reproduces that deviation behind `hf_upscale_quirk` ONLY for this cross-check, so
that every other HQ piece (hq token, hq MLP, compress_vit_feat, embedding_encoder, embedding_maskfeature, mask sum) is pinned.
The upstream behaviour itself (transformer output up-scaled) is the plain-SAM path already pinned in test_oracle_sam_vs_hf."""

import pytest
import torch

from oracle import sam_ref
from sampt_b200 import synth
from tests.test_oracle_sam_vs_hf import _remap as _remap_sam

pytest.importorskip("transformers")


def _remap(sd):
    out = _remap_sam(sd)
    ren = {}
    for k, v in out.items():
        nk = k
        nk = nk.replace("mask_decoder.hf_token.", "mask_decoder.hq_token.")
        if "mask_decoder.hf_mlp." in nk:
            nk = nk.replace("hf_mlp.", "hq_mask_mlp.")
            nk = nk.replace("layers.0.", "proj_in.").replace("layers.2.", "proj_out.").replace("layers.1.", "layers.0.")
        for a, b in (("compress_vit_feat.0.", "compress_vit_conv1."), ("compress_vit_feat.1.", "compress_vit_norm."),
                     ("compress_vit_feat.3.", "compress_vit_conv2."), ("embedding_encoder.0.", "encoder_conv1."),
                     ("embedding_encoder.1.", "encoder_norm."), ("embedding_encoder.3.", "encoder_conv2."),
                     ("embedding_maskfeature.0.", "mask_conv1."), ("embedding_maskfeature.1.", "mask_norm."),
                     ("embedding_maskfeature.3.", "mask_conv2.")):
            nk = nk.replace("mask_decoder." + a, "mask_decoder." + b)
        ren[nk] = v
    return ren


def test_hq_decoder_matches_hf():
    from transformers import SamHQConfig, SamHQModel, SamHQVisionConfig
    from transformers.models.sam_hq.configuration_sam_hq import SamHQMaskDecoderConfig
    cfg = sam_ref.VIT_TEST
    sd = synth.condition_sam(synth.make_state_dict(sam_ref.sam_state_dict_shapes(cfg, hq=True), 13))
    vc = SamHQVisionConfig(hidden_size=cfg.embed_dim, num_hidden_layers=cfg.depth, num_attention_heads=cfg.num_heads,
                           global_attn_indexes=list(cfg.global_attn_indexes), mlp_dim=cfg.embed_dim * 4, output_channels=256,
                           window_size=14)
    m = SamHQModel(SamHQConfig(vision_config=vc, mask_decoder_config=SamHQMaskDecoderConfig(vit_dim=cfg.embed_dim))).eval()
    missing, unexpected = m.load_state_dict(_remap(sd), strict=False)
    assert not unexpected, unexpected
    assert not missing, missing
    g = torch.Generator().manual_seed(4)
    x = torch.randn((1, 3, 1024, 1024), generator=g)
    with torch.no_grad():
        feats, interm = sam_ref.vit_encode(sd, x, cfg, return_interm=True)
        pts = torch.rand((1, 5, 2), generator=g) * 1000
        labels = torch.tensor([[1, 1, 0, 1, 1]])
        mask_in = torch.randn((1, 1, 256, 256), generator=g)
        box = torch.tensor([[100.0, 150.0, 700.0, 640.0]])
        sparse, dense = sam_ref.prompt_encode(sd, (pts, labels), box, mask_in)
        low, iou = sam_ref.mask_decode(sd, feats, sam_ref.get_dense_pe(sd), sparse, dense, False,
                                       hq={"interm": interm[0], "hq_token_only": False, "hf_upscale_quirk": True})
        out = m(image_embeddings=feats, intermediate_embeddings=[i for i in interm], input_points=pts[:, None],
                input_labels=labels[:, None], input_masks=mask_in, input_boxes=box[:, None], multimask_output=False,
                hq_token_only=False)
    theirs = out.pred_masks[0, 0]
    assert (low[0] - theirs).abs().max() < 3e-4 * max(1.0, theirs.abs().max().item())
    assert (iou[0] - out.iou_scores[0, 0]).abs().max() < 1e-4

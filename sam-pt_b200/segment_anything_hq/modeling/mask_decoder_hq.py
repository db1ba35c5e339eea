"""`MaskDecoderHQ` parameter container (m43/sam-hq segment_anything/modeling/mask_decoder_hq.py; kwargs per
/root/reference/configs/model/sam/samhq_vit_huge.yaml:25-27 + mask_decoder/sam.yaml).  Arithmetic: csrc/decoder.cu."""
from torch import nn

from sampt_b200.param_tree import build_param_tree
from segment_anything.modeling.mask_decoder import MaskDecoder


class MaskDecoderHQ(MaskDecoder):
    def __init__(self, *, transformer_dim: int, transformer: nn.Module, num_multimask_outputs: int = 3, activation=nn.GELU,
                 iou_head_depth: int = 3, iou_head_hidden_dim: int = 256, vit_dim: int = 1024) -> None:
        super().__init__(transformer_dim=transformer_dim, transformer=transformer, num_multimask_outputs=num_multimask_outputs,
                         activation=activation, iou_head_depth=iou_head_depth, iou_head_hidden_dim=iou_head_hidden_dim)
        self.vit_dim = vit_dim
        c = transformer_dim
        shapes = {"hf_token.weight": (1, c),
                  "hf_mlp.layers.0.weight": (c, c), "hf_mlp.layers.0.bias": (c,), "hf_mlp.layers.1.weight": (c, c),
                  "hf_mlp.layers.1.bias": (c,), "hf_mlp.layers.2.weight": (c // 8, c), "hf_mlp.layers.2.bias": (c // 8,),
                  "compress_vit_feat.0.weight": (vit_dim, c, 2, 2), "compress_vit_feat.0.bias": (c,),
                  "compress_vit_feat.1.weight": (c,), "compress_vit_feat.1.bias": (c,),
                  "compress_vit_feat.3.weight": (c, c // 8, 2, 2), "compress_vit_feat.3.bias": (c // 8,),
                  "embedding_encoder.0.weight": (c, c // 4, 2, 2), "embedding_encoder.0.bias": (c // 4,),
                  "embedding_encoder.1.weight": (c // 4,), "embedding_encoder.1.bias": (c // 4,),
                  "embedding_encoder.3.weight": (c // 4, c // 8, 2, 2), "embedding_encoder.3.bias": (c // 8,),
                  "embedding_maskfeature.0.weight": (c // 4, c // 8, 3, 3), "embedding_maskfeature.0.bias": (c // 4,),
                  "embedding_maskfeature.1.weight": (c // 4,), "embedding_maskfeature.1.bias": (c // 4,),
                  "embedding_maskfeature.3.weight": (c // 8, c // 4, 3, 3), "embedding_maskfeature.3.bias": (c // 8,)}
        build_param_tree(self, shapes, seed=16384)

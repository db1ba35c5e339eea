"""HQ-SAM's ViT: identical to SAM's but the predictor also consumes the output of the first global-attention block
(`interm_embeddings[0]`, the only one MaskDecoderHQ reads)."""
from segment_anything.modeling.image_encoder import ImageEncoderViT as _Base


class ImageEncoderViT(_Base):
    returns_interm = True

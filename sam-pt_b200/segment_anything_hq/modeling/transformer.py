from segment_anything.modeling.transformer import TwoWayTransformer  # noqa: F401

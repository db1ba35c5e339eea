"""Drop-in `sam_pt` package of the B200-native SAM-PT hot path (mirrors the dotted paths the reference's Hydra
configs target).  Sub-packages are imported lazily: the reference's eager
`sam_pt/point_tracker/__init__.py:1-7` pulls in tensorflow/jax trackers that are out of scope here."""

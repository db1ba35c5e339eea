"""Drop-in `segment_anything_hq` (the reference installs the fork m43/sam-hq @ 75c73fa from git, requirements.txt:30; it is
not vendored).  Exposes the names the reference imports / targets from YAML (configs/model/sam/samhq_vit_huge.yaml,
configs/model/sam_pt.yaml:8)."""
from .modeling import Sam  # noqa: F401
from .predictor import SamPredictor  # noqa: F401

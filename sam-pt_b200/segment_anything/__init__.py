"""Drop-in `segment_anything` (the reference installs it from git, requirements.txt:28; it is not vendored).  Exposes
exactly the names the reference imports / targets from YAML."""
from .modeling import Sam  # noqa: F401
from .predictor import SamPredictor  # noqa: F401

from .pips import Pips  # noqa: F401
from .tracker import PipsPointTracker  # noqa: F401

"""Point trackers on the hot path: PIPS and CoTracker (north star).  The reference additionally imports RAFT,
SuperGlue, TAPIR, TapNet and PIPS++ eagerly (sam_pt/point_tracker/__init__.py:1-7); those are out of scope."""
from .tracker import PointTracker  # noqa: F401


def __getattr__(name):
    if name == "PipsPointTracker":
        from .pips import PipsPointTracker
        return PipsPointTracker
    if name == "CoTrackerPointTracker":
        from .cotracker import CoTrackerPointTracker
        return CoTrackerPointTracker
    if name == "SuperGluePointTracker":  # SamPt only uses it in an isinstance check (sam_pt.py:189)
        class SuperGluePointTracker:  # never instantiated here
            pass
        return SuperGluePointTracker
    raise AttributeError(name)

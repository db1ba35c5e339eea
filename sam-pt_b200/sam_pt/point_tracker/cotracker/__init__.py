from .tracker import CoTrackerPointTracker  # noqa: F401

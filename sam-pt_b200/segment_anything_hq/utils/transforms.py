from segment_anything.utils.transforms import ResizeLongestSide  # noqa: F401

"""Test-only stand-in for matplotlib (absent from this image). The reference's utils/visualizer.py imports
matplotlib.pyplot at module level; the trainer only calls save_sample_grid from it, which needs torchvision + PIL, not
matplotlib (utils/visualizer.py:133-177). Any actual plotting call raises."""


def __getattr__(name):
    raise AttributeError(f"matplotlib stub: {name!r} is not available in the test environment")

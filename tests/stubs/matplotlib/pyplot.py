"""See tests/stubs/matplotlib/__init__.py."""


def __getattr__(name):
    raise AttributeError(f"matplotlib.pyplot stub: {name!r} is not available in the test environment")

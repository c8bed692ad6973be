"""Placeholder for the `roma` package, which the reference imports at module level (must3r/engine/inference.py:7,
must3r/slam/model.py) but only calls inside postprocess(compute_cam=True).  It is not installed in this image; the
parity harness never takes that branch with the reference, so any use raises."""


def __getattr__(name):
    raise ImportError(f"roma.{name}: the real `roma` package is not installed (stub at baseline/stubs/roma)")

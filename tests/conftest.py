import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, HERE):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The built libraries are git-ignored: (re)build them in-tree before the tests -- `make` is
    incremental, so this is a no-op when they are current and a rebuild when a source changed
    (nvcc cross-compiles without a GPU).  A failed build stops the session with the compiler output."""
    import shutil
    import subprocess
    for sub in (os.path.join("slide_slam_b200", "csrc"), "oracle"):
        if shutil.which("make") is None:
            break  # no toolchain on this box: the prebuilt libraries travel with the snapshot
        r = subprocess.run(["make", "-C", os.path.join(ROOT, sub)], capture_output=True, text=True)
        if r.returncode != 0:
            built = os.path.join(ROOT, "slide_slam_b200", "libslide_pr.so") if sub != "oracle" else os.path.join(ROOT, "oracle", "libslide_oracle.so")
            if not os.path.exists(built):
                pytest.exit(f"building {sub} failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}", returncode=3)
            import warnings
            warnings.warn(f"rebuilding {sub} failed, testing the existing library:\n{r.stderr[-2000:]}")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

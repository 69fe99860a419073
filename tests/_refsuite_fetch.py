"""Stage the reference's own test files where the runner (and the GPU box) can see them.

``/root/reference`` exists only in the build container.  Like ``oracle/_ref`` (the reference
extension compiled from its sources), the test files are staged -- unmodified, by this
committed recipe -- under ``baseline/_ref/tests/``: git-ignored (never part of this repo's
history) but not gpurun-ignored, so the ``-m gpu`` run on the B200 box can execute
``test_cuda_kernels.py`` and the CUDA-parametrised cases of the other files.
"""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/tests"
DST = os.path.join(ROOT, "baseline", "_ref", "tests")

# the reference test files that exercise the hot path
HOT_PATH_FILES = (
    "test_iir.py", "test_biquad.py", "test_fused.py", "test_chain_fusion.py", "test_fftconv.py", "test_fir.py",
    "test_filterbank.py", "test_cuda_fallback.py", "test_ops_dispatch.py", "test_iir_gaps.py", "test_filter_base.py",
    "test_filter_utils.py", "test_cuda_kernels.py",
)


def tests_dir() -> str | None:
    """Directory holding the reference test files, or None when neither copy exists."""
    for d in (SRC, DST):
        if os.path.isfile(os.path.join(d, "test_iir.py")):
            return d
    return None


def stage() -> str | None:
    if not os.path.isdir(SRC):
        return tests_dir()
    os.makedirs(DST, exist_ok=True)
    for f in HOT_PATH_FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    return DST


if __name__ == "__main__":
    print(stage())

"""privacy_preserving_sfm_b200 — B200 (sm_100a) implementation of the privacy-preserving-SfM hot
path (line-lifted absolute pose under RANSAC + line-reprojection bundle adjustment).

The product is ``libppsfm_b200.so`` (C-ABI in ``include/ppsfm_b200.h``).  This package is a thin
ctypes host layer over it whose function names, argument meaning and error behaviour mirror the
reference's C++ API (``src/estimators/pose.h``, ``src/optim/ransac.h``,
``src/optim/bundle_adjustment.h``).  There is no CPU fallback: if the shared library or a CUDA
device is missing, calls raise.

Host-side data modules around the path (numpy, no GPU involved, each checked against the
reference's own code where it compiles — tests/test_ref_*.py): ``model_io`` (the fork's text
model format, ``Reconstruction::ReadText`` / ``WriteText`` / ``Normalize``),
``correspondence_graph`` (``CorrespondenceGraph``), ``lifting`` (keypoints -> lifted lines,
camera models in both directions, the database's line blob), ``mapper`` (a test driver with the
control flow of ``IncrementalMapper`` over the GPU operators).
"""
from .binding import (  # noqa: F401
    Context, RANSACOptions, RansacReport, RansacTiming, PpsfmError, load_library, library_path,
    build_library,
)
from .estimators import (  # noqa: F401
    EstimateAbsolutePoseFromLines, RANSAC_P6L, P6LEstimator, ComputeSquaredLineReprojectionError,
    ComputeNumTrials,
)

__all__ = [
    "Context", "RANSACOptions", "RansacReport", "RansacTiming", "PpsfmError", "load_library",
    "library_path", "build_library", "EstimateAbsolutePoseFromLines", "RANSAC_P6L",
    "P6LEstimator", "ComputeSquaredLineReprojectionError", "ComputeNumTrials",
]

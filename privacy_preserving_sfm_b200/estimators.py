"""Host-side mirror of the reference's estimator API for the absolute-pose path.

Names, argument order/meaning and failure behaviour follow the reference:
  * ``EstimateAbsolutePoseFromLines``   — src/estimators/pose.h:110-115, pose.cc:52-94
  * ``RANSAC_P6L`` (= ``RANSAC<P6LEstimator>``) — src/optim/ransac.h:78-137, 178-278
  * ``P6LEstimator.Estimate / Residuals`` — src/estimators/absolute_pose.h:48-75
  * ``ComputeSquaredLineReprojectionError`` — src/estimators/utils.cc:40-89
Everything executes in libppsfm_b200.so on the GPU; nothing here computes on the CPU.
"""
import numpy as np

from . import binding
from .binding import RANSACOptions


def _split_feature_lines(lines2D):
    """Accepts an (n, 3) array, or a pair (lines[n,3], aligned[n]) standing for FeatureLines
    (src/feature/types.h:98-149: Line() + IsAligned())."""
    if isinstance(lines2D, (tuple, list)) and len(lines2D) == 2:
        lines, aligned = lines2D
        return np.asarray(lines, dtype=np.float64), np.asarray(aligned, dtype=np.uint8)
    lines = np.asarray(lines2D, dtype=np.float64)
    return lines, np.zeros(lines.shape[0], dtype=np.uint8)


def ComputeNumTrials(num_inliers, num_samples, confidence, num_trials_multiplier):
    """RANSAC<P6LEstimator>::ComputeNumTrials, src/optim/ransac.h:158-176."""
    return int(binding.load_library().ppsfm_compute_num_trials(
        num_inliers, num_samples, confidence, num_trials_multiplier))


def ComputeSquaredLineReprojectionError(lines2D, points3D, proj_matrix, ctx=None):
    """src/estimators/utils.cc:40-89.  proj_matrix: 3x4 array.  Returns residuals[n]."""
    ctx = ctx or binding.default_context()
    lines, _ = _split_feature_lines(lines2D)
    points3D = np.asarray(points3D, dtype=np.float64)
    if lines.shape[0] != points3D.shape[0]:
        raise binding.PpsfmError("CHECK_EQ(lines2D.size(), points3D.size())")
    model = np.asarray(proj_matrix, dtype=np.float64).reshape(3, 4).T.reshape(-1)  # col-major
    res, _, _ = ctx.line_residuals(lines, points3D, model[None, :], 1.0)
    return res[0]


class P6LEstimator:
    """src/estimators/absolute_pose.h:48-75.  X_t = FeatureLine, Y_t = Vector3d,
    M_t = Matrix3x4d, kMinNumSamples = 6."""
    kMinNumSamples = 6

    def __init__(self, ctx=None):
        self._ctx = ctx or binding.default_context()

    def Estimate(self, lines2D, points3D):
        """6 correspondences -> list of 3x4 poses (0..8)."""
        lines, aligned = _split_feature_lines(lines2D)
        if lines.shape[0] != 6:
            raise binding.PpsfmError("P6LEstimator::Estimate needs exactly 6 correspondences")
        models, nm = self._ctx.p6l_solve_batch(lines, aligned, points3D, np.arange(6)[None, :])
        return [models[0, m].reshape(4, 3).T.copy() for m in range(int(nm[0]))]

    def Residuals(self, lines2D, points3D, proj_matrix):
        return ComputeSquaredLineReprojectionError(lines2D, points3D, proj_matrix, self._ctx)


class RANSAC_P6L:
    """RANSAC<P6LEstimator, InlierSupportMeasurer, RandomSampler> (src/optim/ransac.h)."""

    def __init__(self, options, ctx=None):
        options.Check()
        self.options = options
        self._ctx = ctx or binding.default_context()
        self.estimator = P6LEstimator(self._ctx)

    def Estimate(self, X, Y):
        """X: FeatureLines ((n,3) or (lines, aligned)); Y: points (n,3). Returns a Report."""
        lines, aligned = _split_feature_lines(X)
        Y = np.asarray(Y, dtype=np.float64)
        if lines.shape[0] != Y.shape[0]:
            raise binding.PpsfmError("CHECK_EQ(X.size(), Y.size())")
        rep, mask = self._ctx.ransac_p6l(lines, aligned, Y, self.options)
        rep.inlier_mask = mask if rep.success else np.zeros(0, dtype=np.uint8)
        return rep


def EstimateAbsolutePoseFromLines(options, lines2D, points3D, ctx=None):
    """src/estimators/pose.cc:52-94.

    Returns (ok, qvec[w,x,y,z], tvec, num_inliers, inlier_mask) — the reference's bool return
    plus its four output pointers."""
    options.Check()
    ctx = ctx or binding.default_context()
    lines, aligned = _split_feature_lines(lines2D)
    ok, q, t, ninl, mask, _ = ctx.estimate_absolute_pose_from_lines(lines, aligned, points3D,
                                                                    options)
    return ok, q, t, ninl, mask

"""Seeded synthetic scenes for the parity tests and bench.py (numpy only).

Recipes follow SURVEY.md §8(d), which models them on the reference's own test generators
(src/init/initializer_test.cc:52-137 `setup_plausible_scene` / `setup_random_lines`):
lines are stored in normalised camera coordinates with ||(a, b)|| = 1
(src/feature/types.h:98-138, src/base/cost_functions.h:51-52).
"""
import numpy as np

SCENE_SEED = 20201017


def random_rotation(rng):
    """Uniform random rotation from a normalised Gaussian quaternion (w, x, y, z)."""
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return quat_to_rotmat(q)


def quat_to_rotmat(q):
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def rotmat_to_quat(R):
    """Trace-branch conversion, (w, x, y, z)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        return np.array([w, (R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s,
                         (R[1, 0] - R[0, 1]) * s])
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
    v = np.zeros(3)
    v[i] = 0.5 * s
    s = 0.5 / s
    w = (R[k, j] - R[j, k]) * s
    v[j] = (R[j, i] + R[i, j]) * s
    v[k] = (R[k, i] + R[i, k]) * s
    return np.array([w, v[0], v[1], v[2]])


def model_from_pose(R, t):
    """3x4 [R|t] flattened column-major (Eigen::Matrix3x4d memory order)."""
    return np.concatenate([np.asarray(R).T.reshape(-1), np.asarray(t).reshape(-1)])


def make_abs_pose_scene(n=50000, inlier_ratio=0.30, noise_px=1.0, focal=1000.0,
                        aligned_fraction=0.30, seed=SCENE_SEED):
    """Config 2: line<->3-D-point correspondences for absolute pose.

    Returns dict(lines[n,3], aligned[n] u8, points[n,3], R, t, is_inlier[n]).
    """
    rng = np.random.default_rng(seed)
    R = random_rotation(rng)
    t = rng.uniform(-1.0, 1.0, size=3)
    # camera-frame coordinates: x, y in U[-2, 2], z in U[2, 6]
    pc = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(2, 6, n)], axis=1)
    points = (pc - t) @ R  # X = R^T (pc - t)
    uv = pc[:, :2] / pc[:, 2:3]
    uv_noisy = uv + rng.normal(scale=noise_px / focal, size=uv.shape)
    theta = rng.uniform(0, 2 * np.pi, n)
    a, b = np.cos(theta), np.sin(theta)
    c = -(a * uv_noisy[:, 0] + b * uv_noisy[:, 1])
    lines = np.stack([a, b, c], axis=1)
    is_inlier = rng.uniform(size=n) < inlier_ratio
    # outliers: independent random 3-D point in front of the camera
    n_out = int((~is_inlier).sum())
    pc_out = np.stack([rng.uniform(-2, 2, n_out), rng.uniform(-2, 2, n_out),
                       rng.uniform(2, 6, n_out)], axis=1)
    points[~is_inlier] = (pc_out - t) @ R
    aligned = (rng.uniform(size=n) < aligned_fraction).astype(np.uint8)
    return dict(lines=np.ascontiguousarray(lines), aligned=aligned,
                points=np.ascontiguousarray(points), R=R, t=t, is_inlier=is_inlier,
                focal=focal)


def make_p6l_minimal_problems(count, seed=1):
    """`count` noise-free generic 6-correspondence problems with their generating pose."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        R = random_rotation(rng)
        t = rng.uniform(-1, 1, 3)
        pc = np.stack([rng.uniform(-2, 2, 6), rng.uniform(-2, 2, 6), rng.uniform(2, 6, 6)], axis=1)
        X = (pc - t) @ R
        uv = pc[:, :2] / pc[:, 2:3]
        th = rng.uniform(0, 2 * np.pi, 6)
        a, b = np.cos(th), np.sin(th)
        c = -(a * uv[:, 0] + b * uv[:, 1])
        out.append(dict(lines=np.stack([a, b, c], 1), points=X, R=R, t=t))
    return out


# --------------------------------------------------------------------------------------------
# Bundle-adjustment scenes (configs 3 / 4)
# --------------------------------------------------------------------------------------------
def look_at_rotation(cam_center, target=np.zeros(3), up=np.array([0.0, 1.0, 0.0])):
    z = target - cam_center
    z /= np.linalg.norm(z)
    x = np.cross(up, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return np.stack([x, y, z], axis=0)  # rows = camera axes in world coords -> R (world->cam)


def make_ba_scene(num_cams=100, num_points=30000, obs_per_point=10, focal=1000.0,
                  noise_px=0.5, rot_sigma_deg=0.5, trans_sigma=0.01, point_sigma=0.01,
                  seed=SCENE_SEED):
    """Cameras on a circle of radius 5 looking at the origin (+5 % jitter), one shared PINHOLE
    camera (f, f, 500, 500); each point observed by `obs_per_point` distinct cameras; observed
    line = random-normal line through the noisy projection.  Observations are point-major.

    Returns dict with ground truth (`*_gt`) and the perturbed initial state.
    qvec is (w, x, y, z); poses are world->camera.
    """
    rng = np.random.default_rng(seed)
    ang = 2 * np.pi * np.arange(num_cams) / num_cams
    centers = np.stack([5 * np.cos(ang), np.zeros(num_cams), 5 * np.sin(ang)], axis=1)
    centers *= 1 + 0.05 * rng.normal(size=(num_cams, 1))
    centers[:, 1] += 0.25 * rng.normal(size=num_cams)
    qvecs = np.zeros((num_cams, 4))
    tvecs = np.zeros((num_cams, 3))
    Rs = np.zeros((num_cams, 3, 3))
    for i in range(num_cams):
        R = look_at_rotation(centers[i])
        Rs[i] = R
        q = rotmat_to_quat(R)
        qvecs[i] = q / np.linalg.norm(q)
        tvecs[i] = -R @ centers[i]
    points = rng.uniform(-1, 1, size=(num_points, 3))
    # visibility: obs_per_point distinct cameras per point
    cam_idx = np.empty((num_points, obs_per_point), dtype=np.int32)
    chunk = max(1, (1 << 24) // max(1, num_cams))      # bound the key matrix to ~128 MB
    for lo in range(0, num_points, chunk):
        hi = min(num_points, lo + chunk)
        keys = rng.random((hi - lo, num_cams))
        if obs_per_point < num_cams:
            part = np.argpartition(keys, obs_per_point - 1, axis=1)[:, :obs_per_point]
        else:
            part = np.tile(np.arange(num_cams), (hi - lo, 1))
        cam_idx[lo:hi] = part
    cam_idx.sort(axis=1)
    obs_cam = cam_idx.reshape(-1)
    obs_pt = np.repeat(np.arange(num_points, dtype=np.int32), obs_per_point)
    pc = np.einsum('oij,oj->oi', Rs[obs_cam], points[obs_pt]) + tvecs[obs_cam]
    uv = pc[:, :2] / pc[:, 2:3]
    uv += rng.normal(scale=noise_px / focal, size=uv.shape)
    th = rng.uniform(0, 2 * np.pi, len(obs_cam))
    a, b = np.cos(th), np.sin(th)
    c = -(a * uv[:, 0] + b * uv[:, 1])
    obs_line = np.stack([a, b, c], axis=1)
    # perturbed initial state
    q0 = qvecs.copy()
    t0 = tvecs.copy()
    for i in range(num_cams):
        w = rng.normal(scale=np.deg2rad(rot_sigma_deg) / np.sqrt(3), size=3)
        ang_i = np.linalg.norm(w)
        dq = np.concatenate([[np.cos(ang_i / 2)], np.sin(ang_i / 2) * w / max(ang_i, 1e-300)])
        q0[i] = quat_mul(dq, qvecs[i])
        t0[i] = tvecs[i] * (1 + trans_sigma * rng.normal(size=3))
    p0 = points + point_sigma * 2.0 * rng.normal(size=points.shape)
    # gauge: camera 0 keeps its true pose, camera 1 keeps tvec[0]
    q0[0], t0[0] = qvecs[0], tvecs[0]
    t0[1, 0] = tvecs[1, 0]
    cam_params = np.array([focal, focal, 500.0, 500.0])
    return dict(qvecs_gt=qvecs, tvecs_gt=tvecs, points_gt=points, qvecs=q0, tvecs=t0, points=p0,
                obs_cam=obs_cam.astype(np.int32), obs_pt=obs_pt.astype(np.int32),
                obs_line=np.ascontiguousarray(obs_line), cam_params=cam_params,
                num_cams=num_cams, num_points=num_points)


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw])


def _rotation_between(a, b):
    """Minimal rotation taking direction a onto b (Eigen::Quaterniond::FromTwoVectors)."""
    a, b = a / np.linalg.norm(a), b / np.linalg.norm(b)
    v, c = np.cross(a, b), float(a @ b)
    K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return np.eye(3) + K + K @ K / (1.0 + c)


def make_init_scene(num_points=100, num_aligned=50, num_outliers=0, seed=SCENE_SEED,
                    tilt_deg=0.0):
    """Four-view initialisation scene after the reference's own test recipe
    (src/init/initializer_test.cc:44-137): camera 0 = identity, |t_1| = 1, upright cameras (rotation
    about the gravity axis y), points uniform in [-1, 1]^2 x [0, 1] in front of all cameras,
    `num_aligned` gravity-aligned lines l = normalize(x~ x g) and the rest through random normals,
    outliers = one view's point replaced by a random one.  ``tilt_deg > 0`` additionally rotates
    every camera frame by a random rotation of up to that angle (gravity no longer along the
    camera y axis; the scene keeps all points in front of the gravity-aligned cameras, which is
    what the bearing sign convention of initializer.cc:87-89 assumes).
    Returns lines [4, n, 3], aligned [4, n], gravity [4, 3], poses [4, 3, 4] (ground truth)."""
    rng = np.random.default_rng(seed)
    ey = np.array([0.0, 1.0, 0.0])
    while True:
        cams, tilts = [], []
        for i in range(4):
            P = np.zeros((3, 4))
            if i == 0:
                P[:, :3] = np.eye(3)
            else:
                a = rng.uniform(-np.pi, np.pi)
                P[:, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0],
                                     [-np.sin(a), 0, np.cos(a)]])
                P[:, 3] = rng.uniform(-1, 1, 3)
            if i == 1:
                P[:, 3] /= np.linalg.norm(P[:, 3])
            cams.append(P)
            axis = rng.normal(size=3)
            axis /= np.linalg.norm(axis)
            ang = np.deg2rad(tilt_deg) * rng.uniform(0.3, 1.0)
            K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
            tilts.append(np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K)
        X = rng.uniform(-1, 1, (num_points, 3))
        X[:, 2] = np.abs(X[:, 2])
        Xh = np.concatenate([X, np.ones((num_points, 1))], axis=1)
        z = [Xh @ P.T for P in cams]
        ok = all((zi[:, 2] > 1e-3).all() for zi in z)
        if ok and tilt_deg > 0:   # depth in the gravity-aligned frame Rg Q P X
            for zi, Q in zip(z, tilts):
                Rg = _rotation_between(Q @ ey, ey)
                ok = ok and ((zi @ (Rg @ Q).T)[:, 2] > 1e-3).all()
        if ok:
            break
    x = [zi[:, :2] / zi[:, 2:3] for zi in z]
    order = rng.permutation(num_points)
    for i in range(num_outliers):
        x[rng.integers(0, 4)][order[i]] = rng.uniform(-1, 1, 2)
    is_aligned = rng.permutation(num_points) < num_aligned
    lines = np.zeros((4, num_points, 3))
    gravity = np.zeros((4, 3))
    poses = np.stack(cams)
    for i in range(4):
        g = cams[i][:, 1].copy()
        xh = np.concatenate([x[i], np.ones((num_points, 1))], axis=1)
        nrm = np.where(is_aligned[:, None], g[None, :], rng.uniform(-1, 1, (num_points, 3)))
        l = np.cross(xh, nrm)
        lines[i] = l / np.linalg.norm(l, axis=1, keepdims=True)
        gravity[i] = g
        if tilt_deg > 0:
            Q = tilts[i]
            lines[i] = lines[i] @ Q.T
            gravity[i] = Q @ g
            poses[i] = Q @ poses[i]
    aligned = np.repeat(is_aligned[None, :].astype(np.uint8), 4, axis=0)
    return lines, aligned, gravity, poses

"""Minimal in-memory incremental mapper over the library's operators.

Follows the control flow of ``IncrementalMapper`` / ``IncrementalMapperController``
(src/sfm/incremental_mapper.cc:192-1160, src/controllers/incremental_mapper.cc:382-591) on an
in-memory scene — no database, no images, no correspondence search: the correspondence graph is
given as tracks (one lifted line per image that sees the point):

  RegisterInitialLineImages   four images with gravity -> init.initialize_reconstruction
                              (sfm/incremental_mapper.cc:192-435)
  RegisterNextImage           2D-3D from the tracks -> EstimateAbsolutePoseFromLines (P6L RANSAC
                              on the GPU) -> RefineAbsolutePoseFromLines      (:570-760)
  TriangulateImage            every track with >= 3 registered views and no point yet ->
                              EstimateTriangulationBatch (one GPU call per image)
                              (sfm/incremental_triangulator.cc:468-560)
  AdjustGlobalBundle          FilterObservationsWithNegativeDepth, BundleAdjuster (first image
                              constant, second image tvec[0] constant), then FilterPoints3D
                              (sfm/incremental_mapper.cc:893-945, controllers/...:102-130)

It is a driver for tests and for the full-loop configuration of BASELINE.json, not a
re-implementation of COLMAP's bookkeeping (no local BA, re-triangulation or track merging).
"""
import numpy as np

from . import bundle_adjustment as ba
from . import filters as F
from . import initializer as I
from . import triangulation as T
from .binding import RANSACOptions
from .estimators import EstimateAbsolutePoseFromLines


class Scene:
    """Tracks of lifted lines.  lines[i][p] = (a, b, c) of point p in image i (normalised camera
    coordinates) or NaN where image i does not see p; aligned[p] marks gravity-aligned tracks;
    gravity[i] is the gravity direction in image i; one shared camera (model id, params, size)."""

    def __init__(self, lines, aligned, gravity, camera_model, camera_params, camera_size):
        self.lines = np.asarray(lines, np.float64)          # [N images, P points, 3]
        self.visible = ~np.isnan(self.lines[:, :, 0])
        self.aligned = np.asarray(aligned, bool)            # [P]
        self.gravity = np.asarray(gravity, np.float64)      # [N, 3]
        self.camera_model, self.camera_params = int(camera_model), list(camera_params)
        self.camera_size = tuple(camera_size)
        self.mean_focal = float(np.mean(camera_params[:2] if camera_model in (1, 4) else camera_params[:1]))


class IncrementalMapper:
    def __init__(self, ctx, scene, max_reproj_error_px=12.0, filter_max_reproj_error=4.0,
                 filter_min_tri_angle=1.5, ba_every=4, verbose=False):
        self.ctx, self.scene = ctx, scene
        n, p = scene.visible.shape
        self.qvec = np.zeros((n, 4))
        self.tvec = np.zeros((n, 3))
        self.registered = []                     # image indices in registration order
        self.points = np.full((p, 3), np.nan)
        self.has_point = np.zeros(p, bool)
        self.obs_on = scene.visible.copy()       # observation (i, p) still part of its track
        self.max_reproj_error_px = max_reproj_error_px
        self.filter_max_reproj_error, self.filter_min_tri_angle = filter_max_reproj_error, filter_min_tri_angle
        self.ba_every, self.verbose = ba_every, verbose
        self.log = []

    # ---- RegisterInitialLineImages --------------------------------------------------------
    def register_initial(self, image_ids):
        sc = self.scene
        ids = list(image_ids)
        common = np.flatnonzero(sc.visible[ids].all(axis=0))
        lines = sc.lines[ids][:, common]
        aligned = np.repeat(sc.aligned[common][None, :].astype(np.uint8), 4, axis=0)
        ok, poses, ratio, rep = I.initialize_reconstruction(lines, aligned, sc.gravity[ids])
        if not ok:
            return False
        for k, i in enumerate(ids):
            R, t = poses[k, :, :3], poses[k, :, 3]
            self.qvec[i], self.tvec[i] = _rotmat_to_quat(R), t
            self.registered.append(i)
        self.log.append(("init", ids, ratio))
        self.triangulate_new()
        self.adjust_global_bundle()
        return True

    # ---- RegisterNextImage ------------------------------------------------------------------
    def find_next_image(self):
        sc = self.scene
        score = (self.obs_on & self.has_point[None, :]).sum(axis=1).astype(float)
        score[self.registered] = -1
        i = int(np.argmax(score))
        return i if score[i] >= 6 else None

    def register_next_image(self, i):
        sc = self.scene
        pts = np.flatnonzero(self.obs_on[i] & self.has_point)
        lines2d, aligned, points3d = sc.lines[i, pts], sc.aligned[pts], self.points[pts]
        # sfm/incremental_mapper.cc:673-681
        opt = RANSACOptions(max_error=self.max_reproj_error_px / sc.mean_focal, min_inlier_ratio=0.25,
                            confidence=0.99999, min_num_trials=100, max_num_trials=10000)
        ok, q, t, num_inliers, mask = EstimateAbsolutePoseFromLines(
            opt, (lines2d, aligned.astype(np.uint8)), points3d, ctx=self.ctx)
        if not ok or num_inliers < 15:           # abs_pose_min_num_inliers is 30 in the reference
            return False
        cam = ba.Camera(1, sc.camera_model, sc.camera_params)
        q, t = np.array(q, np.float64), np.array(t, np.float64)
        ropt = ba.AbsolutePoseRefinementOptions()
        if not ba.RefineAbsolutePoseFromLines(ropt, mask, lines2d, points3d, q, t, cam, ctx=self.ctx):
            return False
        self.qvec[i], self.tvec[i] = q, t
        self.registered.append(i)
        # only the RANSAC inliers continue their tracks (:739-752)
        self.obs_on[i, pts[~np.asarray(mask, bool)]] = False
        self.log.append(("register", i, int(num_inliers), len(pts)))
        return True

    # ---- TriangulateImage -------------------------------------------------------------------
    def _track_problem(self, point_ids, use_points=None):
        sc = self.scene
        reg = np.array(self.registered)
        vis = self.obs_on[reg][:, point_ids]                        # [R, T]
        counts = vis.sum(axis=0)
        track_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        r_idx, t_idx = np.nonzero(vis.T)[::-1]                      # track-major order
        order = np.lexsort((r_idx, t_idx))
        r_idx, t_idx = r_idx[order], t_idx[order]
        obs_image = reg[r_idx]
        obs_line = sc.lines[obs_image, point_ids[t_idx]]
        obs_aligned = sc.aligned[point_ids[t_idx]].astype(np.uint8)
        pts = np.zeros((len(point_ids), 3)) if use_points is None else use_points
        pb = F.FilterProblem(self.qvec, self.tvec, np.zeros(len(self.qvec), np.int32),
                             [sc.camera_model], [sc.camera_params], [sc.camera_size], pts,
                             track_start, obs_image, obs_line, obs_aligned)
        return pb, obs_image, point_ids[t_idx]

    def triangulate_new(self):
        reg = np.array(self.registered)
        views = self.obs_on[reg].sum(axis=0)
        # tracks need a non-aligned line (incremental_triangulator.cc:512-515) and >= 3 views
        cand = np.flatnonzero(~self.has_point & (views >= 3) & ~self.scene.aligned)
        if len(cand) == 0:
            return 0
        pb, _, _ = self._track_problem(cand)
        opt = T.EstimateTriangulationOptions(
            min_tri_angle=np.deg2rad(1.5), residual_type=T.ANGULAR_ERROR, max_error=np.deg2rad(2.0),
            confidence=0.9999, min_inlier_ratio=0.02, max_num_trials=10000, exhaustive_threshold=15)
        ok, xyz, mask, _ = T.EstimateTriangulationBatch(self.ctx, pb, opt)
        self.points[cand[ok]] = xyz[ok]
        self.has_point[cand[ok]] = True
        self.log.append(("triangulate", int(ok.sum()), len(cand)))
        return int(ok.sum())

    # ---- AdjustGlobalBundle + filters ----------------------------------------------------------
    def adjust_global_bundle(self, max_num_iterations=50):
        sc = self.scene
        reg = np.array(self.registered)
        pid = np.flatnonzero(self.has_point)
        if len(pid) == 0:
            return
        pb, obs_image, obs_point = self._track_problem(pid, self.points[pid].copy())
        # FilterObservationsWithNegativeDepth before the adjustment (:904)
        # (DeleteObservation removes points whose track is down to three views: all their
        # observations come back flagged)
        _, neg, dead = F.FilterObservationsWithNegativeDepth(self.ctx, pb)
        self.obs_on[obs_image[neg.astype(bool)], obs_point[neg.astype(bool)]] = False
        self.has_point[pid[dead.astype(bool)]] = False
        self.points[pid[dead.astype(bool)]] = np.nan
        keep = ~neg.astype(bool)
        pid = pid[~dead.astype(bool)]
        if len(pid) == 0:
            return
        local_pt = np.searchsorted(pid, obs_point[keep])
        flags = np.ones(len(self.qvec), np.uint8)        # unregistered images: constant, unused
        flags[reg] = 0
        flags[reg[0]] = 1                                 # first image fixed (:907-926)
        flags[reg[1]] = 2                                 # second image: tvec[0] fixed
        arrays = ba.BaArrays(self.qvec, self.tvec, self.points[pid], obs_image[keep], local_pt,
                             sc.lines[obs_image[keep], obs_point[keep]], [sc.camera_model],
                             [sc.camera_params], pose_flags=flags)
        opts = ba.default_solver_options(loss_type=0, max_num_iterations=max_num_iterations,
                                         gradient_tolerance=1.0)   # controllers/...:221-243
        ok, s = ba.solve_arrays(self.ctx, arrays, opts)
        if ok:
            self.qvec[reg], self.tvec[reg] = arrays.qvecs[reg], arrays.tvecs[reg]
            self.points[pid] = arrays.points
        # FilterPoints3D after the adjustment (controllers/incremental_mapper.cc:120-128)
        pb, obs_image, obs_point = self._track_problem(pid, self.points[pid].copy())
        nf, od, pd, _ = F.FilterPoints3D(self.ctx, pb, self.filter_max_reproj_error,
                                         self.filter_min_tri_angle)
        od, pd = od.astype(bool), pd.astype(bool)
        self.obs_on[obs_image[od], obs_point[od]] = False
        self.has_point[pid[pd]] = False
        self.points[pid[pd]] = np.nan
        self.log.append(("global_ba", float(s.initial_cost), float(s.final_cost), int(nf)))

    # ---- the loop (controllers/incremental_mapper.cc:438-591) -------------------------------------
    def run(self, initial_images):
        if not self.register_initial(initial_images):
            return False
        since_ba = 0
        while True:
            i = self.find_next_image()
            if i is None:
                break
            if not self.register_next_image(i):
                # give up on this image (the reference retries other candidates)
                self.obs_on[i, :] = False
                continue
            self.triangulate_new()
            since_ba += 1
            if since_ba >= self.ba_every:
                self.adjust_global_bundle()
                self.triangulate_new()
                since_ba = 0
        self.adjust_global_bundle()
        return True


def _rotmat_to_quat(R):
    from .synthetic import rotmat_to_quat
    q = rotmat_to_quat(np.asarray(R))
    return q / np.linalg.norm(q)


def make_mapper_scene(num_images=12, num_points=600, aligned_fraction=0.4, noise_px=0.3,
                      focal=1000.0, seed=1, visibility=0.8):
    """Upright cameras on an arc looking at a cloud of points; every point is seen by a random
    subset of the images; the four initial images see everything.  Returns (Scene, gt) with
    gt = dict(R [N,3,3], t [N,3], points [P,3])."""
    rng = np.random.default_rng(seed)
    ang = np.linspace(-0.9, 0.9, num_images) + 0.03 * rng.normal(size=num_images)
    R = np.zeros((num_images, 3, 3))
    t = np.zeros((num_images, 3))
    for i, a in enumerate(ang):
        c = np.array([4.0 * np.sin(a), 0.15 * rng.normal(), -4.0 * np.cos(a)])   # camera centre
        yaw = a + 0.05 * rng.normal()
        R[i] = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
        t[i] = -R[i] @ c
    X = rng.uniform(-1, 1, (num_points, 3)) * np.array([1.2, 0.8, 1.2])
    lines = np.full((num_images, num_points, 3), np.nan)
    aligned = rng.uniform(size=num_points) < aligned_fraction
    vis = rng.uniform(size=(num_images, num_points)) < visibility
    vis[:4] = True
    for i in range(num_images):
        pc = X @ R[i].T + t[i]
        uv = pc[:, :2] / pc[:, 2:3] + rng.normal(scale=noise_px / focal, size=(num_points, 2))
        g = R[i][:, 1]
        xh = np.concatenate([uv, np.ones((num_points, 1))], axis=1)
        th = rng.uniform(0, 2 * np.pi, num_points)
        n_rand = np.stack([np.cos(th), np.sin(th), np.zeros(num_points)], axis=1)
        l_al = np.cross(xh, g[None, :])
        a, b = np.cos(th), np.sin(th)
        l_rand = np.stack([a, b, -(a * uv[:, 0] + b * uv[:, 1])], axis=1)
        l = np.where(aligned[:, None], l_al, l_rand)
        l /= np.linalg.norm(l[:, :2], axis=1, keepdims=True)
        ok = vis[i] & (pc[:, 2] > 0.5)
        lines[i, ok] = l[ok]
    gravity = np.stack([R[i][:, 1] for i in range(num_images)])
    scene = Scene(lines, aligned, gravity, 1, [focal, focal, 500.0, 500.0], (1000, 1000))
    return scene, dict(R=R, t=t, points=X)


def pose_errors(mapper, gt):
    """Align the reconstruction to the ground truth (similarity from the camera centres) and
    return (max rotation error [rad], max centre error relative to the scene extent)."""
    from .synthetic import quat_to_rotmat
    reg = np.array(mapper.registered)
    Re = np.stack([quat_to_rotmat(mapper.qvec[i]) for i in reg])
    ce = np.stack([-Re[k].T @ mapper.tvec[i] for k, i in enumerate(reg)])
    cg = np.stack([-gt["R"][i].T @ gt["t"][i] for i in reg])
    mu_e, mu_g = ce.mean(0), cg.mean(0)
    H = (ce - mu_e).T @ (cg - mu_g)
    U, S, Vt = np.linalg.svd(H)
    D = np.diag([1, 1, np.sign(np.linalg.det(Vt.T @ U.T))])
    Rot = Vt.T @ D @ U.T
    scale = np.trace(np.diag(S) @ D) / ((ce - mu_e) ** 2).sum()
    c_al = (scale * (Rot @ (ce - mu_e).T)).T + mu_g
    extent = np.linalg.norm(cg.max(0) - cg.min(0))
    rot_err = 0.0
    for k, i in enumerate(reg):
        Rg = gt["R"][i]
        dR = Re[k] @ Rot.T @ Rg.T
        rot_err = max(rot_err, float(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))))
    return rot_err, float(np.linalg.norm(c_al - cg, axis=1).max() / extent)

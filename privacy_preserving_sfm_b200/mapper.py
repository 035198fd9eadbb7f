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

  AdjustLocalBundle           the new image + the 5 images sharing most points, last one constant,
                              second-to-last tvec[0] constant, short-track points of the new image
                              variable, their views outside the bundle through constant poses,
                              SOFT_L1, then CompleteTracks + FilterPoints3D of the touched points
                              (sfm/incremental_mapper.cc:781-891)
  schedule                    local BA after every registration, global BA when the model has
                              grown by ba_global_{images,points}_{ratio,freq}
                              (controllers/incremental_mapper.cc:490-510)
  WriteText / ReadText        cameras.txt / images.txt / points3D.txt of this fork
                              (base/reconstruction.cc:963-1095: images carry LINES2D[] as
                              (A, B, C, is_aligned, POINT3D_ID))

It is a driver for tests and for the full-loop configuration of BASELINE.json (configs[4]) over the
GPU operators, not a re-implementation of COLMAP's bookkeeping: tracks are given (no
correspondence search, hence no track merging), one shared camera, intrinsics constant (the
fork's defaults, controllers/incremental_mapper.h:81-83).
"""
import numpy as np

from . import bundle_adjustment as ba
from . import filters as F
from . import initializer as I
from . import model_io
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
        self.mean_focal = float(np.mean(camera_params[:2] if camera_model in (1, 4, 5, 6, 7, 10)
                                        else camera_params[:1]))

    @classmethod
    def from_correspondence_graph(cls, graph, image_lines, image_aligned, gravity, camera_model,
                                  camera_params, camera_size, min_track_length=2):
        """The scene of a matched image set: ``graph`` is a correspondence_graph.CorrespondenceGraph
        over image ids index + 1 (what ``CorrespondenceGraph`` holds after the database's matches
        were added, controllers/incremental_mapper.cc:424-436 / base/database_cache.cc),
        ``image_lines[i]`` [n_i, 3] and ``image_aligned[i]`` [n_i] the lifted lines of image i.
        Tracks are the connected components of the graph (``graph.Tracks``); a component with two
        lines of one image is ambiguous and left out, a track must be aligned in all of its views
        or in none (one flag per track in this driver).  Returns (scene, tracks)."""
        tracks = [t for t in graph.Tracks(min_track_length)
                  if len({image_id for image_id, _ in t}) == len(t)]
        lines = np.full((len(image_lines), len(tracks), 3), np.nan)
        aligned = np.zeros(len(tracks), bool)
        for k, t in enumerate(tracks):
            flags = [bool(image_aligned[image_id - 1][line_idx]) for image_id, line_idx in t]
            if any(flags) != all(flags):
                raise ValueError("track %d mixes gravity-aligned and free lines" % k)
            aligned[k] = flags[0]
            for image_id, line_idx in t:
                lines[image_id - 1, k] = image_lines[image_id - 1][line_idx]
        return cls(lines, aligned, gravity, camera_model, camera_params, camera_size), tracks


class IncrementalMapper:
    def __init__(self, ctx, scene, max_reproj_error_px=12.0, filter_max_reproj_error=4.0,
                 filter_min_tri_angle=1.5, ba_every=4, verbose=False, local_ba=False,
                 ba_global_images_ratio=1.1, ba_global_points_ratio=1.1,
                 ba_global_images_freq=500, ba_global_points_freq=250000,
                 ba_local_num_images=6, ba_ctx=None, normalize=False,
                 local_bundle_selection="overlap", ba_local_min_tri_angle=6.0):
        """local_ba = False: global BA every `ba_every` images (the small-scene mode of the
        tests); True: the controller's schedule — local BA after every image, global BA when the
        model has grown by the ba_global_* ratios.  ba_ctx: context for the GLOBAL adjustments
        (e.g. one with a communicator: sharded over the GPUs); default ctx."""
        self.local_ba = local_ba
        self.max_init_tracks = 2000
        self.ba_global = (ba_global_images_ratio, ba_global_points_ratio, ba_global_images_freq,
                          ba_global_points_freq)
        self.ba_local_num_images = ba_local_num_images
        self.ba_ctx = ba_ctx if ba_ctx is not None else ctx
        self.timing = {}
        self.ctx, self.scene = ctx, scene
        n, p = scene.visible.shape
        self.qvec = np.zeros((n, 4))
        self.tvec = np.zeros((n, 3))
        self.registered = []                     # image indices in registration order
        self.points = np.full((p, 3), np.nan)
        self.has_point = np.zeros(p, bool)
        self.obs_on = scene.visible.copy()       # observation (i, p) still part of its track
        # sparse view of the (static) visibility: the images that see point p, ascending
        pi, ii = np.nonzero(scene.visible.T)
        self._csc_ptr = np.concatenate([[0], np.cumsum(np.bincount(pi, minlength=p))]).astype(np.int64)
        self._csc_img = ii.astype(np.int64)
        # find_next_image: visible-point counts of the unregistered images, kept up to date from
        # the changes of has_point (their obs_on rows do not change before they are registered)
        self._score = np.zeros(n, np.int64)
        self._score_has_point = np.zeros(p, bool)
        self.max_reproj_error_px = max_reproj_error_px
        self.filter_max_reproj_error, self.filter_min_tri_angle = filter_max_reproj_error, filter_min_tri_angle
        self.ba_every, self.verbose = ba_every, verbose
        # Reconstruction::Normalize after every global adjustment (sfm/incremental_mapper.cc:934-936):
        # off by default — a similarity of the whole model, it changes no residual, only the gauge
        # the result is reported in (pose_errors aligns by a similarity anyway)
        self.normalize = normalize
        # "overlap": the images sharing most points with the new one (the mode every measured run
        # used); "reference": IncrementalMapper::FindLocalBundle in full (find_local_bundle)
        self.local_bundle_selection = local_bundle_selection
        self.ba_local_min_tri_angle = ba_local_min_tri_angle
        self.log = []

    # ---- RegisterInitialLineImages --------------------------------------------------------
    def register_initial(self, image_ids):
        sc = self.scene
        ids = list(image_ids)
        common = np.flatnonzero(sc.visible[ids].all(axis=0))
        # (the LO-MSAC control flow is host code, its candidate models are scored on the GPU —
        # every model carries every track; a few thousand tracks are what a real
        # four-view match set holds)
        common = common[:self.max_init_tracks]
        lines = sc.lines[ids][:, common]
        aligned = np.repeat(sc.aligned[common][None, :].astype(np.uint8), 4, axis=0)
        ok, poses, ratio, rep = I.initialize_reconstruction(lines, aligned, sc.gravity[ids],
                                                            ctx=self.ctx)
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
        # score[i] = number of triangulated points image i sees; only needed for unregistered
        # images, whose rows of obs_on still equal the scene's visibility: update the counts by
        # the points whose has_point flag changed since the last call
        changed = np.flatnonzero(self.has_point != self._score_has_point)
        if len(changed):
            gained = changed[self.has_point[changed]]
            lost = changed[~self.has_point[changed]]
            if len(gained):
                self._score += self.scene.visible[:, gained].sum(axis=1)
            if len(lost):
                self._score -= self.scene.visible[:, lost].sum(axis=1)
            self._score_has_point[changed] = self.has_point[changed]
        score = self._score.astype(float)
        score[self.registered] = -1
        i = int(np.argmax(score))
        return i if score[i] >= 6 else None

    def _views(self, point_ids):
        """Active observations of the given points from registered images: (image, index into
        point_ids), track-major, the views of a track in registration order."""
        n = len(self.qvec)
        rank = np.full(n, -1, np.int64)
        rank[self.registered] = np.arange(len(self.registered))
        point_ids = np.asarray(point_ids, np.int64)
        starts = self._csc_ptr[point_ids]
        lens = self._csc_ptr[point_ids + 1] - starts
        total = int(lens.sum())
        t_idx = np.repeat(np.arange(len(point_ids)), lens)
        offs = np.arange(total) - np.repeat(np.cumsum(lens) - lens, lens) + np.repeat(starts, lens)
        img = self._csc_img[offs]
        keep = (rank[img] >= 0) & self.obs_on[img, point_ids[t_idx]]
        img, t_idx = img[keep], t_idx[keep]
        order = np.lexsort((rank[img], t_idx))
        return img[order], t_idx[order]

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
        point_ids = np.asarray(point_ids, np.int64)
        obs_image, t_idx = self._views(point_ids)                   # track-major order
        counts = np.bincount(t_idx, minlength=len(point_ids))
        track_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        obs_line = sc.lines[obs_image, point_ids[t_idx]]
        obs_aligned = sc.aligned[point_ids[t_idx]].astype(np.uint8)
        pts = np.zeros((len(point_ids), 3)) if use_points is None else use_points
        pb = F.FilterProblem(self.qvec, self.tvec, np.zeros(len(self.qvec), np.int32),
                             [sc.camera_model], [sc.camera_params], [sc.camera_size], pts,
                             track_start, obs_image, obs_line, obs_aligned)
        return pb, obs_image, point_ids[t_idx]

    def triangulate_new(self):
        reg = np.array(self.registered)
        # tracks need a non-aligned line (incremental_triangulator.cc:512-515) and >= 3 views
        open_pts = np.flatnonzero(~self.has_point & ~self.scene.aligned)
        views = np.bincount(self._views(open_pts)[1], minlength=len(open_pts))
        cand = open_pts[views >= 3]
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
        ok, s = ba.solve_arrays(self.ba_ctx, arrays, opts)
        if ok:
            self.qvec[reg], self.tvec[reg] = arrays.qvecs[reg], arrays.tvecs[reg]
            self.points[pid] = arrays.points
        if self.normalize:
            self.normalize_scene()
        # FilterPoints3D after the adjustment (controllers/incremental_mapper.cc:120-128)
        pb, obs_image, obs_point = self._track_problem(pid, self.points[pid].copy())
        nf, od, pd, _ = F.FilterPoints3D(self.ctx, pb, self.filter_max_reproj_error,
                                         self.filter_min_tri_angle)
        od, pd = od.astype(bool), pd.astype(bool)
        self.obs_on[obs_image[od], obs_point[od]] = False
        self.has_point[pid[pd]] = False
        self.points[pid[pd]] = np.nan
        self.log.append(("global_ba", float(s.initial_cost), float(s.final_cost), int(nf)))

    def normalize_scene(self):
        """Reconstruction::Normalize (base/reconstruction.cc:302-398; model_io.normalize_scene):
        registered images' centres to extent 10 around their robust mean, points with them."""
        reg = np.array(self.registered)
        pid = np.flatnonzero(self.has_point)
        tvec, pts, scale, _ = model_io.normalize_scene(self.qvec[reg], self.tvec[reg], self.points[pid])
        self.tvec[reg], self.points[pid] = tvec, pts
        return scale

    def find_local_bundle(self, i):
        """IncrementalMapper::FindLocalBundle (sfm/incremental_mapper.cc:993-1160): the registered
        images sharing points with image i by descending number of shared observations; if there
        are more than ba_local_num_images - 1 of them, those whose triangulation angle towards
        image i (75th percentile over image i's points, base/triangulation.cc:84-118) clears a
        threshold are preferred, over eight successively relaxed (angle, overlap) thresholds, and
        the rest is filled with the most overlapping ones.  Returns image indices in selection
        order.  (Ties of the overlap count, which the reference's unordered map + std::sort leave
        unspecified, are taken in registration order.)"""
        seen_idx = np.flatnonzero(self.obs_on[i] & self.has_point)
        if len(seen_idx) == 0:
            return []
        img_s, _ = self._views(seen_idx)
        reg = np.array(self.registered)
        shared = np.bincount(img_s, minlength=len(self.qvec))[reg]
        shared[reg == i] = 0
        order = np.argsort(-shared, kind="stable")
        order = order[shared[order] > 0]
        overlapping, counts = reg[order], shared[order]
        num_eff = min(self.ba_local_num_images - 1, len(overlapping))
        if len(overlapping) == num_eff:
            return [int(v) for v in overlapping]
        min_tri = np.deg2rad(self.ba_local_min_tri_angle)
        n_pts = float(len(seen_idx))
        thresholds = [(min_tri / 1.0, 0.6 * n_pts), (min_tri / 1.5, 0.6 * n_pts), (min_tri / 2.0, 0.5 * n_pts),
                      (min_tri / 2.5, 0.4 * n_pts), (min_tri / 3.0, 0.3 * n_pts), (min_tri / 4.0, 0.2 * n_pts),
                      (min_tri / 5.0, 0.1 * n_pts), (min_tri / 6.0, 0.1 * n_pts)]
        centre = model_io.projection_centers(self.qvec[[i]], self.tvec[[i]])[0]
        pts = self.points[seen_idx]
        ray1 = ((pts - centre) ** 2).sum(axis=1)
        tri_angle = np.full(len(overlapping), -1.0)
        used = np.zeros(len(overlapping), bool)
        local = []
        for angle_thr, overlap_thr in thresholds:
            for k in range(len(overlapping)):
                if counts[k] < overlap_thr:
                    break
                if used[k]:
                    continue
                if tri_angle[k] < 0.0:
                    c2 = model_io.projection_centers(self.qvec[[overlapping[k]]], self.tvec[[overlapping[k]]])[0]
                    ray2 = ((pts - c2) ** 2).sum(axis=1)
                    den = 2.0 * np.sqrt(ray1 * ray2)
                    with np.errstate(invalid="ignore", divide="ignore"):
                        ang = np.abs(np.arccos((ray1 + ray2 - ((centre - c2) ** 2).sum()) / den))
                    ang = np.where(den == 0.0, 0.0, np.minimum(ang, np.pi - ang))
                    idx = max(0, min(len(ang) - 1, int(round(75 / 100 * (len(ang) - 1)))))
                    tri_angle[k] = np.partition(ang, idx)[idx]
                if tri_angle[k] >= angle_thr:
                    local.append(int(overlapping[k]))
                    used[k] = True
                    if len(local) >= num_eff:
                        break
            if len(local) >= num_eff:
                break
        for k in range(len(overlapping)):
            if len(local) >= num_eff:
                break
            if not used[k]:
                local.append(int(overlapping[k]))
                used[k] = True
        return local

    # ---- AdjustLocalBundle (sfm/incremental_mapper.cc:781-891) ----------------------------------
    def adjust_local_bundle(self, i, max_num_iterations=25):
        sc = self.scene
        reg = np.array(self.registered)
        seen = self.obs_on[i] & self.has_point                    # points of the new image
        if not seen.any():
            return
        # FindLocalBundle: the images sharing most points with image i
        seen_idx = np.flatnonzero(seen)
        img_s, t_s = self._views(seen_idx)                        # registered views of its points
        shared_all = np.bincount(img_s, minlength=len(self.qvec))
        shared = shared_all[reg]
        shared[reg == i] = -1
        order = np.argsort(-shared, kind="stable")[:self.ba_local_num_images - 1]
        local = [int(reg[k]) for k in order if shared[k] > 0]
        if self.local_bundle_selection == "reference":
            local = self.find_local_bundle(i)
        if not local:
            return
        bundle = [i] + local
        in_bundle = np.zeros(len(self.qvec), bool)
        in_bundle[bundle] = True
        # variable points: those of the new image with a short track (kMaxTrackLength = 15)
        variable = np.zeros(len(seen), bool)
        track_len = np.bincount(t_s, minlength=len(seen_idx))
        variable[seen_idx[track_len <= 15]] = True
        # observations: everything the bundle images see (AddImageToProblem), plus the views of
        # the variable points from registered images outside the bundle (AddPointToProblem);
        # ordered by (image, point)
        oi, op = [], []
        for b_img in bundle:
            pts_b = np.flatnonzero(self.obs_on[b_img] & self.has_point)
            oi.append(np.full(len(pts_b), b_img, np.int64))
            op.append(pts_b)
        out = variable[seen_idx[t_s]] & ~in_bundle[img_s]
        oi.append(img_s[out])
        op.append(seen_idx[t_s[out]])
        obs_image, obs_point = np.concatenate(oi), np.concatenate(op)
        order = np.lexsort((obs_point, obs_image))
        obs_image, obs_point = obs_image[order], obs_point[order]
        pts_any = np.unique(np.concatenate(op[:len(bundle)]))
        flags = np.ones(len(self.qvec), np.uint8)                 # outside the bundle: constant
        flags[bundle] = 0
        if len(local) == 1:                                       # (:828-839)
            flags[local[0]] = 1
            flags[i] = 2
        else:
            flags[local[-1]] = 1
            flags[local[-2]] = 2
        local_pt = np.searchsorted(pts_any, obs_point)
        # points whose track is not completely inside the problem are constant (ParameterizePoints)
        arrays = ba.BaArrays(self.qvec, self.tvec, self.points[pts_any], obs_image, local_pt,
                             sc.lines[obs_image, obs_point], [sc.camera_model],
                             [sc.camera_params], pose_flags=flags,
                             point_const=(~variable[pts_any]).astype(np.uint8))
        opts = ba.default_solver_options(loss_type=1, loss_scale=1.0,   # controllers/...:196-219
                                         max_num_iterations=max_num_iterations,
                                         gradient_tolerance=10.0)
        ok, s = ba.solve_arrays(self.ctx, arrays, opts)
        if ok:
            self.qvec[bundle], self.tvec[bundle] = arrays.qvecs[bundle], arrays.tvecs[bundle]
            self.points[pts_any] = arrays.points
        touched = np.flatnonzero(variable)
        completed = self.complete_tracks(touched)
        # FilterPoints3D of the changed points (:882-888)
        nf = 0
        if len(touched):
            pb, obs_image, obs_point = self._track_problem(touched, self.points[touched].copy())
            nf, od, pd, _ = F.FilterPoints3D(self.ctx, pb, self.filter_max_reproj_error,
                                             self.filter_min_tri_angle)
            od, pd = od.astype(bool), pd.astype(bool)
            self.obs_on[obs_image[od], obs_point[od]] = False
            self.has_point[touched[pd]] = False
            self.points[touched[pd]] = np.nan
        self.log.append(("local_ba", i, len(bundle), float(s.initial_cost), float(s.final_cost),
                         int(completed), int(nf)))

    # ---- CompleteTracks (sfm/incremental_triangulator.cc, CompleteTracks / CompleteImage) --------
    def complete_tracks(self, point_ids):
        """Views of a triangulated point that were dropped earlier (RANSAC outlier at registration
        time, filtered) re-join the track if the point now reprojects within the filter
        threshold and in front of the camera."""
        sc = self.scene
        reg = np.array(self.registered)
        if len(point_ids) == 0:
            return 0
        cand = sc.visible[np.ix_(reg, point_ids)] & ~self.obs_on[np.ix_(reg, point_ids)] & \
            self.has_point[point_ids][None, :]
        ri, pi = np.nonzero(cand)
        if len(ri) == 0:
            return 0
        img, pts = reg[ri], point_ids[pi]
        q = self.qvec[img] / np.linalg.norm(self.qvec[img], axis=1, keepdims=True)
        w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        Rm = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], 1),
                       np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], 1),
                       np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1)],
                      axis=1)
        pc = np.einsum("nij,nj->ni", Rm, self.points[pts]) + self.tvec[img]
        ok = pc[:, 2] > 1e-9
        uv = pc[:, :2] / np.where(ok, pc[:, 2], 1.0)[:, None]
        l = sc.lines[img, pts]
        dist_px = np.abs((l[:, 0] * uv[:, 0] + l[:, 1] * uv[:, 1] + l[:, 2])) * sc.mean_focal
        good = ok & (dist_px < self.filter_max_reproj_error)
        self.obs_on[img[good], pts[good]] = True
        return int(good.sum())

    # ---- WriteText / ReadText (base/reconstruction.cc:543-553, 721-1095; model_io.py) -------------
    def to_model(self):
        """The reconstruction as model_io.Model: image ids are index + 1, the lines of an image are
        its visible tracks in point order, point ids are point index + 1."""
        sc = self.scene
        nparams = model_io.CAMERA_MODEL_NUM_PARAMS[sc.camera_model]
        cams = {1: model_io.Camera(sc.camera_model, sc.camera_size[0], sc.camera_size[1],
                                   sc.camera_params[:nparams])}
        reg = sorted(self.registered)
        images = {}
        for i in reg:
            vis = np.flatnonzero(sc.visible[i])
            has = self.obs_on[i, vis] & self.has_point[vis]
            images[i + 1] = model_io.Image(self.qvec[i], self.tvec[i], 1, "image%06d.jpg" % i,
                                           sc.lines[i, vis], sc.aligned[vis],
                                           np.where(has, vis + 1, -1))
        is_reg = np.zeros(len(self.qvec), bool)
        is_reg[reg] = True
        row = np.cumsum(is_reg) - 1                              # image -> row of line_idx
        # line_idx: rank of a point among an image's lines
        line_idx = np.cumsum(sc.visible[reg], axis=1, dtype=np.int32) - 1
        points = {}
        for p in np.flatnonzero(self.has_point):
            imgs = np.flatnonzero(self.obs_on[:, p] & is_reg)
            points[int(p) + 1] = model_io.Point3D(self.points[p],
                                                  np.stack([imgs + 1, line_idx[row[imgs], p]], 1))
        return model_io.Model(cams, images, points)

    def write_text(self, path):
        """cameras.txt, images.txt (LINES2D[] as (A, B, C, is_aligned, POINT3D_ID)) and
        points3D.txt (TRACK[] as (IMAGE_ID, line_idx)) in the record layout of the reference's
        WriteText, with 17 significant digits everywhere (the reference itself keeps six digits of
        poses, lines and camera parameters: model_io.write_model_text)."""
        model_io.write_model_text(path, self.to_model(), reference_precision=False)

    @staticmethod
    def read_text(path):
        """Reads the three files back WITHOUT the float narrowing of the reference's reader
        (model_io.read_model_text(path, reference_precision=False)): dict(cameras {id: (model, w,
        h, params)}, images {id: (qvec, tvec, camera_id, name, lines [n, 5])}, points {id: (xyz,
        error, track [m, 2])})."""
        m = model_io.read_model_text(path, reference_precision=False)
        cams = {c: (cam.model_name, cam.width, cam.height, cam.params) for c, cam in m.cameras.items()}
        images = {i: (im.qvec, im.tvec, im.camera_id, im.name,
                      np.column_stack([im.lines, im.aligned.astype(np.float64),
                                       im.point3D_ids.astype(np.float64)]).reshape(-1, 5))
                  for i, im in m.images.items()}
        points = {p: (pt.xyz, pt.error, pt.track) for p, pt in m.points3D.items()}
        return dict(cameras=cams, images=images, points=points)

    # ---- the loop (controllers/incremental_mapper.cc:438-591) -------------------------------------
    def run(self, initial_images, max_images=None):
        import time
        t0 = time.perf_counter()
        tm = self.timing
        for k in ("init", "register", "triangulate", "local_ba", "global_ba", "find_next"):
            tm.setdefault(k, 0.0)

        def timed(key, fn, *a):
            t = time.perf_counter()
            r = fn(*a)
            tm[key] += time.perf_counter() - t
            return r

        if not timed("init", self.register_initial, initial_images):
            return False
        since_ba = 0
        prev_images, prev_points = len(self.registered), int(self.has_point.sum())
        while max_images is None or len(self.registered) < max_images:
            i = timed("find_next", self.find_next_image)
            if i is None:
                break
            if not timed("register", self.register_next_image, i):
                # give up on this image (the reference retries other candidates)
                self.obs_on[i, :] = False
                continue
            timed("triangulate", self.triangulate_new)
            if self.local_ba:
                # controllers/incremental_mapper.cc:484-510: local BA, then global BA when the
                # model has grown enough since the last one
                timed("local_ba", self.adjust_local_bundle, i)
                ri, rp, fi, fp = self.ba_global
                n_img, n_pts = len(self.registered), int(self.has_point.sum())
                if (n_img >= ri * prev_images or n_img >= fi + prev_images or
                        n_pts >= rp * prev_points or n_pts >= fp + prev_points):
                    timed("global_ba", self.adjust_global_bundle)
                    timed("triangulate", self.triangulate_new)
                    prev_images, prev_points = len(self.registered), int(self.has_point.sum())
            else:
                since_ba += 1
                if since_ba >= self.ba_every:
                    timed("global_ba", self.adjust_global_bundle)
                    timed("triangulate", self.triangulate_new)
                    since_ba = 0
        timed("global_ba", self.adjust_global_bundle)
        tm["total"] = time.perf_counter() - t0
        return True


_COMBOS3 = {}


def _combos3(c):
    """All 3-subsets of range(c) in the order of the reference's i < j < k loops."""
    if c not in _COMBOS3:
        i, j, k = np.meshgrid(np.arange(c), np.arange(c), np.arange(c), indexing="ij")
        keep = (i < j) & (j < k)
        _COMBOS3[c] = np.stack([i[keep], j[keep], k[keep]], 1)
    return _COMBOS3[c]


def find_initial_image_sets(graph, image_aligned, check_image_ids, min_num_aligned_tracks=20,
                            min_num_random_tracks=20):
    """The candidate search of ``IncrementalMapper::RegisterInitialLineImages``
    (src/sfm/incremental_mapper.cc:192-421) on a correspondence_graph.CorrespondenceGraph
    (image ids index + 1, ``image_aligned[i]`` the is_aligned flags of image i's lines).

    For every line of every check image (the reference draws up to ten of them at random,
    :296-306: here they are an argument) its correspondences of the SAME alignment are taken; with
    three or more of them every 3-subset plus the line itself is a four-view track, keyed by its
    image set (ids ascending) and stored once (:253-283, :336-358).  An image set qualifies with at
    least 20 aligned and 20 unaligned tracks (:398-412); the qualifying sets are returned by
    descending number of aligned tracks (:421-426: the unaligned weight is zero; ties, which
    ``std::sort`` leaves unspecified, in ascending image-set order).

    Returns [dict(image_set (4 ids), aligned_tracks [na, 4], unaligned_tracks [nu, 4])]: line
    indices per image of the set, rows in lexicographic order (the order of the reference's
    ``std::set``) — aligned rows then unaligned rows are the ``lines`` the reference hands to
    ``init::initialize_reconstruction`` (:459-481)."""
    rows = []
    csr = graph.CSR()
    # is_aligned of the line every correspondence points to (one lookup for the whole graph)
    ids = np.array(sorted(csr["start"]), np.int64)
    sizes = np.array([len(image_aligned[i - 1]) for i in ids], np.int64)
    base = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    flat = np.concatenate([np.asarray(image_aligned[i - 1], bool) for i in ids]) if len(ids) else np.zeros(0, bool)
    dst_aligned = flat[base[np.searchsorted(ids, csr["dst_img"])] + csr["dst_line"]]
    for image_id in check_image_ids:
        flags = np.asarray(image_aligned[image_id - 1], bool)
        start = csr["start"][image_id]                                  # images_.at(image_id)
        for line_idx in range(len(flags)):
            lo, hi = start[line_idx], start[line_idx + 1]
            if hi - lo < 3:
                continue
            same = dst_aligned[lo:hi] == flags[line_idx]
            di, dl = csr["dst_img"][lo:hi][same], csr["dst_line"][lo:hi][same]
            if len(di) < 3:
                continue
            c3 = _combos3(len(di))
            img = np.concatenate([np.full((len(c3), 1), image_id, np.int64), di[c3]], 1)
            idx = np.concatenate([np.full((len(c3), 1), line_idx, np.int64), dl[c3]], 1)
            order = np.argsort(img, axis=1, kind="stable")
            img, idx = np.take_along_axis(img, order, 1), np.take_along_axis(idx, order, 1)
            distinct = (np.diff(img, axis=1) > 0).all(axis=1)          # track_candidate.size() == 4
            al = np.full((int(distinct.sum()), 1), int(flags[line_idx]), np.int64)
            rows.append(np.concatenate([al, img[distinct], idx[distinct]], 1))
    if not rows:
        return []
    rows = np.concatenate(rows)
    rows = rows[np.lexsort(rows.T[::-1])]                               # lexicographic by row
    rows = rows[np.concatenate([[True], (np.diff(rows, axis=0) != 0).any(axis=1)])]   # std::set: every track once
    start = np.flatnonzero(np.concatenate([[True], (np.diff(rows[:, :5], axis=0) != 0).any(axis=1)]))
    stop = np.concatenate([start[1:], [len(rows)]])
    by_set = {}
    for s, e in zip(start.tolist(), stop.tolist()):
        key = rows[s, :5].tolist()
        by_set.setdefault(tuple(key[1:]), {})[key[0]] = rows[s:e, 5:]
    out = []
    for image_set in sorted(by_set):
        tracks = by_set[image_set]
        if len(tracks.get(1, ())) >= min_num_aligned_tracks and len(tracks.get(0, ())) >= min_num_random_tracks:
            out.append(dict(image_set=image_set, aligned_tracks=tracks[1], unaligned_tracks=tracks[0]))
    out.sort(key=lambda d: -len(d["aligned_tracks"]))                  # stable
    return out


def select_initial_images(graph, image_lines, image_aligned, gravity, check_image_ids, options=None,
                          init_min_num_inliers=0, max_num_init_tries=10, ctx=None):
    """The selection loop of ``RegisterInitialLineImages`` (:428-541): the first ten candidate sets
    go through ``init::initialize_reconstruction`` (models scored on the GPU of ``ctx`` if given,
    else the library's host estimators), the one with the best inlier ratio wins; fails without a
    success or with fewer than ``init_min_num_inliers`` inliers (:536-539).

    Returns (ok, image_set, poses [4, 3, 4], inlier_ratio, tried [(image_set, ok, ratio)])."""
    best = (False, None, None, 0.0)
    best_inliers, tried = 0, []
    for cand in find_initial_image_sets(graph, image_aligned, check_image_ids)[:max_num_init_tries]:
        ids = cand["image_set"]
        tracks = np.concatenate([cand["aligned_tracks"], cand["unaligned_tracks"]])
        lines = np.stack([np.asarray(image_lines[i - 1], np.float64)[tracks[:, k]] for k, i in enumerate(ids)])
        aligned = np.zeros((4, len(tracks)), np.uint8)
        aligned[:, :len(cand["aligned_tracks"])] = 1
        ok, poses, ratio, _ = I.initialize_reconstruction(
            lines, aligned, np.asarray(gravity, np.float64)[[i - 1 for i in ids]], options, ctx=ctx)
        tried.append((ids, bool(ok), float(ratio)))
        if ok and ratio > best[3]:
            best = (True, ids, poses, float(ratio))
            best_inliers = int(ratio * len(tracks))
    if not best[0] or best_inliers < init_min_num_inliers:
        return False, None, None, 0.0, tried
    return best + (tried,)


def _rotmat_to_quat(R):
    from .synthetic import rotmat_to_quat
    q = rotmat_to_quat(np.asarray(R))
    return q / np.linalg.norm(q)


def make_mapper_scene(num_images=12, num_points=600, aligned_fraction=0.4, noise_px=0.3,
                      focal=1000.0, seed=1, visibility=0.8, rings=1):
    """Upright cameras on an arc looking at a cloud of points; every point is seen by a random
    subset of the images; the four initial images see everything.  Returns (Scene, gt) with
    gt = dict(R [N,3,3], t [N,3], points [P,3])."""
    rng = np.random.default_rng(seed)
    ang = np.linspace(-0.9, 0.9, num_images) + 0.03 * rng.normal(size=num_images)
    R = np.zeros((num_images, 3, 3))
    t = np.zeros((num_images, 3))
    for i, a in enumerate(ang):
        ring = i % max(1, rings)                 # multi-ring trajectory: radius / height per ring
        rad = 4.0 + 0.8 * ring
        c = np.array([rad * np.sin(a), 0.15 * rng.normal() + 0.5 * ring,
                      -rad * np.cos(a)])                                          # camera centre
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

"""In-memory correspondence graph:
privacy_preserving_sfm_b200/correspondence_graph.py against the REFERENCE'S OWN
CorrespondenceGraph (src/base/correspondence_graph.cc compiled from /root/reference into
oracle/_ref/libref_filter.so, oracle/ref/ref_corr_graph.cc) — CPU.

Both are fed the same pairwise matches (clean tracks, plus duplicates, conflicting matches,
out-of-range indices, self matches, a pair added twice and in both orders, images without any
match) and must answer every query identically: image / pair / observation / correspondence
counts before and after Finalize, the per-pair counts, every line's correspondences IN ORDER, the
transitive closure at every transitivity (the reference's removal of the query line by
overwriting the first entry with the last one included), the matches between two images and the
two-view test.  Then the tracks (connected components) against the generating tracks, a mapper
scene rebuilt from pairwise matches, and the initial image-set search of
RegisterInitialLineImages (src/sfm/incremental_mapper.cc:192-541) against a literal restatement,
with the selection loop run on the library's host estimators.

Skipped where neither oracle/_ref/libref_filter.so nor /root/reference exists."""
import numpy as np
import pytest

from privacy_preserving_sfm_b200.correspondence_graph import CorrespondenceGraph, ImagePairToPairId


@pytest.fixture(scope="module")
def ref():
    import oracle.reference as R
    if not R.available():
        pytest.skip("oracle/_ref/libref_filter.so not built and /root/reference absent")
    return R


def _scene(num_images, num_points, visibility, seed, extra_lines=3):
    """visible [n, p]; per image the line index of every visible point (shuffled) + spare lines."""
    rng = np.random.default_rng(seed)
    visible = rng.uniform(size=(num_images, num_points)) < visibility
    line_of = -np.ones((num_images, num_points), np.int64)
    num_lines = []
    for i in range(num_images):
        vis = np.flatnonzero(visible[i])
        n = len(vis) + extra_lines
        line_of[i, vis] = rng.permutation(n)[:len(vis)]
        num_lines.append(n)
    return visible, line_of, num_lines


def _pairwise(visible, line_of, pair_fraction, seed, keep=0.9):
    """[(id1, id2, matches)] for a random subset of the image pairs, image ids = index + 1."""
    rng = np.random.default_rng(seed)
    n = len(visible)
    calls = []
    for i in range(n):
        for j in range(i + 1, n):
            if rng.uniform() > pair_fraction:
                continue
            both = np.flatnonzero(visible[i] & visible[j])
            both = both[rng.uniform(size=len(both)) < keep]              # the matcher misses some
            m = np.stack([line_of[i, both], line_of[j, both]], 1)
            m = m[rng.permutation(len(m))]
            calls.append((i + 1, j + 1, m) if rng.uniform() < 0.7 else (j + 1, i + 1, m[:, ::-1]))
    return calls


def _feed(graphs, num_lines, calls, finalize=True):
    for g in graphs:
        for i, n in enumerate(num_lines):
            g.AddImage(i + 1, n)
        for id1, id2, m in calls:
            g.AddCorrespondences(id1, id2, m)
        if finalize:
            g.Finalize()


def _same_answers(a, b, num_lines, transitivities=(1, 2, 3, 50)):
    assert a.NumImages() == b.NumImages() and a.NumImagePairs() == b.NumImagePairs()
    assert a.NumCorrespondencesBetweenImages() == b.NumCorrespondencesBetweenImages()
    n = len(num_lines)
    for i in range(1, n + 1):
        assert a.ExistsImage(i) == b.ExistsImage(i)
        if not b.ExistsImage(i):
            with pytest.raises(KeyError):                                  # images_.at(image_id) throws
                a.NumObservationsForImage(i)
            continue
        assert a.NumObservationsForImage(i) == b.NumObservationsForImage(i)
        assert a.NumCorrespondencesForImage(i) == b.NumCorrespondencesForImage(i)
        for line in range(num_lines[i - 1]):
            assert a.HasCorrespondences(i, line) == b.HasCorrespondences(i, line)
            assert a.IsTwoViewObservation(i, line) == b.IsTwoViewObservation(i, line)
            for t in transitivities:
                assert a.FindTransitiveCorrespondences(i, line, t) == \
                    b.FindTransitiveCorrespondences(i, line, t), (i, line, t)
        for j in range(1, n + 1):
            if j != i and b.ExistsImage(j):
                assert a.NumCorrespondencesBetweenImages(i, j) == b.NumCorrespondencesBetweenImages(i, j)
                assert a.FindCorrespondencesBetweenImages(i, j) == b.FindCorrespondencesBetweenImages(i, j)


@pytest.mark.parametrize("seed,pair_fraction", [(1, 1.0), (2, 0.5), (3, 0.25)])
def test_clean_tracks_identical_to_the_reference(ref, seed, pair_fraction):
    visible, line_of, num_lines = _scene(9, 60, 0.5, seed)
    calls = _pairwise(visible, line_of, pair_fraction, seed)
    a, b = CorrespondenceGraph(), ref.CorrespondenceGraph()
    _feed((a, b), num_lines, calls)
    _same_answers(a, b, num_lines)
    assert a.NumImagePairs() == len(calls) and 2 <= a.NumImages() <= 9     # unmatched images are erased


def test_dropped_matches_and_repeated_pairs_identical_to_the_reference(ref, capfd):
    visible, line_of, num_lines = _scene(7, 40, 0.6, seed=11)
    num_lines.append(5)                                                    # image 8: never matched
    calls = _pairwise(visible, line_of, 0.8, seed=11)
    rng = np.random.default_rng(5)
    noisy = []
    for id1, id2, m in calls:
        m = m.copy()
        if len(m) > 4:
            extra = m[rng.integers(0, len(m), 3)].copy()
            extra[1, 1] = m[rng.integers(0, len(m)), 1]                    # another line: conflicts on one side
            extra[2, 0] = num_lines[id1 - 1] + rng.integers(0, 3)          # line index out of range
            m = np.concatenate([m[:2], extra[:1], m[2:], extra[1:]])       # duplicates inside the call
        noisy.append((id1, id2, m))
    noisy.append((3, 3, np.array([[0, 1], [2, 2]])))                       # self matches: ignored
    id1, id2, m = noisy[0]
    noisy.append((id2, id1, m[:, ::-1]))                                   # the same pair again, swapped:
    noisy.append((id1, id2, np.array([[num_lines[id1 - 1] - 1, num_lines[id2 - 1] - 1]])))
    noisy.append((1, 8, np.array([[99, 0], [0, 77]])))                     # a pair with nothing kept
    a, b = CorrespondenceGraph(), ref.CorrespondenceGraph()
    _feed((a, b), num_lines, noisy, finalize=False)
    assert a.NumImages() == b.NumImages() == 8
    assert a.NumCorrespondencesBetweenImages() == b.NumCorrespondencesBetweenImages()
    assert a.NumCorrespondencesBetweenImages(1, 8) == 0 and ImagePairToPairId(8, 1) in a.NumCorrespondencesBetweenImages()
    for i in range(1, 9):                                                  # queries work before Finalize
        assert a.NumCorrespondencesForImage(i) == b.NumCorrespondencesForImage(i)
        for line in range(num_lines[i - 1]):
            assert a.FindCorrespondences(i, line) == b.FindCorrespondences(i, line)
    a.Finalize()
    b.Finalize()
    assert a.NumImages() == b.NumImages() == 7 and not a.ExistsImage(8)   # erased: no observation
    _same_answers(a, b, num_lines)
    assert "Duplicate correspondence" in capfd.readouterr().out            # the reference said so
    with pytest.raises(IndexError):
        a.FindCorrespondences(1, num_lines[0])                             # corrs.at(line_idx)
    with pytest.raises(ValueError):
        a.AddImage(1, 3)                                                   # CHECK(!ExistsImage)


def test_tracks_are_the_generating_tracks():
    visible, line_of, num_lines = _scene(10, 80, 0.45, seed=21)
    calls = _pairwise(visible, line_of, 1.0, seed=21, keep=1.0)          # complete matching
    g = CorrespondenceGraph()
    _feed((g,), num_lines, calls)
    tracks = g.Tracks()
    want = {frozenset((i + 1, int(line_of[i, p])) for i in np.flatnonzero(visible[:, p]))
            for p in range(visible.shape[1]) if visible[:, p].sum() >= 2}
    assert {frozenset(t) for t in tracks} == want and len(tracks) == len(want)
    assert len(g.Tracks(min_length=4)) == sum(len(t) >= 4 for t in tracks)
    assert all(len({i for i, _ in t}) == len(t) for t in tracks)           # one line per image
    assert tracks == sorted(tracks, key=lambda t: t[0]) and all(t == sorted(t) for t in tracks)
    # transitivity: what Find collects from any line of a track is the rest of the track
    for t in tracks[:20]:
        for node in t:
            assert sorted(g.FindTransitiveCorrespondences(*node, 50)) == sorted(x for x in t if x != node)


def test_mapper_scene_from_pairwise_matches():
    """mapper.Scene.from_correspondence_graph: the (image, track) line table the mapper driver
    takes, rebuilt from per-image lines and pairwise matches, equals the generating scene."""
    from privacy_preserving_sfm_b200 import mapper as M
    scene, _ = M.make_mapper_scene(num_images=8, num_points=120, seed=4, visibility=0.6)
    rng = np.random.default_rng(8)
    n, p = scene.visible.shape
    line_of = -np.ones((n, p), np.int64)
    image_lines, image_aligned = [], []
    for i in range(n):
        vis = np.flatnonzero(scene.visible[i])
        perm = rng.permutation(len(vis))
        line_of[i, vis[perm]] = np.arange(len(vis))
        image_lines.append(scene.lines[i, vis[perm]])
        image_aligned.append(scene.aligned[vis[perm]])
    g = CorrespondenceGraph()
    for i in range(n):
        g.AddImage(i + 1, len(image_lines[i]))
    for i in range(n):
        for j in range(i + 1, n):
            both = np.flatnonzero(scene.visible[i] & scene.visible[j])
            g.AddCorrespondences(i + 1, j + 1, np.stack([line_of[i, both], line_of[j, both]], 1))
    g.Finalize()
    rebuilt, tracks = M.Scene.from_correspondence_graph(
        g, image_lines, image_aligned, scene.gravity, scene.camera_model, scene.camera_params,
        scene.camera_size)
    seen = np.flatnonzero(scene.visible.sum(axis=0) >= 2)                  # points seen twice or more
    assert rebuilt.visible.shape == (n, len(seen)) == (n, len(tracks))
    # tracks are ordered by their first (image, line): map them back to the generating points
    point_of = [int(np.flatnonzero(line_of[t[0][0] - 1] == t[0][1])[0]) for t in tracks]
    assert sorted(point_of) == seen.tolist()
    assert np.array_equal(rebuilt.visible, scene.visible[:, point_of])
    assert np.array_equal(rebuilt.lines[rebuilt.visible], scene.lines[:, point_of][rebuilt.visible])
    assert np.array_equal(rebuilt.aligned, scene.aligned[point_of])


def _initial_sets_literal(g, image_aligned, check_image_ids, min_al=20, min_un=20):
    """incremental_mapper.cc:253-426 statement by statement (sets of tuples, itertools)."""
    from itertools import combinations
    all_tracks = {True: {}, False: {}}
    for image_id in check_image_ids:
        for line_idx, flag in enumerate(image_aligned[image_id - 1]):
            corrs = [c for c in g.FindCorrespondences(image_id, line_idx)
                     if bool(image_aligned[c[0] - 1][c[1]]) == bool(flag)]
            if len(corrs) < 3:
                continue
            for trio in combinations(corrs, 3):
                cand = {}
                for image, line in ((image_id, line_idx),) + trio:       # std::set keyed by image id
                    cand.setdefault(image, line)
                if len(cand) == 4:
                    key = tuple(sorted(cand))
                    all_tracks[bool(flag)].setdefault(key, set()).add(tuple(cand[i] for i in key))
    numbers = [(key, len(tr), len(all_tracks[False][key])) for key, tr in sorted(all_tracks[True].items())
               if len(tr) >= min_al and len(all_tracks[False].get(key, ())) >= min_un]
    numbers.sort(key=lambda t: -t[1])
    return numbers, all_tracks


def test_initial_image_sets_and_selection():
    """mapper.find_initial_image_sets against a literal restatement of the reference's candidate
    search; mapper.select_initial_images (host estimators, no GPU) picks the tried set with the
    best inlier ratio and its poses explain the tracks."""
    from privacy_preserving_sfm_b200 import mapper as M
    scene, gt = M.make_mapper_scene(num_images=7, num_points=260, seed=6, visibility=0.75, noise_px=0.2)
    rng = np.random.default_rng(2)
    n, p = scene.visible.shape
    line_of = -np.ones((n, p), np.int64)
    image_lines, image_aligned = [], []
    for i in range(n):
        vis = np.flatnonzero(scene.visible[i])
        perm = rng.permutation(len(vis))
        line_of[i, vis[perm]] = np.arange(len(vis))
        image_lines.append(scene.lines[i, vis[perm]])
        image_aligned.append(scene.aligned[vis[perm]])
    g = CorrespondenceGraph()
    for i in range(n):
        g.AddImage(i + 1, len(image_lines[i]))
    for i in range(n):
        for j in range(i + 1, n):
            both = np.flatnonzero(scene.visible[i] & scene.visible[j])
            both = both[rng.uniform(size=len(both)) < 0.9]
            g.AddCorrespondences(i + 1, j + 1, np.stack([line_of[i, both], line_of[j, both]], 1))
    g.Finalize()
    check = [2, 5, 6]
    got = M.find_initial_image_sets(g, image_aligned, check)
    want, all_tracks = _initial_sets_literal(g, image_aligned, check)
    assert len(got) == len(want) > 3
    assert [(d["image_set"], len(d["aligned_tracks"]), len(d["unaligned_tracks"])) for d in got] == want
    for d in got[:5]:
        assert [tuple(r) for r in d["aligned_tracks"].tolist()] == sorted(all_tracks[True][d["image_set"]])
        assert [tuple(r) for r in d["unaligned_tracks"].tolist()] == sorted(all_tracks[False][d["image_set"]])
        assert any(i in check for i in d["image_set"])                    # every track holds its reference line
    ok, image_set, poses, ratio, tried = M.select_initial_images(
        g, image_lines, image_aligned, scene.gravity, check, max_num_init_tries=3)
    assert ok and len(tried) == 3 and [t[0] for t in tried] == [d["image_set"] for d in got[:3]]
    assert ratio == max(t[2] for t in tried if t[1])
    assert image_set == next(t[0] for t in tried if t[1] and t[2] == ratio)   # first best: ">" at :505
    assert ratio > 0.9 and poses.shape == (4, 3, 4)
    # the poses explain the four-view tracks: relative rotations equal the generating ones
    ids = [i - 1 for i in image_set]
    for k in range(1, 4):
        rel = poses[k, :, :3] @ poses[0, :, :3].T
        rel_gt = gt["R"][ids[k]] @ gt["R"][ids[0]].T
        assert np.degrees(np.arccos(np.clip((np.trace(rel @ rel_gt.T) - 1) / 2, -1, 1))) < 0.5


def test_graph_from_visibility_is_the_pairwise_graph(ref):
    """correspondence_graph.graph_from_visibility (config 5: "synthetic CorrespondenceGraph from
    ground-truth visibility") == the reference's class fed pair by pair with the same matches."""
    from privacy_preserving_sfm_b200.correspondence_graph import graph_from_visibility
    visible, line_of, _ = _scene(9, 70, 0.4, seed=31, extra_lines=0)
    visible[8] = False                                                     # an image that sees nothing
    g, line_of2, num_lines = graph_from_visibility(visible, np.where(visible, line_of, -1))
    b = ref.CorrespondenceGraph()
    for i, n in enumerate(num_lines):
        b.AddImage(i + 1, n)
    for i in range(9):
        for j in range(i + 1, 9):
            both = np.flatnonzero(visible[i] & visible[j])
            if len(both):
                b.AddCorrespondences(i + 1, j + 1, np.stack([line_of[i, both], line_of[j, both]], 1))
    b.Finalize()
    _same_answers(g, b, num_lines, transitivities=(1, 3))
    assert not g.ExistsImage(9)
    g2, default_line_of, _ = graph_from_visibility(visible)                # default: rank among visible
    assert np.array_equal(default_line_of[0][visible[0]], np.arange(visible[0].sum()))
    assert len(g2.Tracks()) == int((visible.sum(axis=0) >= 2).sum())


@pytest.mark.parametrize("seed", range(12))
def test_random_dirty_match_sets_identical_to_the_reference(ref, capfd, seed):
    """Fuzz: few images, few lines, many random matches — so that duplicates, conflicting matches,
    repeated / swapped pairs, self matches and out-of-range indices all occur together."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(3, 6))
    num_lines = [int(rng.integers(1, 7)) for _ in range(n)]
    calls = []
    for _ in range(int(rng.integers(4, 14))):
        i, j = (int(v) for v in rng.integers(1, n + 1, size=2))
        k = int(rng.integers(0, 9))
        m = np.stack([rng.integers(0, num_lines[i - 1] + 2, size=k),
                      rng.integers(0, num_lines[j - 1] + 2, size=k)], 1)
        calls.append((i, j, m))
    a, b = CorrespondenceGraph(), ref.CorrespondenceGraph()
    _feed((a, b), num_lines, calls, finalize=False)
    for i in range(1, n + 1):
        assert a.NumCorrespondencesForImage(i) == b.NumCorrespondencesForImage(i)
        for line in range(num_lines[i - 1]):
            assert a.FindCorrespondences(i, line) == b.FindCorrespondences(i, line), (seed, i, line)
    a.Finalize()
    b.Finalize()
    _same_answers(a, b, num_lines, transitivities=(1, 2, 4))
    capfd.readouterr()                                                     # the reference's warnings

"""In-memory correspondence graph: host-side mirror of the reference's
``CorrespondenceGraph`` (src/base/correspondence_graph.{h,cc}) — the structure between the
matcher's pairwise line matches and the mapper: which line of which image corresponds to which,
how many images / pairs / observations there are, and the transitive closure the triangulator
walks (``IncrementalTriangulator::Find``, src/sfm/incremental_triangulator.cc:426-466).

Same member names, argument meaning and results as the reference, including
  * ``AddCorrespondences`` (:80-161): self-matches are ignored; a match with a line index out of
    range is dropped; a match whose line already has a correspondence into the other image (on
    either side, earlier calls and earlier matches of the same call included) is a duplicate and
    dropped; the per-image and per-pair counts hold the matches that were kept; the pair exists
    from the first call even if nothing was kept;
  * ``Finalize`` (:55-71): ``NumObservationsForImage`` = lines with at least one correspondence;
    images without any are erased (``ExistsImage`` turns false, queries on them raise);
  * ``FindTransitiveCorrespondences`` (:163-223): breadth first, level by level, in insertion
    order of every line's correspondences, and — as the reference does — the query line itself is
    removed by overwriting the first entry with the last one;
  * ``FindCorrespondencesBetweenImages`` (:225-248), ``IsTwoViewObservation`` (:250-262).
Storage is an edge list in insertion order with a CSR index per (image, line) built on demand
(the reference keeps a vector of vectors per image).

``Tracks`` is this package's addition: the connected components of the graph, i.e. the tracks
``mapper.Scene`` takes as given (for complete pairwise matching they are what the reference's
triangulator assembles through ``Find`` with transitivity 1).

Pinned against the reference's own class compiled from /root/reference
(oracle/_ref/libref_filter.so, oracle/ref/ref_corr_graph.cc): tests/test_ref_correspondence_graph.py.
"""
import numpy as np

K_MAX_NUM_IMAGES = 2 ** 31 - 1                       # Database::kMaxNumImages (base/database.cc:229-230)


def ImagePairToPairId(image_id1, image_id2):
    """Database::ImagePairToPairId (src/base/database.h:311-322)."""
    if image_id1 > image_id2:
        return K_MAX_NUM_IMAGES * image_id2 + image_id1
    return K_MAX_NUM_IMAGES * image_id1 + image_id2


def _all_distinct(v):
    if len(v) < 256:
        return len(set(v.tolist())) == len(v)
    return len(np.unique(v)) == len(v)


class CorrespondenceGraph:
    def __init__(self):
        self._num_lines = {}                         # image_id -> number of lines
        self._num_corrs = {}                         # image_id -> correspondences kept
        self._num_obs = {}                           # image_id -> lines with a correspondence (Finalize)
        self._pairs = {}                             # pair_id -> correspondences kept
        self._kept = {}                              # pair_id -> [(lines of the lower id, of the higher id)]
        self._edges = []                             # (image_id1, image_id2, kept matches [n, 2])
        self._index = None

    # ---- construction ---------------------------------------------------------------------------
    def AddImage(self, image_id, num_lines):
        if image_id in self._num_lines:
            raise ValueError("CHECK(!ExistsImage(image_id))")
        self._num_lines[int(image_id)] = int(num_lines)
        self._num_corrs[int(image_id)] = 0
        self._num_obs[int(image_id)] = 0
        self._index = None

    def AddCorrespondences(self, image_id1, image_id2, matches):
        """matches [n, 2]: (line_idx1, line_idx2).  Returns the number of matches kept."""
        image_id1, image_id2 = int(image_id1), int(image_id2)
        if image_id1 == image_id2:
            return 0
        n1, n2 = self._num_lines[image_id1], self._num_lines[image_id2]       # images_.at()
        m = np.asarray(matches, np.int64).reshape(-1, 2)
        pair_id = ImagePairToPairId(image_id1, image_id2)
        self._pairs.setdefault(pair_id, 0)
        if len(m) and (m.min() < 0 or m[:, 0].max() >= n1 or m[:, 1].max() >= n2):
            m = m[(m[:, 0] >= 0) & (m[:, 0] < n1) & (m[:, 1] >= 0) & (m[:, 1] < n2)]
        swap = image_id1 > image_id2
        earlier = self._kept.setdefault(pair_id, [])
        if not earlier and _all_distinct(m[:, 0]) and _all_distinct(m[:, 1]):
            keep = m                                                          # no duplicate possible
        else:
            used1, used2 = set(), set()
            for lo, hi in earlier:
                a, b = (hi, lo) if swap else (lo, hi)
                used1.update(a.tolist())
                used2.update(b.tolist())
            rows = []
            for k, (a, b) in enumerate(m.tolist()):
                if a in used1 or b in used2:
                    continue
                used1.add(a)
                used2.add(b)
                rows.append(k)
            keep = m[rows]
        if len(keep):
            earlier.append((keep[:, 1], keep[:, 0]) if swap else (keep[:, 0], keep[:, 1]))
            self._edges.append((image_id1, image_id2, keep))
            self._num_corrs[image_id1] += len(keep)
            self._num_corrs[image_id2] += len(keep)
            self._pairs[pair_id] += len(keep)
            self._index = None
        return len(keep)

    def Finalize(self):
        idx = self._build()
        erased = False
        for image_id in list(self._num_lines):
            start = idx["start"][image_id]
            obs = int((np.diff(start) > 0).sum())
            if obs == 0:
                for d in (self._num_lines, self._num_corrs, self._num_obs):
                    del d[image_id]
                erased = True
            else:
                self._num_obs[image_id] = obs
        if erased:
            self._index = None

    # ---- CSR index: per image, start[line] .. start[line + 1] into (dst_image, dst_line) ------------
    def _build(self):
        if self._index is not None:
            return self._index
        ids = sorted(self._num_lines)
        base = {}
        total = 0
        for i in ids:
            base[i] = total
            total += self._num_lines[i]
        if self._edges and ids:
            ids_arr = np.array(ids, np.int64)
            base_arr = np.array([base[i] for i in ids], np.int64)
            lens = np.array([len(m) for _, _, m in self._edges], np.int64)
            id1 = np.repeat(np.array([e[0] for e in self._edges], np.int64), lens)
            id2 = np.repeat(np.array([e[1] for e in self._edges], np.int64), lens)
            m = np.concatenate([e[2] for e in self._edges])
            k1 = np.minimum(np.searchsorted(ids_arr, id1), len(ids) - 1)
            k2 = np.minimum(np.searchsorted(ids_arr, id2), len(ids) - 1)
            alive = (ids_arr[k1] == id1) & (ids_arr[k2] == id2)   # erased by Finalize: nothing points to them
            id1, id2, m, k1, k2 = id1[alive], id2[alive], m[alive], k1[alive], k2[alive]
            # corrs1.emplace_back(image_id2, idx2); corrs2.emplace_back(image_id1, idx1) per match
            src = np.empty(2 * len(m), np.int64)
            src[0::2], src[1::2] = base_arr[k1] + m[:, 0], base_arr[k2] + m[:, 1]
            dst_img = np.empty(2 * len(m), np.int64)
            dst_img[0::2], dst_img[1::2] = id2, id1
            dst_line = np.empty(2 * len(m), np.int64)
            dst_line[0::2], dst_line[1::2] = m[:, 1], m[:, 0]
            order = np.argsort(src, kind="stable")    # insertion order within every line
            src, dst_img, dst_line = src[order], dst_img[order], dst_line[order]
        else:
            src = dst_img = dst_line = np.zeros(0, np.int64)
        offsets = np.searchsorted(src, np.arange(total + 1))
        start = {i: offsets[base[i]:base[i] + self._num_lines[i] + 1] for i in ids}
        self._index = dict(start=start, dst_img=dst_img, dst_line=dst_line)
        return self._index

    def CSR(self):
        """The index itself, for vectorised walks over many lines: dict(start {image_id: offsets
        [num_lines + 1]}, dst_img, dst_line) — line l of image i corresponds to
        (dst_img[k], dst_line[k]) for k in start[i][l] .. start[i][l + 1], in insertion order."""
        return self._build()

    def _corrs(self, image_id, line_idx):
        idx = self._build()
        start = idx["start"][image_id]                                        # images_.at(image_id)
        if not 0 <= line_idx < len(start) - 1:
            raise IndexError("corrs.at(line_idx)")
        return idx["dst_img"][start[line_idx]:start[line_idx + 1]], \
            idx["dst_line"][start[line_idx]:start[line_idx + 1]]

    # ---- queries ----------------------------------------------------------------------------------
    def NumImages(self):
        return len(self._num_lines)

    def NumImagePairs(self):
        return len(self._pairs)

    def ExistsImage(self, image_id):
        return image_id in self._num_lines

    def NumObservationsForImage(self, image_id):
        return self._num_obs[image_id]

    def NumCorrespondencesForImage(self, image_id):
        return self._num_corrs[image_id]

    def NumCorrespondencesBetweenImages(self, image_id1=None, image_id2=None):
        """With two ids: the count of that pair (0 if the pair does not exist); without: {pair_id: count}."""
        if image_id1 is None:
            return dict(self._pairs)
        return self._pairs.get(ImagePairToPairId(image_id1, image_id2), 0)

    def FindCorrespondences(self, image_id, line_idx):
        """[(image_id, line_idx)] in insertion order."""
        di, dl = self._corrs(image_id, line_idx)
        return list(zip(di.tolist(), dl.tolist()))

    def HasCorrespondences(self, image_id, line_idx):
        return len(self._corrs(image_id, line_idx)[0]) > 0

    def FindTransitiveCorrespondences(self, image_id, line_idx, transitivity):
        if transitivity == 1:
            return self.FindCorrespondences(image_id, line_idx)
        if not self.HasCorrespondences(image_id, line_idx):
            return []
        found = [(image_id, line_idx)]
        seen = {(image_id, line_idx)}
        begin, end = 0, 1
        for _ in range(int(transitivity)):
            for k in range(begin, end):
                for corr in self.FindCorrespondences(*found[k]):
                    if corr not in seen:
                        seen.add(corr)
                        found.append(corr)
            begin, end = end, len(found)
            if begin == end:
                break
        if len(found) > 1:
            found[0] = found[-1]
        found.pop()
        return found

    def FindCorrespondencesBetweenImages(self, image_id1, image_id2):
        """[(line_idx1, line_idx2)] by ascending line_idx1."""
        if self.NumCorrespondencesBetweenImages(image_id1, image_id2) == 0:
            return []
        idx = self._build()
        start = idx["start"][image_id1]
        sel = np.flatnonzero(idx["dst_img"][start[0]:start[-1]] == image_id2) + start[0]
        line1 = np.searchsorted(start, sel, side="right") - 1
        return list(zip(line1.tolist(), idx["dst_line"][sel].tolist()))

    def IsTwoViewObservation(self, image_id, line_idx):
        di, dl = self._corrs(image_id, line_idx)
        if len(di) != 1:
            return False
        return len(self._corrs(int(di[0]), int(dl[0]))[0]) == 1

    # ---- tracks: connected components (this package's addition, see the module docstring) ----------
    def Tracks(self, min_length=2):
        """[[(image_id, line_idx), ...], ...]: the connected components with at least ``min_length``
        lines, each sorted by (image_id, line_idx), ordered by their first element."""
        idx = self._build()
        ids = sorted(self._num_lines)
        base, total = {}, 0
        for i in ids:
            base[i] = total
            total += self._num_lines[i]
        img_of = np.repeat(np.array(ids, np.int64), [self._num_lines[i] for i in ids]) if ids else np.zeros(0, np.int64)
        base_arr = np.array([base[i] for i in ids], np.int64)
        ids_arr = np.array(ids, np.int64)
        # edges (node -> node): the CSR arrays are ordered by source node
        src = np.concatenate([np.repeat(np.arange(base[i], base[i] + self._num_lines[i]),
                                        np.diff(idx["start"][i])) for i in ids]) if ids else np.zeros(0, np.int64)
        parent = np.arange(total)
        if len(src):
            dst = base_arr[np.searchsorted(ids_arr, idx["dst_img"])] + idx["dst_line"]
            try:
                from scipy.sparse import coo_matrix
                from scipy.sparse.csgraph import connected_components
                adj = coo_matrix((np.ones(len(src), np.int8), (src, dst)), shape=(total, total))
                _, label = connected_components(adj, directed=False)
                # label every node by the smallest node of its component (order of first element)
                first = np.full(label.max() + 1, total, np.int64)
                np.minimum.at(first, label, np.arange(total))
                parent = first[label]
            except ImportError:                       # label propagation by pointer jumping
                while True:
                    m = np.minimum(parent[src], parent[dst])
                    new = parent.copy()
                    np.minimum.at(new, src, m)
                    np.minimum.at(new, dst, m)
                    new = new[new]
                    if np.array_equal(new, parent):
                        break
                    parent = new
        order = np.argsort(parent, kind="stable")
        labels = parent[order]
        cuts = np.flatnonzero(np.diff(labels)) + 1
        tracks = []
        for comp in np.split(order, cuts):
            if len(comp) >= min_length:
                tracks.append([(int(img_of[n]), int(n - base[int(img_of[n])])) for n in comp])
        return tracks


def graph_from_visibility(visible, line_of=None, finalize=True):
    """The synthetic correspondence graph of a generated scene ("synthetic
    CorrespondenceGraph from ground-truth visibility"): image ids index + 1, every pair of images
    matched on the points both see.  visible [n_images, n_points] bool; line_of [n_images,
    n_points]: the line index of a point in an image (default: its rank among the image's visible
    points).  Returns (graph, line_of, num_lines)."""
    visible = np.asarray(visible, bool)
    n, _ = visible.shape
    if line_of is None:
        line_of = np.where(visible, np.cumsum(visible, axis=1) - 1, -1)
    num_lines = [int(line_of[i][visible[i]].max()) + 1 if visible[i].any() else 0 for i in range(n)]
    g = CorrespondenceGraph()
    for i in range(n):
        g.AddImage(i + 1, num_lines[i])
    # all (point, view a < view b) triples, grouped by image pair
    pt, img = np.nonzero(visible.T)                      # sorted by point, then image
    start = np.searchsorted(pt, np.arange(visible.shape[1] + 1))
    length = np.diff(start)
    a_list, b_list = [], []
    for L in np.unique(length[length >= 2]):
        pts = np.flatnonzero(length == L)
        views = img[start[pts][:, None] + np.arange(L)[None, :]]        # [points of this length, L]
        ia, ib = np.triu_indices(L, 1)
        a_list.append(np.stack([views[:, ia].ravel(), views[:, ib].ravel(),
                                np.repeat(pts, len(ia))], 1))
    if a_list:
        tri = np.concatenate(a_list)
        order = np.argsort((tri[:, 0] * n + tri[:, 1]) * visible.shape[1] + tri[:, 2], kind="stable")
        tri = tri[order]                              # by image pair, then point
        matches = np.stack([line_of[tri[:, 0], tri[:, 2]], line_of[tri[:, 1], tri[:, 2]]], 1)
        key = tri[:, 0] * n + tri[:, 1]
        cuts = np.concatenate([[0], np.flatnonzero(np.diff(key)) + 1, [len(tri)]])
        first = tri[cuts[:-1], :2] + 1
        for (i, j), lo, hi in zip(first.tolist(), cuts[:-1].tolist(), cuts[1:].tolist()):
            g.AddCorrespondences(i, j, matches[lo:hi])
    if finalize:
        g.Finalize()
    return g, line_of, num_lines

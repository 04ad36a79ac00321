"""ORACLE (test infrastructure): CPU restatement of the reference's merge / clustering path.

Follows /root/reference/tree_learn/util/pipeline.py:
  ensemble_ref            :113-141  (round(2) group-by-mean, output sorted by (x,y,z))
  get_instances_ref       :145-169
  group_dbscan_ref        :173-180  (DBSCAN(eps, min_samples=2) == connected components of the
                                     <=eps graph, singletons noise, labels by lowest member index)
  make_labels_consecutive :195-206
  assign_remaining_ref    :287-296  (kNN(5) uniform vote, ties -> smallest label)
Pinned against the reference's OWN functions (sklearn/pandas, imported from /root/reference in
this container) by tests/golden/make_golden.py -> tests/golden/cluster_*.npz.
numpy + scipy only.  Only tests/, smoke() and bench.py's cpu_baseline legs may import it.
"""
import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components
from scipy.spatial import cKDTree


def round2_key(coords):
    """Integer key of pandas `round(2)` on float32 columns: rint(x*100) evaluated in float32."""
    c = np.asarray(coords, dtype=np.float32)
    return np.rint(c * np.float32(100.0)).astype(np.int64)


def ensemble_ref(coords, semantic_scores, semantic_labels, offset_predictions, offset_labels,
                 instance_labels, feats, input_feats):
    key = round2_key(coords)
    order = np.lexsort((key[:, 2], key[:, 1], key[:, 0]))
    ks = key[order]
    head = np.ones(len(ks), dtype=bool)
    head[1:] = np.any(ks[1:] != ks[:-1], axis=1)
    gid = np.cumsum(head) - 1
    ng = int(gid[-1]) + 1 if len(gid) else 0
    cnt = np.bincount(gid, minlength=ng).astype(np.float64)

    def mean(a):
        a = np.asarray(a)
        a2 = a.reshape(len(a), -1)[order].astype(np.float64)
        out = np.zeros((ng, a2.shape[1]))
        np.add.at(out, gid, a2)
        return out / cnt[:, None]

    out_coords = (ks[head].astype(np.float32) / np.float32(100.0)).astype(np.float32)
    return (out_coords, mean(semantic_scores).astype(np.float32),
            mean(semantic_labels).astype(np.int64).flatten(), mean(offset_predictions).astype(np.float32),
            mean(offset_labels).astype(np.float32), mean(instance_labels).astype(np.int64).flatten(),
            mean(feats).astype(np.float32), mean(input_feats).astype(np.float32))


def make_labels_consecutive_ref(labels, start_num):
    palette = np.unique(labels)
    return np.searchsorted(palette, labels) + start_num


def radius_components(points, radius):
    """sklearn DBSCAN(eps=radius, min_samples=2).labels_ : components of the d<=radius graph
    (distances in float64 on the given values), singletons -1, ids by lowest member index."""
    x = np.asarray(points, dtype=np.float64)
    n = len(x)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    pairs = cKDTree(x).query_pairs(radius, output_type='ndarray')
    # cKDTree prunes with squared distances like sklearn's KDTree; recheck exactly as rdist <= r*r
    d = x[pairs[:, 0]] - x[pairs[:, 1]]
    rd = np.zeros(len(pairs))
    for j in range(x.shape[1]):
        rd = rd + d[:, j] * d[:, j]
    pairs = pairs[rd <= radius * radius]
    g = coo_matrix((np.ones(len(pairs)), (pairs[:, 0], pairs[:, 1])), shape=(n, n))
    _, comp = connected_components(g, directed=False)
    size = np.bincount(comp)
    first = np.full(comp.max() + 1, n, dtype=np.int64)
    np.minimum.at(first, comp, np.arange(n))
    valid = size >= 2
    rank = -np.ones(len(size), dtype=np.int64)
    order = np.argsort(first[valid], kind='stable')
    ids = np.nonzero(valid)[0][order]
    rank[ids] = np.arange(len(ids))
    return rank[comp]


def group_dbscan_ref(cluster_coords, radius, npoint_thr, not_assigned_label, start_num_preds):
    labels = radius_components(cluster_coords, radius)
    nums, counts = np.unique(labels, return_counts=True)
    valid = nums[(counts >= npoint_thr) & (nums != -1)]
    ind_valid = np.isin(labels, valid)
    out = np.full(len(labels), not_assigned_label, dtype=np.int64)
    if ind_valid.any():
        out[ind_valid] = make_labels_consecutive_ref(labels[ind_valid], start_num_preds)
    return out


def softmax_tree_mask(logits, tree_class, thresh):
    l = np.asarray(logits, dtype=np.float32)
    m = l.max(axis=1, keepdims=True)
    e = np.exp(l - m)
    p = e / e.sum(axis=1, keepdims=True)
    return p[:, tree_class] >= thresh


def get_instances_ref(coords, offset, logits, tree_conf_thresh, tau_vert, tau_off, tau_group, tau_min,
                      verticality_feat, tree_class=0, non_trees_label=0, not_assigned_label=-1,
                      start_num_preds=1, tree_mask=None):
    cluster_coords = (np.asarray(coords) + np.asarray(offset))[:, :3]
    if tree_mask is None:
        tree_mask = softmax_tree_mask(logits, tree_class, tree_conf_thresh)
    mask = tree_mask & (np.asarray(verticality_feat) > tau_vert) & (np.abs(offset[:, 2]) < tau_off)
    ind = np.where(mask)[0]
    pred = non_trees_label * np.ones(len(cluster_coords))
    pred[tree_mask] = not_assigned_label
    pred[ind] = group_dbscan_ref(cluster_coords[ind][:, :2], tau_group, tau_min, not_assigned_label, start_num_preds)
    return pred.astype(np.int64)


def assign_remaining_ref(coords, predictions, remaining_points_idx=-1, n_neighbors=5):
    pred = np.copy(predictions)
    q = np.argwhere(pred == remaining_points_idx).reshape(-1)
    r = np.argwhere(pred != remaining_points_idx).reshape(-1)
    if len(q) == 0:
        return pred.astype(np.int64)
    x = np.asarray(coords, dtype=np.float64)
    _, nn = cKDTree(x[r]).query(x[q], k=n_neighbors)
    lab = pred[r][nn.reshape(len(q), -1)]
    out = np.empty(len(q), dtype=pred.dtype)
    for i in range(len(q)):
        vals, counts = np.unique(lab[i], return_counts=True)
        out[i] = vals[np.argmax(counts)]      # first max over sorted classes = smallest label on ties
    pred[q] = out
    return pred.astype(np.int64)


# ---- HDBSCAN (reference util/pipeline.py:184-191 -> sklearn.cluster.HDBSCAN(min_cluster_size=m)) -----------------------
# sklearn (>= 1.3; 1.9.0 in this image) is a third-party dependency of the reference; its algorithm for dense Euclidean
# input is restated here (sklearn/cluster/_hdbscan/{hdbscan.py::_hdbscan_prims, _linkage.pyx, _tree.pyx}):
#   core distance = distance to the min_samples-th neighbour incl. the point itself (min_samples = min_cluster_size),
#   Prim MST of max(core_a, core_b, d_ab) from node 0, strict '<' relaxations, first minimal index wins,
#   np.argsort of the edge weights, union-find dendrogram, condensed tree, stability, excess of mass, labelling.
# Pinned against sklearn.cluster.HDBSCAN / the reference's group_hdbscan by tests/golden/make_golden_hdbscan.py.
def core_distances_ref(points, k):
    p = np.asarray(points, dtype=np.float64)
    out = np.empty(len(p))
    for i in range(len(p)):                      # d2 = dx*dx + dy*dy in fp64, like KDTree's reduced distance
        d2 = (p[i, 0] - p[:, 0]) ** 2 + (p[i, 1] - p[:, 1]) ** 2
        out[i] = np.sqrt(np.partition(d2, k - 1)[k - 1])
    return out


def prim_mst_ref(points, core):
    p = np.asarray(points, dtype=np.float64)
    n = len(p)
    in_tree = np.zeros(n, dtype=bool)
    min_reach = np.full(n, np.inf)
    source = np.ones(n, dtype=np.int64)
    src, dst, w = np.empty(n - 1, np.int64), np.empty(n - 1, np.int64), np.empty(n - 1)
    cur = 0
    for i in range(n - 1):
        in_tree[cur] = True
        d = np.sqrt((p[cur, 0] - p[:, 0]) ** 2 + (p[cur, 1] - p[:, 1]) ** 2)
        mrd = np.maximum(np.maximum(core[cur], core), d)
        upd = (mrd < min_reach) & ~in_tree
        min_reach[upd] = mrd[upd]
        source[upd] = cur
        cand = np.where(in_tree, np.inf, min_reach)
        new = int(np.argmin(cand))               # first minimal index, like the ascending strict-'<' scan
        src[i], dst[i], w[i] = source[new], new, cand[new]
        cur = new
    return src, dst, w


def hdbscan_tree_labels_ref(src, dst, w, n, min_cluster_size):
    """Edges sorted by weight -> labels (-1 noise).  Pure-Python port for small n."""
    m = n - 1
    parent = np.full(2 * n - 1, -1, dtype=np.int64)
    size = np.ones(2 * n - 1, dtype=np.int64)
    left, right, csize = np.empty(m, np.int64), np.empty(m, np.int64), np.empty(m, np.int64)

    def find(x):
        r = x
        while parent[r] != -1:
            r = parent[r]
        while parent[x] != -1 and parent[x] != r:
            parent[x], x = r, parent[x]
        return r
    for i in range(m):
        a, b = find(int(src[i])), find(int(dst[i]))
        left[i], right[i], csize[i] = a, b, size[a] + size[b]
        parent[a] = parent[b] = n + i
        size[n + i] = csize[i]

    def bfs(start):
        out = [start]
        h = 0
        while h < len(out):
            x = out[h]
            h += 1
            if x >= n:
                out += [int(left[x - n]), int(right[x - n])]
        return out
    root = 2 * m
    relabel = {root: n}
    ignore = set()
    rows, nxt = [], n + 1
    for node in bfs(root):
        if node in ignore or node < n:
            continue
        l, r, dist = int(left[node - n]), int(right[node - n]), w[node - n]
        lam = 1.0 / dist if dist > 0 else np.inf
        lc = csize[l - n] if l >= n else 1
        rc = csize[r - n] if r >= n else 1

        def spill(frm):
            for x in bfs(frm):
                if x < n:
                    rows.append((relabel[node], x, lam, 1))
                ignore.add(x)
        if lc >= min_cluster_size and rc >= min_cluster_size:
            relabel[l] = nxt
            rows.append((relabel[node], nxt, lam, lc))
            relabel[r] = nxt + 1
            rows.append((relabel[node], nxt + 1, lam, rc))
            nxt += 2
        elif lc < min_cluster_size and rc < min_cluster_size:
            spill(l)
            spill(r)
        elif lc < min_cluster_size:
            relabel[r] = relabel[node]
            spill(l)
        else:
            relabel[l] = relabel[node]
            spill(r)
    nc = nxt - n
    birth, stab = np.zeros(nc), np.zeros(nc)
    for p_, c, lam, s in rows:
        if c >= n:
            birth[c - n] = lam
    for p_, c, lam, s in rows:
        stab[p_ - n] += (lam - birth[p_ - n]) * s
    kids = [[] for _ in range(nc)]
    for p_, c, lam, s in rows:
        if s > 1:
            kids[p_ - n].append(c - n)
    is_cluster = np.ones(nc, dtype=bool)
    is_cluster[0] = False
    for c in range(nc - 1, 0, -1):
        sub = sum(stab[k] for k in kids[c]) if kids[c] else 0.0
        if sub > stab[c]:
            is_cluster[c] = False
            stab[c] = sub
        else:
            q = list(kids[c])
            while q:
                x = q.pop()
                is_cluster[x] = False
                q += kids[x]
    label_of = np.cumsum(is_cluster) - 1
    up = {}

    def top(x):
        while x in up:
            x = up[x]
        return x
    for p_, c, lam, s in rows:      # child rows hang under their parent unless the child is a selected cluster
        if not (c >= n and is_cluster[c - n]):
            up[c] = p_
    labels = np.empty(n, dtype=np.int64)
    for i in range(n):
        c = top(i)
        labels[i] = label_of[c - n] if c != n else -1
    return labels


def hdbscan_ref(points, min_cluster_size):
    pts = np.asarray(points)
    core = core_distances_ref(pts, min_cluster_size)
    src, dst, w = prim_mst_ref(pts, core)
    order = np.argsort(w)
    return hdbscan_tree_labels_ref(src[order], dst[order], w[order], len(pts), min_cluster_size)


def group_hdbscan_ref(cluster_coords, npoint_thr, not_assigned_label, start_num_preds):
    labels = hdbscan_ref(cluster_coords, npoint_thr)
    nums, cnt = np.unique(labels, return_counts=True)
    valid = nums[(cnt >= npoint_thr) & (nums != -1)]
    ind = np.isin(labels, valid)
    out = np.full(len(labels), not_assigned_label, dtype=np.int64)
    if ind.any():
        out[ind] = make_labels_consecutive_ref(labels[ind], start_num_preds)
    return out

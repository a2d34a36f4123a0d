"""Oracle-independent validity check of a succinct de Bruijn graph (SURVEY.md B.6a invariants 1-5).

The only consumer of the graph is `megahit_core assemble -s <prefix>` (reference call site
/root/reference/assemble/assemble_wrapper.py:264-295), which never sees node labels: it navigates with rank/select over
W and last (LF mapping) and reads explicit labels only for tips.  `decode()` does the same -- it recovers every node
label from W / last / tip / tip labels ALONE -- and `expected_edges()` builds the canonical (k+1)-mer multiset straight
from the reads with numpy.  Neither touches oracle/ or tests/pymodel.py, so agreement is evidence that the emitted arrays
are a well-formed BOSS graph of exactly the solid edge set, whatever produced them (oracle or GPU).

Orientation: megahit stores sequences reversed; a stored item (b, X = x0..x(k-1)) is the edge b.X of the stored strings,
items are sorted by X, W = b + 1 (first occurrence of b in the (k-1)-prefix group of X) or b + 5 (repeat), 0 for b = '$'.
Following W from item i reaches the node b.x0..x(k-2), i.e. the r-th node whose label starts with b, where r is the rank
of i among the non-minus items with the same W.  Tips (x(k-1) = '$') are not nodes for that count (last = 0).
"""
import numpy as np

COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


def unpack_labels(words, nchars):
    """uint32 [n, nw] left-aligned 2-bit strings -> uint8 [n, nchars]."""
    words = np.asarray(words, dtype=np.uint32)
    n = words.shape[0]
    out = np.zeros((n, nchars), np.uint8)
    for j in range(nchars):
        out[:, j] = (words[:, j >> 4] >> np.uint32(30 - 2 * (j & 15))) & 3
    return out


def decode(w, last, tip, tip_labels, k):
    """Node label of every item from W / last / tip / tip labels only.

    Returns (labels uint8 [n, k] with chars 1..4 = ACGT and 0 = '$' (only the last char of a tip), node_of_item).
    Raises AssertionError when the arrays are not a consistent BOSS graph (invariants 1, 2)."""
    w = np.asarray(w).astype(np.int64)
    last = np.asarray(last).astype(np.int64)
    tip = np.asarray(tip).astype(bool)
    n = len(w)
    assert not (tip & (last == 1)).any(), "a tip item carries last = 1"
    # nodes = runs of non-tip items closed by last = 1
    node_of = np.cumsum(last) - last
    V = int(last.sum())
    nontip = ~tip
    if nontip.any():
        assert last[np.nonzero(nontip)[0][-1]] == 1, "the final node is not closed by last = 1"
    # invariant 1: every node has exactly one non-minus in-edge
    nonminus = (w >= 1) & (w <= 4)
    assert int(nonminus.sum()) == V, f"invariant 1: {int(nonminus.sum())} non-minus edges for {V} nodes"
    # first char of node v: nodes are sorted by label, cnt[c] of them start with c (invariant 2 is what makes this a bijection)
    cnt = np.array([(w == c + 1).sum() for c in range(4)], dtype=np.int64)
    F = np.concatenate([[0], np.cumsum(cnt)])
    first = np.repeat(np.arange(4, dtype=np.uint8), cnt)            # [V]
    # predecessor edge of node v: the (v - F[c])-th non-minus item with W = c + 1 (item order)
    pred_item = np.empty(V, np.int64)
    for c in range(4):
        pred_item[F[c]:F[c + 1]] = np.nonzero(w == c + 1)[0]
    tips_idx = np.nonzero(tip)[0]
    tip_rank = np.cumsum(tip) - 1                                   # item -> row of tip_labels
    tl = unpack_labels(tip_labels, k - 1) if len(tips_idx) else np.zeros((0, k - 1), np.uint8)
    assert tl.shape[0] == len(tips_idx), "tip label count disagrees with the tip flags"
    pred_is_tip = tip[pred_item]
    pred_node = np.where(pred_is_tip, -1, node_of[pred_item])
    labels = np.zeros((V, k), np.uint8)
    done = np.zeros(V, bool)
    cur = np.arange(V, dtype=np.int64)
    for j in range(k):
        act = np.nonzero(~done)[0]
        if len(act) == 0:
            break
        u = cur[act]
        labels[act, j] = first[u] + 1
        hit = pred_is_tip[u]
        if hit.any() and j + 1 < k:
            # the in-edge comes out of a tip with explicit label T = t0..t(k-2)$: the rest of the label is T[0 .. k-2-j]
            rows = act[hit]
            t = tl[tip_rank[pred_item[u[hit]]]]
            labels[rows, j + 1:] = t[:, :k - 1 - j] + 1
        done[act[hit]] = True
        cur[act[~hit]] = pred_node[u[~hit]]
    out = np.zeros((n, k), np.uint8)
    out[nontip] = labels[node_of[nontip]]
    if len(tips_idx):
        out[tips_idx, :k - 1] = tl + 1
    return out, node_of


def check_graph(w, last, tip, mul, tip_labels, k):
    """Invariants 1-3 and 5 on the arrays alone; returns the decoded edge table (uint8 [E, k+1] stored orientation, mul)."""
    w = np.asarray(w)
    last = np.asarray(last)
    tip = np.asarray(tip)
    mul = np.asarray(mul)
    labels, node_of = decode(w, last, tip, tip_labels, k)
    # invariant 3
    assert (mul[tip == 1] == 0).all() and (mul[w == 0] == 0).all(), "invariant 3: a dummy carries a multiplicity"
    # invariant 2: nodes whose label starts with c == items with W = c + 1 (checked on the DECODED labels: closes the loop)
    nontip_last = (tip == 0) & (last == 1)
    for c in range(4):
        assert int((labels[nontip_last, 0] == c + 1).sum()) == int((w == c + 1).sum()), f"invariant 2 fails for base {c}"
    # invariant 5: items non-decreasing in label order ('$' lowest), equal labels contiguous = one node
    if len(labels) > 1:
        a, b = labels[:-1], labels[1:]
        diff = a != b
        anyd = diff.any(axis=1)
        fd = np.argmax(diff, axis=1)
        rows = np.nonzero(anyd)[0]
        assert (a[rows, fd[rows]] < b[rows, fd[rows]]).all(), "invariant 5: items are not sorted by node label"
        # a node's items are contiguous and a new label starts a new node
        same = ~anyd
        nt = (tip[:-1] == 0) & (tip[1:] == 0)
        assert (node_of[1:][same & nt] == node_of[:-1][same & nt]).all(), "equal labels split over two nodes"
        assert (node_of[1:][anyd & nt] != node_of[:-1][anyd & nt]).all(), "two labels inside one node"
    # minus flags: W = b + 5 iff b already occurred in the same (k-1)-prefix group (tips take part)
    b = ((w.astype(np.int64) - 1) % 4).astype(np.uint8)
    if len(labels):
        grp = np.concatenate([[0], np.cumsum((labels[1:, :k - 1] != labels[:-1, :k - 1]).any(axis=1))])
        has_b = np.nonzero(w > 0)[0]
        key = grp[has_b] * 4 + b[has_b]
        _, first_idx = np.unique(key, return_index=True)
        is_first = np.zeros(len(has_b), bool)
        is_first[first_idx] = True
        assert ((w[has_b] <= 4) == is_first).all(), "minus flags disagree with the (k-1)-prefix groups"
        # last = 1 exactly on the final item of each distinct k-th char of a group ... i.e. of each node (checked above)
    real = (w > 0) & (tip == 0)
    edges = np.concatenate([b[real, None] + 1, labels[real]], axis=1)
    return edges, mul[real]


def _kmers(bases, starts, k1):
    """all k1-mers of the reads, stored orientation (reads reversed), uint8 [n, k1] chars 1..4."""
    bases = np.asarray(bases, dtype=np.uint8)
    out = []
    for r in range(len(starts) - 1):
        s = bases[starts[r]:starts[r + 1]][::-1]
        if len(s) >= k1:
            out.append(np.lib.stride_tricks.sliding_window_view(s, k1))
    if not out:
        return np.zeros((0, k1), np.uint8)
    return np.concatenate(out) + 1


def _rc(x):
    return (5 - x)[:, ::-1]


def _rows_view(a):
    a = np.ascontiguousarray(a)
    return a.view(np.dtype((np.void, a.shape[1]))).ravel()


def expected_edges(bases, starts, k, m):
    """(k+1)-mers of the reads with canonical count >= m, closed under reverse complement: (uint8 [E, k+1], mult)."""
    km = _kmers(bases, starts, k + 1)
    if len(km) == 0:
        return km, np.zeros(0, np.int64)
    rc = _rc(km)
    # canonical = lexicographic min of the two
    diff = km != rc
    fd = np.argmax(diff, axis=1)
    idx = np.arange(len(km))
    use_rc = diff.any(axis=1) & (rc[idx, fd] < km[idx, fd])
    canon = np.where(use_rc[:, None], rc, km)
    u, cnts = np.unique(_rows_view(canon), return_counts=True)
    u = u.view(np.uint8).reshape(-1, k + 1)
    keep = cnts >= m
    u, cnts = u[keep], np.minimum(cnts[keep], 65535)
    r = _rc(u)
    pal = (u == r).all(axis=1)
    allk = np.concatenate([u, r[~pal]])
    allc = np.concatenate([cnts, cnts[~pal]])
    return allk, allc


def assert_round_trip(g, bases, starts, k, m):
    """g: dict with w, last, tip, mul, tip_labels.  Invariant 4: decoded edges == solid (k+1)-mers closed under revcomp."""
    edges, mul = check_graph(g["w"], g["last"], g["tip"], g["mul"], g["tip_labels"], k)
    exp_e, exp_c = expected_edges(bases, starts, k, m)
    assert len(edges) == len(exp_e), f"invariant 4: graph holds {len(edges)} edges, reads give {len(exp_e)}"
    if len(edges) == 0:
        return 0
    o1 = np.argsort(_rows_view(edges), kind="stable")
    o2 = np.argsort(_rows_view(exp_e), kind="stable")
    assert np.array_equal(edges[o1], exp_e[o2]), "invariant 4: edge sets differ"
    assert np.array_equal(np.asarray(mul)[o1].astype(np.int64), exp_c[o2].astype(np.int64)), "invariant 4: multiplicities differ"
    return len(edges)

"""Oracle (test infrastructure): rectangular linear sum assignment.

``match_by_tag`` calls ``scipy.optimize.linear_sum_assignment``
(mindpose/utils/match.py:8-11, :100).  scipy is a third-party dependency that is
not under /root/reference (requirements.txt: ``scipy>=1.5.4``; 1.18.1 in this
image, compiled ``_lsap`` module only), so its published algorithm is restated
here: the shortest-augmenting-path method of D. F. Crouse, "On implementing 2D
rectangular assignment algorithms", IEEE T-AES 52(4), 2016, as implemented in
scipy's ``rectangular_lsap.cpp``.  With ``use_rounded_norm`` the cost matrix is
integer valued and full of ties, so which optimum is returned depends on the
scan order; the details that fix it are kept:

* rows are augmented in index order; the matrix is transposed first when it has
  more rows than columns;
* the list of unscanned columns starts in REVERSE index order
  (``remaining[it] = nc - it - 1``) and a scanned column is replaced by the last
  entry of the list;
* among equal shortest path costs the first one in list order wins, except that
  an equal-cost column that is still unassigned (a sink) replaces it;
* dual variables are updated after every augmentation.

PINNED by tests/golden/lsap_ref.npz (scipy outputs on tie-heavy matrices) and,
when scipy is importable, by a live differential test.
"""
import numpy as np


def linear_sum_assignment(cost):
    """Returns (row_ind, col_ind) like scipy.optimize.linear_sum_assignment (minimise)."""
    cost = np.asarray(cost, dtype=np.float64)
    if cost.ndim != 2:
        raise ValueError("expected a matrix (2-D array), got a %r array" % (cost.shape,))
    if np.any(np.isnan(cost)) or np.any(np.isneginf(cost)):
        raise ValueError("matrix contains invalid numeric entries")
    nr, nc = cost.shape
    if nr == 0 or nc == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    transpose = nc < nr
    if transpose:
        cost = np.ascontiguousarray(cost.T)
        nr, nc = nc, nr

    u = np.zeros(nr)
    v = np.zeros(nc)
    shortest = np.zeros(nc)
    path = np.full(nc, -1, dtype=np.int64)
    col4row = np.full(nr, -1, dtype=np.int64)
    row4col = np.full(nc, -1, dtype=np.int64)

    for cur_row in range(nr):
        # ---- shortest augmenting path from cur_row
        min_val = 0.0
        remaining = [nc - it - 1 for it in range(nc)]
        num_remaining = nc
        sr = np.zeros(nr, dtype=bool)
        sc = np.zeros(nc, dtype=bool)
        shortest[:] = np.inf
        sink = -1
        i = cur_row
        while sink == -1:
            index = -1
            lowest = np.inf
            sr[i] = True
            for it in range(num_remaining):
                j = remaining[it]
                r = min_val + cost[i, j] - u[i] - v[j]
                if r < shortest[j]:
                    path[j] = i
                    shortest[j] = r
                if shortest[j] < lowest or (shortest[j] == lowest and row4col[j] == -1):
                    lowest = shortest[j]
                    index = it
            min_val = lowest
            if min_val == np.inf:
                raise ValueError("cost matrix is infeasible")
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            sc[j] = True
            num_remaining -= 1
            remaining[index] = remaining[num_remaining]
        # ---- dual update
        u[cur_row] += min_val
        for r_ in range(nr):
            if sr[r_] and r_ != cur_row:
                u[r_] += min_val - shortest[col4row[r_]]
        for c_ in range(nc):
            if sc[c_]:
                v[c_] -= min_val - shortest[c_]
        # ---- augment
        j = sink
        while True:
            i = path[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur_row:
                break

    if transpose:
        order = np.argsort(col4row, kind="stable")
        return col4row[order].astype(np.int64), order.astype(np.int64)
    return np.arange(nr, dtype=np.int64), col4row.astype(np.int64)

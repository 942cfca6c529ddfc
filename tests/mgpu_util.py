"""Host-side helpers for the multi-GPU tests: a NumPy DFS pre-order of an LDD (what K3 builds on device) and a
NumPy emulation of what each rank's kernels compute in the partitioned scan routing (csrc/route_scan.cu)."""
import numpy as np


def dfs_order(st):
    """LddStruct -> (pre, sub_size): pre[i] = DFS pre-order rank of canonical cell i (donors in ascending index,
    roots in ascending index), sub_size[rank] = upstream cell count.  Mirrors finalize_cell_order() in ldd_build.cu."""
    n = st.n
    pre = np.zeros(n, np.int64)
    roots = np.flatnonzero(st.down < 0)
    off = np.concatenate([[0], np.cumsum(st.nup[roots])[:-1]])
    pre[roots] = off
    don_ptr, don_idx = st.don_ptr.astype(np.int64), st.don_idx.astype(np.int64)
    parent = np.repeat(np.arange(n), np.diff(don_ptr))
    nupd = st.nup[don_idx]
    csum = np.cumsum(nupd) - nupd
    seg_first = np.repeat(csum[np.minimum(don_ptr[:-1], max(len(csum) - 1, 0))] if len(csum) else np.zeros(0), np.diff(don_ptr))
    sib_off = csum - seg_first
    lp = st.level_ptr
    for lev in range(st.nlevels - 1, 0, -1):
        cells = st.order[lp[lev]:lp[lev + 1]].astype(np.int64)
        a, b = don_ptr[cells], don_ptr[cells + 1]
        cnt = b - a
        if cnt.sum() == 0:
            continue
        idx = np.repeat(a, cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        pre[don_idx[idx]] = pre[parent[idx]] + 1 + sib_off[idx]
    sub_size = np.empty(n, np.int64)
    sub_size[pre] = st.nup
    return pre, sub_size


def device_order_structures(st, pre):
    """(down, level) by device (DFS) index from the canonical LddStruct arrays."""
    n = st.n
    down = np.full(n, -1, np.int64)
    has = st.down >= 0
    down[pre[has]] = pre[st.down[has]]
    level = np.empty(n, np.int64)
    level[pre] = st.level
    return down, level


def trunk_select(sub_size, level, down, min_level, min_cells, bounds):
    """NumPy restatement of sphy_trunk_select (csrc/route_scan.cu): the listed cells (ascending device index), their
    upstream counts, the trunk flag, the last listed slot inside each upstream range and, per entry, the list slot of
    its receiver when both are trunk cells (-1 otherwise)."""
    sub_size = np.asarray(sub_size, np.int64)
    n = sub_size.size
    j = np.arange(n)
    rank = lambda p: np.searchsorted(bounds, p, side="right") - 1  # noqa: E731
    trunk = (sub_size >= min_cells) | (level >= min_level)
    listed = trunk | (rank(j) != rank(j + sub_size - 1))
    cell = np.flatnonzero(listed)
    is_trunk = trunk[cell]
    last = np.searchsorted(cell, cell + sub_size[cell]) - 1
    parent = down[cell]
    pslot = np.searchsorted(cell, np.maximum(parent, 0))
    ok = (parent >= 0) & (pslot < cell.size)
    ok[ok] &= cell[pslot[ok]] == parent[ok]
    ok &= is_trunk
    ok[ok] &= is_trunk[pslot[ok]]
    return cell, sub_size[cell], is_trunk, last, np.where(ok, pslot, -1)


def rank_begin(values_local, sub_size_local, pub_local, listed_local):
    """What sphy_route_scan_begin leaves behind on one rank: the float64 range sums of its ordinary cells (NaN at the
    listed cells, which are finished from the published prefixes) and the block it publishes."""
    v = values_local.astype(np.float64)
    n = v.size
    S = np.cumsum(v)
    Sex = S - v
    r = np.arange(n)
    e = r + sub_size_local - 1
    acc = np.full(n, np.nan)
    loc = e < n
    acc[loc] = S[e[loc]] - Sex[r[loc]]
    acc[listed_local] = np.nan
    send = np.concatenate([[S[-1]], S[pub_local]])
    return acc, send


def trunk_finish(plan, gathered, is_trunk, last, parent_slot):
    """k_trunk_drift + k_trunk_apply (csrc/route_scan.cu): exact range sums E of the listed cells from the published
    blocks of all ranks (`gathered[q]` = [range total, prefixes at plan['pub'][q]]), the local rounding terms g of the
    trunk cells, their subtree sums D over the list, and the REAL4 result rnd(rnd(E) + D)."""
    world = len(gathered)
    base = np.concatenate([[0.0], np.cumsum([g[0] for g in gathered])])[:world]

    def prefix(rank, slot):
        out = np.zeros(rank.size)
        for q in range(world):
            sel = rank == q
            out[sel] = base[q] + np.asarray(gathered[q])[1 + slot[sel]]
        return out

    E = prefix(plan["end_rank"], plan["end_slot"]) - prefix(plan["before_rank"], plan["before_slot"])
    r = E.astype(np.float32).astype(np.float64)
    X = E.copy()
    has = parent_slot >= 0
    np.add.at(X, parent_slot[has], (r - E)[has])                 # every trunk donor hands over rnd(E_t), not E_t
    g = np.where(is_trunk, X.astype(np.float32).astype(np.float64) - r, 0.0)
    G = np.concatenate([[0.0], np.cumsum(g)])
    D = np.where(is_trunk, G[last + 1] - G[np.arange(E.size)], 0.0)
    return (r + D).astype(np.float32)

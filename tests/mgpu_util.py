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


def rank_begin(values_local, sub_size_local, me, stride):
    """What sphy_route_scan_begin leaves behind on one rank: local accumulations (NaN where the range ends on another
    rank) and the buffer it publishes."""
    v = values_local.astype(np.float64)
    n = v.size
    S = np.cumsum(v)
    Sex = S - v
    r = np.arange(n)
    e = r + sub_size_local - 1
    acc = np.full(n, np.nan)
    loc = e < n
    acc[loc] = S[e[loc]] - Sex[r[loc]]
    send = np.zeros(stride)
    send[0] = S[-1]
    send[1:1 + len(me["pub_e"])] = S[me["pub_e"]]
    partial = S[-1] - Sex[me["remote_r"]]            # r .. end of this rank's range
    return acc, send, partial


def rank_finish(acc, partial, me, rank, gathered):
    """sphy_route_scan_finish: trunk cells whose range ends on another rank."""
    acc = acc.copy()
    for i, r in enumerate(me["remote_r"]):
        q, slot = int(me["remote_rank"][i]), int(me["remote_slot"][i])
        s = partial[i] + sum(gathered[u][0] for u in range(rank + 1, q)) + gathered[q][1 + slot]
        acc[r] = s
    return acc

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


def trunk_select(sub_size, level, down, min_level, min_cells, bounds, cut_cell=None):
    """NumPy restatement of sphy_trunk_select / sphy_trunk_select_cuts (csrc/route_scan.cu): the listed cells (ascending
    device index), their upstream counts, the trunk flag, the last listed slot inside each upstream range and, per entry,
    the list slot of its receiver when both are trunk cells (-1 otherwise).  With `cut_cell` (lake / reservoir cells,
    ascending) every cell whose upstream range contains one of them is listed as well."""
    sub_size = np.asarray(sub_size, np.int64)
    n = sub_size.size
    j = np.arange(n)
    rank = lambda p: np.searchsorted(bounds, p, side="right") - 1  # noqa: E731
    trunk = (sub_size >= min_cells) | (level >= min_level)
    listed = trunk | (rank(j) != rank(j + sub_size - 1))
    if cut_cell is not None and len(cut_cell):
        cut = np.asarray(cut_cell, np.int64)
        nxt = np.searchsorted(cut, j)                                # first structure cell at or after j
        listed |= (nxt < cut.size) & (cut[np.minimum(nxt, cut.size - 1)] <= j + sub_size - 1)
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


def reslake_rule(S):
    """Stand-in for lakes.QLake / reservoirs.QRes in the host-logic emulation (the real functions are per-structure device
    code, tested on the GPU): a fixed fraction of the storage leaves per day."""
    return (np.float32(0.05) * S).astype(np.float32)


class ReslakeEmulation:
    """What sphy_route_reslake_begin / _finish compute on `world` ranks (csrc/route_scan.cu, csrc/route.cu), in NumPy: the
    plain scan on every rank, the exchange block with the per-structure extras, and the replicated structure update + the
    correction of the listed cells.  The storage update of day t (it needs the discharge at the structures' donors, which
    only exists after the listed cells are finished) is applied at the head of day t + 1 from values that travel in that
    day's exchange block, or by flush()."""

    def __init__(self, sub_size, level, down, bounds, cut_cell, donors, qfrac, kx, vol, S0, rule=reslake_rule):
        self.rule = rule                              # storage [ns] -> outflow [ns] (the stand-in, or the port's QLake / QRes)
        from sphy_b200 import multigpu
        self.b = np.asarray(bounds, np.int64)
        self.world = self.b.size - 1
        self.sub_size = np.asarray(sub_size, np.int64)
        self.cell, self.sub, _, _, _ = trunk_select(self.sub_size, level, down, 2 ** 31 - 1, 2 ** 62, self.b, cut_cell)
        self.plan = multigpu.plan_trunk(self.cell, self.sub, self.b)
        self.rl = multigpu.plan_reslake(self.cell, self.sub, cut_cell, donors, self.b)
        self.qfrac = qfrac
        self.kx, self.omk, self.vol = np.float32(kx), np.float32(1 - kx), np.float32(vol)
        self.S = S0.astype(np.float32).copy()
        self.ns = len(cut_cell)
        self.Qout = np.zeros(self.ns, np.float32)
        self.Qin = np.zeros(self.ns, np.float32)
        n = self.sub_size.size
        self.q = np.zeros(n, np.float32)             # QRAold, m3/s
        self.qday = np.zeros(n, np.float32)
        self.pending = False
        self.x_cell = np.where(self.rl["x_rank"] >= 0, self.b[np.maximum(self.rl["x_rank"], 0)] + self.rl["x_local"], -1)

    def _extras(self, totr):
        """Every rank writes the values of the cells it owns; the reader takes each slot from its owner's block."""
        x = np.zeros(self.ns * 9)
        has = self.x_cell >= 0
        is_totr = (np.arange(self.ns * 9) % 9) == 8
        x[has & ~is_totr] = self.qday[self.x_cell[has & ~is_totr]]
        x[has & is_totr] = totr[self.x_cell[has & is_totr]]
        return x

    def _apply_lag(self, x):
        v = x.reshape(self.ns, 9)[:, :8]
        s = np.zeros(self.ns)
        for q in range(8):
            s = s + v[:, q]                          # REAL8 sum in donor order, like k_rl_post
        self.Qin = s.astype(np.float32)
        self.S = ((self.S - self.Qout) + self.Qin).astype(np.float32)
        self.pending = False

    def flush(self):
        if self.pending:
            self._apply_lag(self._extras(np.zeros_like(self.q)))

    def step(self, totr, exchange=None):
        """One day.  `exchange(blocks, stride)` stands for the all-gather of the ranks' blocks (default: in-process); a block is
        [range total, prefixes at the rank's published positions, padding, the per-structure extras this rank owns]."""
        f32 = np.float32
        xg = self._extras(totr)
        nx = self.ns * 9
        stride = int(self.plan["stride"]) + nx
        # the material of EVERY cell, structure cells included (advanced_routing.py:159 zeroes them): what they add to a range
        # sum is taken out again with E(k) of the structure, which holds the same terms
        m = (self.vol * totr).astype(np.float32)
        blend = lambda f, qold: (self.omk * f + ((qold * f32(24)) * f32(3600)) * self.kx).astype(np.float32)  # noqa: E731
        sends = []
        for p in range(self.world):
            lo, hi = int(self.b[p]), int(self.b[p + 1])
            mine = self.cell[(self.cell >= lo) & (self.cell < hi)] - lo
            acc, send = rank_begin(m[lo:hi], self.sub_size[lo:hi], self.plan["pub"][p], mine)
            ordinary = ~np.isnan(acc)
            qd = blend(np.nan_to_num(acc).astype(np.float32), self.q[lo:hi])
            self.qday[lo:hi][ordinary] = qd[ordinary]
            block = np.zeros(stride)
            block[:send.size] = send
            block[stride - nx:] = np.where(self.rl["x_rank"] == p, xg, 0.0)          # what k_rl_part_gather leaves on rank p
            sends.append(block)
        if exchange is not None:
            sends = exchange(sends, stride)
        # ---- after the exchange: the same on every rank; every extra is read from its owner's block
        x = np.zeros(nx)
        for i in np.flatnonzero(self.rl["x_rank"] >= 0):
            x[i] = sends[self.rl["x_rank"][i]][stride - nx + i]
        if self.pending:
            self._apply_lag(x)
        xt = x.reshape(self.ns, 9)[:, 8].astype(np.float32)
        self.S = (self.S + self.vol * xt).astype(np.float32)                # advanced_routing.py:136-138
        self.Qout = self.rule(self.S)
        world = self.world
        base = np.concatenate([[0.0], np.cumsum([g[0] for g in sends])])[:world]

        def prefix(rank, slot):
            out = np.zeros(rank.size)
            for r in range(world):
                sel = rank == r
                out[sel] = base[r] + np.asarray(sends[r])[1 + slot[sel]]
            return out

        P = self.plan
        E = prefix(P["end_rank"], P["end_slot"]) - prefix(P["before_rank"], P["before_slot"])
        ptr, idx = self.rl["entry_ptr"], self.rl["entry_idx"]
        # a structure keeps everything of its upstream range (nested structures included) and hands on its outflow
        W = E[self.rl["cut_slot"]] - self.Qout.astype(np.float64)
        F = E - np.array([W[idx[ptr[s]:ptr[s + 1]]].sum() for s in range(E.size)])
        qd = blend(F.astype(np.float32), self.q[self.cell])
        ec = self.rl["entry_cut"]
        qd[ec >= 0] = self.Qout[ec[ec >= 0]]
        self.qday[self.cell] = qd
        self.q = (self.qday / f32(86400)).astype(np.float32)
        self.pending = True
        return self.q

"""numpy model of the arithmetic the sm_100a chamfer kernel performs (csrc/dtfill_k2_chamfer.cuh) -- TEST CODE.

It mirrors, lane for lane, what one warp does for one task (a band of rows of one frame): packed 32-bit
keys ``dist:11 | order:4 | label:17``, the 7-candidate stencil as unsigned minima, the in-lane sequential
(min,+) scan, the cross-lane Hillis-Steele carry in a widened ``dist:15 | label:17`` form, and the carry
application.  Stencil candidates use even order values so that a stored key may keep a residual order bit
(0/1, left by the carry application, which is not cleared) without changing any comparison.  tests/test_kernel_model.py checks it against the oracle; it exists so the kernel's packing,
tie-breaking and band/halo logic are verified on the CPU before a GPU is involved.
"""
from __future__ import annotations

import numpy as np

DSH, OSH = 21, 17
LMASK = np.uint64((1 << 17) - 1)
ORDMASK = np.uint64(15 << 17)
KEYMASK = np.uint64(0xFFFFFFFF)
D1 = np.uint64(1 << DSH)


def _clr(k):
    return k & ~ORDMASK & KEYMASK


def _add(k, cost, order):
    r = k + np.uint64((cost << DSH) | (order << OSH))
    assert (r >> np.uint64(32)).max() == 0, "dist field overflowed 11 bits"
    return r


# (dy, dx, cost) in OpenCV's comparison order
FWD = [(-2, -1, 3), (-2, +1, 3), (-1, -2, 3), (-1, -1, 2), (-1, 0, 1), (-1, +1, 2), (-1, +2, 3)]
BWD = [(+2, +1, 3), (+2, -1, 3), (+1, +2, 3), (+1, +1, 2), (+1, 0, 1), (+1, -1, 2), (+1, -2, 3)]


def _row_scan(c, ppl, order_scan, reverse, init_key, clamp_dist):
    """c: uint64 [32*ppl] stencil results (order bits may be set).  Returns final cleared row."""
    lanes = 32
    v = c.reshape(lanes, ppl).copy()
    if reverse:
        v = v[::-1, ::-1].copy()            # mirror so the scan always runs towards higher index
    U = np.empty_like(v)
    U[:, 0] = _clr(v[:, 0])
    step = np.uint64((1 << DSH) | (order_scan << OSH))
    for i in range(1, ppl):
        U[:, i] = _clr(np.minimum(v[:, i], U[:, i - 1] + step))
    # cross-lane carry in the widened form dist:14 | label:18
    e = U[:, ppl - 1]
    E = ((e >> np.uint64(DSH)) << np.uint64(OSH)) | (e & LMASK)
    d = 1
    while d < lanes:
        other = E.copy()
        other[d:] = E[:-d]                  # shfl_up: lanes < d keep their own value
        t = other + np.uint64((d * ppl) << OSH)
        take = (t | LMASK) < E
        E = np.where(take, t, E)
        d *= 2
    cin = np.empty_like(E)
    cin[1:] = E[:-1]
    cin[0] = (np.uint64(clamp_dist) << np.uint64(OSH))
    cd = np.minimum(cin >> np.uint64(OSH), np.uint64(clamp_dist))
    cin_key = (cd << np.uint64(DSH)) | np.uint64(1 << OSH) | (cin & LMASK)
    cin_key[0] = (np.uint64(clamp_dist) << np.uint64(DSH)) | np.uint64(1 << OSH)
    out = np.empty_like(U)
    for i in range(ppl):
        t = cin_key + np.uint64((i + 1) << DSH)
        assert (t >> np.uint64(32)).max() == 0
        out[:, i] = np.minimum(U[:, i], t)            # residual order bit 0/1 is kept
    if reverse:
        out = out[::-1, ::-1]
    return out.reshape(-1).copy()


def chamfer_band(src: np.ndarray, rank: np.ndarray, lo: int, hi: int, ppl: int):
    """src bool [H,W] source map, rank int [H,W] (1-based raster rank of sources), rows [lo,hi) processed
    as if they were the whole image.  Returns (dist [hi-lo,W], label [hi-lo,W])."""
    H, W = src.shape
    Wp = 32 * ppl
    assert W <= Wp
    init = H + W + 8
    assert 2 * H + W + ppl + 12 <= 2047 and int(rank.max()) < (1 << 17)
    clamp = 2047 - ppl - 1
    INIT = np.uint64(init << DSH)
    n = hi - lo
    pad = np.zeros((n, Wp), bool); pad[:, :W] = src[lo:hi]
    rk = np.zeros((n, Wp), np.uint64); rk[:, :W] = rank[lo:hi]
    colpad = np.arange(Wp) >= W

    def shifted(row, dx):
        ext = np.concatenate([np.full(2, INIT, np.uint64), row, np.full(2, INIT, np.uint64)])
        return ext[2 + dx: 2 + dx + Wp]

    F = np.empty((n, Wp), np.uint64)
    A = np.full(Wp, INIT, np.uint64); Bp = A.copy()          # rows y-1 and y-2
    for y in range(n):
        c = None
        for o, (dy, dx, cost) in enumerate(FWD):
            cand = _add(shifted(A if dy == -1 else Bp, dx), cost, 2 * o)
            c = cand if c is None else np.minimum(c, cand)
        c = np.where(pad[y], rk[y], c)
        row = _row_scan(c, ppl, 14, False, INIT, clamp)
        row = np.where(colpad, INIT, row)
        F[y] = row
        Bp, A = A, row
    R = np.empty((n, Wp), np.uint64)
    A = np.full(Wp, INIT, np.uint64); Bp = A.copy()          # rows y+1 and y+2
    for y in range(n - 1, -1, -1):
        c = F[y].copy()
        for o, (dy, dx, cost) in enumerate(BWD):
            cand = _add(shifted(A if dy == 1 else Bp, dx), cost, 2 * (o + 1))
            c = np.minimum(c, cand)
        c = _clr(c)
        row = _row_scan(c, ppl, 1, True, INIT, clamp)
        row = np.where(colpad, INIT, row)
        R[y] = row
        Bp, A = A, row
    dist = (R >> np.uint64(DSH)).astype(np.int64)[:, :W]
    label = (R & LMASK).astype(np.int64)[:, :W]
    dist = np.where(dist >= init, 65533, dist)
    return dist, label


# ---- coarse planning bound (guaranteed upper bound on the row maximum of dt) --------------------------------
def coarse_row_bound(src: np.ndarray, ch: int, cw: int, halves: bool = False) -> np.ndarray:
    """Upper bound U[y] >= max_x dt(y,x) from a CH x CW cell-occupancy grid: an exact anisotropic city-block
    distance on the cell grid (vertical step ch, horizontal step cw) plus the in-cell slack.

    ``halves`` (evaluated, NOT what k1b_scan_compact does -- see profiles/r02_experiments.txt): occupancy per HALF cell
    (cw/2 columns).  A cell whose two halves
    both hold a source has a source column within cw/2 - 1 of every column of its own span, so a pixel k cells away
    horizontally is at most cw*k + cw/2 - 1 columns from one of its sources (instead of cw*k + cw - 1): such a cell
    starts the distance propagation at 0, a cell with one occupied half at cw/2, and the in-cell slack that is added at
    the end is (ch - 1) + (cw/2 - 1).  Valid, but a cell row's bound is the maximum over its cells and stays where it was."""
    H, W = src.shape
    nh, nw = -(-H // ch), -(-W // cw)
    BIG = 1 << 20
    D = np.full((nh, nw), BIG, np.int64)
    hw = cw // 2
    for cy in range(nh):
        for cx in range(nw):
            blk = src[cy * ch:(cy + 1) * ch, cx * cw:(cx + 1) * cw]
            if not halves or cw % 2:
                if blk.any():
                    D[cy, cx] = 0
            else:
                left, right = blk[:, :hw].any(), blk[:, hw:].any()
                if left and right:
                    D[cy, cx] = 0
                elif left or right:
                    D[cy, cx] = hw
    for cy in range(nh):
        for cx in range(nw):
            if cy: D[cy, cx] = min(D[cy, cx], D[cy - 1, cx] + ch)
            if cx: D[cy, cx] = min(D[cy, cx], D[cy, cx - 1] + cw)
    for cy in range(nh - 1, -1, -1):
        for cx in range(nw - 1, -1, -1):
            if cy < nh - 1: D[cy, cx] = min(D[cy, cx], D[cy + 1, cx] + ch)
            if cx < nw - 1: D[cy, cx] = min(D[cy, cx], D[cy, cx + 1] + cw)
    slack = (ch - 1) + ((hw - 1) if (halves and cw % 2 == 0) else (cw - 1))
    cellmax = D.max(axis=1) + slack
    return np.repeat(cellmax, ch)[:H]


# ---- rows above the first source row (kernel k3_sky) -----------------------------------------------------------
def sky_rows_closed_form(dt: np.ndarray, lbl: np.ndarray, S: int):
    """Rows [0,S) of (dt, lbl) from rows S and S+1 alone, for a frame whose first source row f satisfies S+1 <= f:
    dt(y,x) = dt(S,x) + (S-y); the label follows t(x) diagonal steps down the distance profile of row S towards
    its valley column, then straight down (see csrc/dtfill_k3_sky.cuh)."""
    H, W = dt.shape
    g = dt[S].astype(np.int64)
    INF = np.int64(1) << 40
    right = np.concatenate([g[1:], [INF]])
    left = np.concatenate([[INF], g[:-1]])
    s = np.where(right == g - 1, 1, np.where(left == g - 1, -1, 0))
    t = np.zeros(W, np.int64)
    for x in range(W):
        xx = x
        while s[xx] != 0:
            xx += s[x]
            t[x] += 1
    odt, ol = dt.copy(), lbl.copy()
    xs = np.arange(W)
    for y in range(S):
        j = (S - y + 1) >> 1
        valley = t < j
        row = np.where(valley, S, y + 2 * j)
        col = np.where(valley, xs + s * t, xs + s * j)
        ol[y] = lbl[row, col]
        odt[y] = g + (S - y)
    return odt, ol


def numpy_pairwise_sum(a, T=None):
    """np.add.reduce over a contiguous 1-D array, restated (numpy/_core/src/umath/loops_utils.h.src pairwise sum): the
    order csrc/dtfill_k4_exact.cuh reproduces on the GPU.  T: accumulation dtype (default: a's)."""
    import numpy as np
    a = np.asarray(a)
    T = T or a.dtype.type

    def rec(lo, n):
        if n < 8:
            r = T(0.0)
            for i in range(n):
                r = T(r + a[lo + i])
            return r
        if n <= 128:
            r = [a[lo + j] for j in range(8)]
            i = 8
            while i < n - (n % 8):
                for j in range(8):
                    r[j] = T(r[j] + a[lo + i + j])
                i += 8
            res = T(T(T(r[0] + r[1]) + T(r[2] + r[3])) + T(T(r[4] + r[5]) + T(r[6] + r[7])))
            while i < n:
                res = T(res + a[lo + i])
                i += 1
            return res
        n2 = n // 2
        n2 -= n2 % 8
        return T(rec(lo, n2) + rec(lo + n2, n - n2))

    return rec(0, len(a))

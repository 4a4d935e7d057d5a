"""Test-only stand-in for the per-band device kernels: computes the band boundary summaries that
dtb_flowacc_band / dtb_hand emit (include/dtb200.h) with plain NumPy walks, so that the band driver's
exchange + boundary-graph solve (descriptools_b200/bands.py) can be exercised on CPU with gloo."""
import numpy as np

import oracle

OFF = {1: (0, 1), 2: (1, 1), 4: (1, 0), 8: (1, -1), 16: (0, -1), 32: (-1, -1), 64: (-1, 0), 128: (-1, 1)}
KIND_RIVER, KIND_FAIL, KIND_EXIT = 1, 2, 3


def _code_at(d8, above, below, r, c):
    rows, cols = d8.shape
    if c < 0 or c >= cols:
        return 0
    if 0 <= r < rows:
        return int(d8[r, c])
    if r == -1 and above is not None:
        return int(above[c])
    if r == rows and below is not None:
        return int(below[c])
    return 0


def _fed(d8, above, below, side, c):
    rows = d8.shape[0]
    if side == 0:
        return d8[0, c] != 0 and (_code_at(d8, above, below, -1, c - 1) == 2 or _code_at(d8, above, below, -1, c) == 4
                                  or _code_at(d8, above, below, -1, c + 1) == 8)
    return d8[rows - 1, c] != 0 and (_code_at(d8, above, below, rows, c - 1) == 128 or _code_at(d8, above, below, rows, c) == 64
                                     or _code_at(d8, above, below, rows, c + 1) == 32)


def flowacc_summary(d8, above, below):
    """int64 [6, cols] as Band.flowacc_summary: exit_above, exit_below, term_above, term_below, d8 rows"""
    rows, cols = d8.shape
    acc0, _ = oracle.flow_accumulation(d8)  # band alone: moves that leave the band are dropped
    out = np.zeros((6, cols), np.int64)
    out[2:4] = -2
    for side, r, halo in ((0, 0, above), (1, rows - 1, below)):
        if halo is None:
            continue
        for c in range(cols):
            code = int(d8[r, c])
            if code in OFF:
                dr, dc = OFF[code]
                if (dr < 0 if side == 0 else dr > 0) and _code_at(d8, above, below, r + dr, c + dc) != 0:
                    out[side, c] = acc0[r, c] + 1
            if _fed(d8, above, below, side, c):
                rr, cc, t = r, c, -1
                for _ in range(rows * cols + 1):
                    code = int(d8[rr, cc])
                    if code not in OFF:
                        break
                    dr, dc = OFF[code]
                    r2, c2 = rr + dr, cc + dc
                    if _code_at(d8, above, below, r2, c2) == 0:
                        break
                    if r2 < 0 or r2 >= rows:
                        t = ((0 if r2 < 0 else 1) << 30) | cc
                        break
                    rr, cc = r2, c2
                out[2 + side, c] = t
    out[4], out[5] = d8[0], d8[rows - 1]
    return out


def hand_summary(d8, above, below, river, dem, acc, row_offset):
    """int64 [8, cols] as Band.hand_summary"""
    rows, cols = d8.shape
    out = np.zeros((8, cols), np.int64)
    out[1] = out[5] = -100
    zbits = np.zeros((2, cols), np.float64)
    for side, r, halo in ((0, 0, above), (1, rows - 1, below)):
        if halo is None:
            continue
        for c in range(cols):
            if not _fed(d8, above, below, side, c):
                continue
            rr, cc, nc, nd = r, c, 0, 0
            state = KIND_FAIL << 62
            for _ in range(rows * cols + 1):
                if river[rr, cc] == 1:
                    state = (KIND_RIVER << 62) | (nd << 47) | (nc << 32) | (rr * cols + cc)
                    out[4 * side + 1, c] = (row_offset + rr) * cols + cc
                    zbits[side, c] = dem[rr, cc]
                    out[4 * side + 3, c] = acc[rr, cc]
                    break
                code = int(d8[rr, cc])
                if code not in OFF:
                    break
                dr, dc = OFF[code]
                r2, c2 = rr + dr, cc + dc
                if _code_at(d8, above, below, r2, c2) == 0:
                    break
                if dr != 0 and dc != 0:
                    nd += 1
                else:
                    nc += 1
                if r2 < 0 or r2 >= rows:
                    state = (KIND_EXIT << 62) | (nd << 47) | (nc << 32) | (1 << 31) | ((0 if r2 < 0 else 1) << 30) | c2
                    break
                rr, cc = r2, c2
            out[4 * side, c] = np.array(state, dtype=np.uint64).astype(np.int64)
    out[2], out[6] = zbits[0].view(np.int64), zbits[1].view(np.int64)
    return out

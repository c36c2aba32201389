"""Host emulation of loops_b200/csrc/spmv_merge2.cuh (the second-generation
merge-path SpMV kernel), statement by statement: row-end flags, 16-byte chunks in
staged coordinates (skew = s.y & 3), boundary masks, the rare 257th chunk, the
segmented fold over chunks in sequence order, the per-tile carry and the fix-up
pass. Same fp32 operations in the same order as the kernel, so on a GPU box the
kernel's y can be compared with this BIT FOR BIT on any input (the order of the
adds inside a row is part of what is emulated).

Test infrastructure only (used by tests/test_merge2_emul.py and the GPU tests).
"""
import numpy as np

TILE = 1024
F32 = np.float32


def merge_coords(row_end, T, A, items=TILE, diagonals=None):
    """S(b * items), b = 0..M: first i in [max(d-A,0), min(d,T)) with
    row_end[i] > d - i - 1 (reference util/search.hxx:34-60)."""
    W = T + A
    if diagonals is None:
        M = (W + items - 1) // items
        diagonals = [min(b * items, W) if b < M else W for b in range(M + 1)]
        # the kernel searches at b*items even past W for the last entry: same answer (T, A)
    out = []
    for d in diagonals:
        lo, hi = max(d - A, 0), min(d, T)
        while lo < hi:
            mid = (lo + hi) >> 1
            if row_end[mid] <= d - mid - 1:
                lo = mid + 1
            else:
                hi = mid
        out.append((min(lo, T), d - lo))
    return out


def spmv_merge2(off, idx, val, x, threads=128, ell_pitch=None, coords=None, y_init=None):
    """y of the emulated kernel. `off` CSR offsets (ignored when ell_pitch is
    given: then tile_end(i) = (i+1)*pitch and idx < 0 marks padding)."""
    if ell_pitch is None:
        T = len(off) - 1
        A = int(off[-1])
        row_end = np.asarray(off[1:], np.int64)
    else:
        T = len(idx) // ell_pitch if ell_pitch else 0
        A = len(idx)
        row_end = (np.arange(T, dtype=np.int64) + 1) * ell_pitch
    accumulate = y_init is not None          # the kernel's ACCUM mode: y += A x
    y = np.array(y_init, F32) if accumulate else np.full(T, np.nan, F32)
    if T == 0:
        return y
    if A == 0:
        if not accumulate:
            y[:] = 0
        return y
    if coords is None:
        coords = merge_coords(row_end, T, A)
    M = len(coords) - 1
    CH = TILE // (4 * threads)
    nchunk_threads = CH * threads            # 256
    carry_row = np.zeros(M, np.int64)
    carry_val = np.zeros(M, F32)
    val = np.asarray(val, F32)
    x = np.asarray(x, F32)
    for j in range(M):
        (sx, sy), (ex, ey) = coords[j], coords[j + 1]
        nt, na = ex - sx, ey - sy
        skew = sy & 3
        lo, hi = skew, skew + na
        flags = np.zeros(TILE + 8, bool)
        rowat = np.zeros(TILE + 8, np.int64)
        for i in range(nt):
            e = int(row_end[sx + i]) - sy
            b = int(row_end[sx + i - 1]) - sy if i > 0 else 0
            if e > b:
                pos = skew + e - 1
                assert not flags[pos]
                flags[pos] = True
                rowat[pos] = i
            elif not accumulate:
                y[sx + i] = 0.0

        def product(pos):
            if pos < lo or pos >= hi:
                return F32(0)
            a = sy + (pos - skew)
            c = idx[a]
            xx = x[c] if (ell_pitch is None or c >= 0) else F32(0)
            return F32(val[a]) * F32(xx)

        # per chunk: (any, head, head_row, tail), stores of rows inside the chunk
        run = F32(0)          # fold over the sequence so far: tail since the last closed row
        nchunks = nchunk_threads + (1 if hi > 4 * nchunk_threads else 0)
        # warp aggregates are folded warp by warp in the kernel; the association of the adds is
        # (((lane0 + lane1) + ...) within Hillis-Steele order -- emulate that order exactly
        c = 0
        seq = []   # per thread-chunk: (got, head, head_row, tail)
        for c in range(nchunk_threads):
            p0 = 4 * c
            r = F32(0)
            got = False
            hd, hrow = F32(0), 0
            span = range(4)
            for q in span:
                r = F32(r + product(p0 + q))
                if p0 < hi and flags[p0 + q]:
                    if not got:
                        hd, hrow, got = r, rowat[p0 + q], True
                    else:
                        rr = sx + rowat[p0 + q]
                        y[rr] = F32(y[rr] + r) if accumulate else r
                    r = F32(0)
            if c == nchunk_threads - 1 and hi > 4 * nchunk_threads:     # spill chunk continues the last thread
                for q in range(4):
                    r = F32(r + product(4 * nchunk_threads + q))
                    if flags[4 * nchunk_threads + q]:
                        if not got:
                            hd, hrow, got = r, rowat[4 * nchunk_threads + q], True
                        else:
                            rr = sx + rowat[4 * nchunk_threads + q]
                            y[rr] = F32(y[rr] + r) if accumulate else r
                        r = F32(0)
            seq.append((got, hd, hrow, r))
        # A warp owns 32*CH consecutive chunks = CH "slots" of 32 lanes. Per slot: Hillis-Steele
        # segmented inclusive scan of the tails; the warp chains its slots in registers
        # (wv, wa); the per-warp aggregates are folded in order after the barrier.
        nwarps = threads // 32
        warp_out = []      # per warp: (wv, wa, [per slot: (lanes, v, gots, acc_val, acc_any)])
        for w in range(nwarps):
            wv, wa = F32(0), False
            slots = []
            for u in range(CH):
                base = (w * CH + u) * 32
                lanes = seq[base: base + 32]
                gots = [l[0] for l in lanes]
                v = [l[3] for l in lanes]
                start = []
                for L in range(32):
                    if gots[L]:
                        start.append(L)
                    else:
                        below = [k for k in range(L) if gots[k]]
                        start.append(below[-1] if below else 0)
                d = 1
                while d < 32:
                    nv = list(v)
                    for L in range(32):
                        if L - d >= start[L]:
                            nv[L] = F32(v[L - d] + v[L])
                    v = nv
                    d <<= 1
                slots.append((lanes, v, gots, wv, wa))
                if any(gots):
                    wv, wa = v[31], True
                else:
                    wv = F32(wv + v[31])
            warp_out.append((wv, wa, slots))
        run = F32(0)
        for w in range(nwarps):
            before = run
            wv, wa, slots = warp_out[w]
            for lanes, v, gots, acc_val, acc_any in slots:
                for L in range(32):
                    got, hd, hrow, _ = lanes[L]
                    if got:
                        excl = v[L - 1] if L > 0 else F32(0)
                        carry_in = excl
                        if not any(gots[:L]):
                            prior = acc_val if acc_any else F32(before + acc_val)
                            carry_in = F32(prior + excl)
                        out = F32(carry_in + hd)
                        y[sx + hrow] = F32(y[sx + hrow] + out) if accumulate else out
            run = wv if wa else F32(run + wv)
        carry_row[j] = sx + nt
        carry_val[j] = run
    # fix-up (spmv_merge_fixup_kernel): carries naming the same row are summed left to right, then added
    j = 0
    while j < M:
        row = carry_row[j]
        k = j
        acc = carry_val[j]
        while k + 1 < M and carry_row[k + 1] == row:
            k += 1
            acc = F32(acc + carry_val[k])
        if row < T:
            y[row] = F32(acc + y[row])
        j = k + 1
    return y

"""Host emulation of the band-tiled SpMV kernel (loops_b200/csrc/spmv_tiled.cuh)
over the image the plan builder produces: walks every consumer warp's stream
step by step with the kernel's acquire/release rules for the x ring, decodes
every entry back to (row, col, value), and accumulates y in the kernel's order.
It checks the format invariants the kernel relies on and raises AssertionError
when one is broken. Test infrastructure only."""
import ctypes as C

import numpy as np

FLAG = np.uint32(0x80000000)


def build_image(lib, rows, cols, off, idx, val, geometry):
    from loops_b200 import _lib
    off = np.ascontiguousarray(off, np.int32)
    idx = np.ascontiguousarray(idx, np.int32)
    val = np.ascontiguousarray(val, np.float32)
    geo = (C.c_int32 * 6)(*geometry)
    h = C.c_void_p()
    rc = lib.loopsb_tiled_image_build_host(rows, cols, off.ctypes.data, idx.ctypes.data if idx.size else None,
                                           val.ctypes.data if val.size else None, C.byref(geo), C.byref(h))
    if rc != 0:
        return rc, None
    info = _lib.TiledInfo()
    assert lib.loopsb_tiled_image_info(h, C.byref(info)) == 0
    ps = [C.c_void_p() for _ in range(6)]
    assert lib.loopsb_tiled_image_arrays(h, *[C.byref(p) for p in ps]) == 0
    g = info.as_dict()
    ns = g["nb"] * g["q"] * g["warps"]

    def arr(p, n, ct, dt):
        if n == 0:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,)).astype(dt, copy=True)

    img = {
        "g": g,
        "steps": arr(ps[0], (g["total_steps"] + g["es"]) * 256, C.c_uint32, np.uint32).reshape(-1, 256),
        "stream_base": arr(ps[1], ns + 1, C.c_int32, np.int64),
        "fs": arr(ps[2], ns * g["nband"], C.c_uint16, np.int64).reshape(ns, g["nband"]),
        "le": arr(ps[3], ns * g["nband"], C.c_uint16, np.int64).reshape(ns, g["nband"]),
        "blk_begin": arr(ps[4], g["nb"] + 1, C.c_int32, np.int64),
        "warp_begin": arr(ps[5], g["nb"] * g["q"] * (g["warps"] + 1), C.c_int32, np.int64).reshape(-1, g["warps"] + 1),
    }
    lib.loopsb_tiled_image_free(h)
    return 0, img


def emulate(img, x, rows, cols):
    """Returns (y, triples) where triples = sorted array of decoded (row, col, value bits)."""
    g = img["g"]
    nb, q, W, cb, xb, es = g["nb"], g["q"], g["warps"], g["cb"], g["xb"], g["es"]
    rb, rw, cq, nband = g["rb"], g["rw"], g["cq"], g["nband"]
    zero_slot = xb * cb
    blk = img["blk_begin"]
    assert blk[0] == 0 and blk[-1] == rows and np.all(np.diff(blk) >= 0) and np.all(np.diff(blk) <= rb)
    y_parts = np.zeros((q, rows), np.float32)
    out_r, out_c, out_v = [], [], []
    lane_of = np.repeat(np.arange(32), 4)
    for cta in range(nb * q):
        rbi, qi = divmod(cta, q)
        ys = np.zeros(rb + 1, np.float32)
        wb = img["warp_begin"][cta]
        assert wb[0] == 0 and wb[-1] == blk[rbi + 1] - blk[rbi] and np.all(np.diff(wb) >= 0)
        for w in range(W):
            s_id = cta * W + w
            base, end = img["stream_base"][s_id], img["stream_base"][s_id + 1]
            fs, le = img["fs"][s_id], img["le"][s_id]
            acq = 0
            resident = [-1] * xb     # band sitting in each ring slot (as this warp sees it)
            held = []                # bands acquired and not yet released, oldest first
            done = set()             # bands released
            # (round-1 format: whole prefetch groups, (end - base) % es == 0; the kernel no longer needs it)

            def acquire(keep):
                nonlocal acq
                old = resident[acq % xb]
                assert old < 0 or old in done, ("x ring would deadlock: slot still held", cta, w, acq, old)
                resident[acq % xb] = acq
                if keep:
                    held.append(acq)
                else:
                    done.add(acq)
                acq += 1

            for s in range(end - base):
                words = img["steps"][base + s]
                ids, vals = words[:128], words[128:].view(np.float32)
                meta = 0
                for lane in range(32):
                    meta |= int(ids[4 * lane] >> np.uint32(31)) << lane
                assert not np.any(ids.reshape(32, 4)[:, 1:] & FLAG), "bit 31 of slots 1..3 must be clear"
                lead, plen = (meta >> 1) & 0xFFF, (meta >> 13) & 0xF
                pat, nrel = (meta >> 17) & 0xFF, (meta >> 25) & 0xF
                for _ in range(lead):
                    acquire(False)
                for i in range(plen):
                    acquire(bool((pat >> i) & 1))
                # cross-check with the band tables the builder derived the word from
                assert acq == int(np.sum(fs <= s)), ("acquire count differs from the band table", cta, w, s)
                rel = 0   # entries may only reference held bands (checked below)
                lc_enc = (ids & np.uint32(0xFFFF)).astype(np.int64)
                lr = ((ids >> np.uint32(16)) & np.uint32(0x7FFF)).astype(np.int64)
                dirty = bool(meta & 1)
                pad = lr == rb
                assert np.all(lc_enc[pad] == zero_slot) and np.all(vals[pad] == 0), "bad padding entry"
                real = ~pad
                assert np.all(lc_enc[real] < zero_slot)
                k = lc_enc // cb
                band = np.array([resident[int(kk)] if kk < xb else -1 for kk in k])
                assert np.all(band[real] >= 0)
                assert all(int(bb) in held for bb in np.unique(band[real])), ("entry of a band not held", cta, w, s)
                col = qi * cq + band * cb + (lc_enc - k * cb)
                row = blk[rbi] + lr
                assert np.all(col[real] < cols) and np.all(row[real] < rows)
                assert np.all((lr[real] >= wb[w]) & (lr[real] < wb[w + 1])), "row outside the warp's sub-block"
                xv = np.where(real, x[np.where(real, col, 0)], np.float32(0)).astype(np.float32)
                prod = (vals * xv).astype(np.float32)
                # rows that the fast path cannot combine: not one contiguous slot
                # range, or a range over three or more lanes
                slots_of = {}
                for i in range(128):
                    if not pad[i]:
                        slots_of.setdefault(int(lr[i]), []).append(i)
                bad_rows = {r for r, sl in slots_of.items() if sl[-1] - sl[0] + 1 != len(sl)}
                long_rows = {r for r, sl in slots_of.items()
                             if r not in bad_rows and sl[-1] // 4 - sl[0] // 4 >= 2}
                assert dirty == bool(bad_rows), ("dirty bit mismatch", cta, w, s)
                assert bool(meta & (1 << 29)) == bool(long_rows), ("long-run bit mismatch", cta, w, s)
                if not dirty:
                    # fast path: run sums inside a lane; the parts of a run in later lanes are
                    # handed back to the lane where the run starts, which updates y once
                    for r, sl in slots_of.items():
                        lane_sums = {}
                        for i in sl:
                            lane_sums[i // 4] = np.float32(lane_sums.get(i // 4, np.float32(0)) + prod[i]) \
                                if (i // 4) in lane_sums else prod[i]
                        lanes = sorted(lane_sums)
                        rest = np.float32(0)
                        for ln in reversed(lanes[1:]):
                            rest = np.float32(lane_sums[ln] + rest)
                        ys[r] = np.float32(ys[r] + np.float32(lane_sums[lanes[0]] + rest))
                else:
                    # dirty step: slot by slot, equal rows applied in lane order
                    for j in range(4):
                        for i in range(j, 128, 4):
                            if not pad[i]:
                                ys[lr[i]] = np.float32(ys[lr[i]] + prod[i])
                out_r.append(row[real]); out_c.append(col[real]); out_v.append(vals[real].view(np.uint32))
                # bands whose last entry is in this step are released, oldest first
                assert nrel <= len(held)
                for _ in range(nrel):
                    bnd = held.pop(0)
                    assert le[bnd] == s + 1, ("released a band before its last entry", cta, w, s, bnd)
                    done.add(bnd)
                assert all(le[bnd] > s + 1 for bnd in held), ("band kept past its last entry", cta, w, s)
            assert not held
            while acq < nband:
                acquire(False)
            assert done == set(range(nband))
        y_parts[qi, blk[rbi]: blk[rbi + 1]] = ys[: blk[rbi + 1] - blk[rbi]]
    y = y_parts[0].copy()
    for qq in range(1, q):
        y = (y + y_parts[qq]).astype(np.float32)
    tr = np.stack([np.concatenate(out_r) if out_r else np.zeros(0, np.int64),
                   np.concatenate(out_c) if out_c else np.zeros(0, np.int64),
                   np.concatenate(out_v).astype(np.int64) if out_v else np.zeros(0, np.int64)], axis=1)
    return y, tr

"""Lane-level numpy model of pv_kernel_ring.cuh (frame 1024: one warp per channel pair; frame
2048: two warps per pair; every thread owns 16 complex points of the half-size FFT).

TEST INFRASTRUCTURE: the model restates, array-of-threads style, exactly the index arithmetic
of the CUDA kernel (ring-order FFT, padded exchange slots, run-based peak / owner scan, two
ordered shift sub-steps, stale upper bins) so that the design can be checked against the CPU
oracle without a GPU (tests/test_ring_model.py), including the shared-memory bank-conflict
degree of every access pattern.

Key identity the kernel relies on: with both rings aligned to the time
cursor t (frame sample n lives at ring index (n + t) mod N), the FFT of the ring-ordered
windowed frame is U[k] = X[k] exp(-j 2 pi k t / N), so the per-region rotation
exp(j 2 pi (bs - b) t / N) of shiftPeaks (phase-vocoder.js:155-170) becomes the identity:
V[b + delta] += U[b], and the inverse FFT comes out in ring order as well.
"""
from __future__ import annotations

import numpy as np

ROWS = False             # gather_lanes: destinations in row order (lane l handles bins TP r + l) instead of runs
MIDDLE = "scatter"       # "gather": prototype of the next middle (destination-order gather, see gather_shift)
EXACT = False            # experiment: exact first-writer classification (no zero fill when contracting)
F32 = np.float32
C64 = np.complex64
INVALID_DELTA = 0x3000


def xslot(k):
    """16-byte slot of spectrum bin k (one pad slot per 16 bins)"""
    k = np.asarray(k)
    return k + (k >> 4)


class Geo:
    def __init__(self, n: int):
        assert n in (256, 512, 1024, 2048, 4096)
        self.N = n
        self.M = n // 2                  # complex FFT length
        self.NB = self.M + 1
        self.TP = n // 32                # threads per channel pair
        self.R1 = self.M // 64           # radix of the first pass (8 or 16)
        self.KS = self.M // 8            # stride between the outputs of one last-pass butterfly
        self.NJ = n // 128               # ring blocks of 128 samples
        self.XSLOTS = self.M + self.M // 16 + 2
        # exchange slot of element i (0..63) of row k1: rows of 64 slots at stride 65; frame 512 (radix-4
        # first pass: a quarter-warp of pass 3 spans two k2 groups) pads every group of 8 and uses stride 74
        self.RS = 65 if self.R1 >= 8 else 74 if self.R1 == 4 else 76
        self.GP = 0 if self.R1 >= 8 else 1
        self.EX_SLOTS = self.RS * (self.R1 - 1) + 64 + 8 * self.GP
        # bytes per pair: frame 512 keeps two pairs per warp 16 banks apart (32-bit plane accesses)
        self.PAIR_BYTES = max(self.XSLOTS, self.EX_SLOTS) * 16
        if self.TP < 32:                 # pairs of one warp: 16 (two pairs) / 8 (four pairs) banks apart
            self.PAIR_BYTES += ((64 if self.TP == 16 else 32) - self.PAIR_BYTES) % 128
        self.T = np.arange(self.TP)

    def exs(self, k1, i):
        i = np.asarray(i)
        return self.RS * np.asarray(k1) + i + self.GP * (i >> 3)


def tables(g: Geo, overlaps: int):
    n = g.N
    i = np.arange(n)
    win = (0.5 * (1 - np.cos(2 * np.pi * i / n))).astype(F32)
    win_out = (win * F32(1.0 / (2 * n)) * F32(1.0 / overlaps)).astype(F32)
    tw = np.exp(-2j * np.pi * i / n).astype(C64)
    for q in range(4):
        tw[q * n // 4] = [1, -1j, -1, 1j][q]
    return win, win_out, tw


def delta_table(g: Geo, pf32: np.float32) -> np.ndarray:
    """round(p * pitchFactor) - p (pv:125, round half up of the exact f64 product), or INVALID
    when the shifted peak is beyond nb (pv:127 `break`: this and all later peaks are dropped)."""
    p = np.arange(g.NB + 1, dtype=np.float64)
    ps = np.floor(p * np.float64(pf32) + 0.5)
    d = (ps - p).astype(np.int64)
    d[ps > g.NB] = INVALID_DELTA
    return d


class Conflicts:
    """worst wavefront count (relative to the minimum) per warp access, per pattern name"""

    def __init__(self):
        self.worst = {}
        self.sparse = {}

    gather_log = None       # experiments: byte addresses (-1: lane idle) per 64-bit gather access
    scatter_log = None      # experiments: list of (destination bins, active lanes) per 32-bit scatter access
    pair_bytes = 0          # frame 512: a warp holds two pairs, the second one this many bytes further

    def note(self, name, byte_addr, width):
        byte_addr = np.asarray(byte_addr).reshape(-1)
        if byte_addr.size == 16:
            byte_addr = np.concatenate([byte_addr, byte_addr + self.pair_bytes])
        elif byte_addr.size == 8:
            byte_addr = np.concatenate([byte_addr + q * self.pair_bytes for q in range(4)])
        for w0 in range(0, byte_addr.size, 32):          # one warp at a time
            self._warp(name, byte_addr[w0:w0 + 32], width)

    def note_sparse(self, name, word_idx):
        """32-bit accesses with idle lanes (index -1): wavefronts per warp access accumulated as
        (count, total) in self.sparse[name]"""
        word_idx = np.asarray(word_idx).reshape(-1)
        for w0 in range(0, word_idx.size, 32):
            a = word_idx[w0:w0 + 32]
            a = a[a >= 0]
            if a.size == 0:
                continue
            banks = {}
            for x in set(int(v) for v in a):
                banks.setdefault(x % 32, set()).add(x)
            wf = max(len(v) for v in banks.values())
            c = self.sparse.setdefault(name, [0, 0])
            c[0] += 1; c[1] += wf

    def _warp(self, name, byte_addr, width):
        assert byte_addr.size == 32
        group = {4: 32, 8: 16, 16: 8}[width]
        total = 0
        for g in range(0, 32, group):
            a = byte_addr[g:g + group]
            words = set()
            for x in a:
                for w in range(width // 4):
                    words.add(int(x) // 4 + w)
            banks = {}
            for w in words:
                banks.setdefault(w % 32, set()).add(w)
            total += max((len(v) for v in banks.values()), default=0)
        ideal = width // 4
        self.worst[name] = max(self.worst.get(name, 0), total / ideal)


def dft(x, inv=False):
    """x[R, ...] complex -> DFT over axis 0 (forward e^-, inverse e^+, unnormalised)"""
    r = x.shape[0]
    w = np.exp((2j if inv else -2j) * np.pi * np.outer(np.arange(r), np.arange(r)) / r)
    return np.tensordot(w, x.astype(np.complex128), axes=(1, 0)).astype(C64)


def gather_shift(g: Geo, xfull, peaks, dtab, contract, cf: Conflicts | None = None):
    """PROTOTYPE of a middle that touches the spectrum once: the shifted spectrum
    in DESTINATION order.  Every peak writes one descriptor at the first destination bin of its region
    (delta, overlap with the previous region, source end); a thread owns 16 consecutive destination
    bins, finds for each the latest descriptor at or below it (forward fill, carry from the threads
    below) and reads its one or two sources: no zero fill, no second sub-step, and the result stays in
    registers until the 128-bit store that feeds the unsplit.

    xfull: complex [M + N/8 + 1], bins 0 .. M + N/8 (the stale extension rebuilt once);
    peaks: ascending peak bins; returns Y[NB].  Equivalent to shiftPeaks (pv:119-173) for pitch
    factors in [0.75, 64]: at most two regions reach one destination bin."""
    N, M, NB = g.N, g.M, g.NB
    src_end_max = xfull.shape[0]                       # sources beyond the first stale level do not exist
    desc = {}                                          # destination bin -> (delta, overlap, source end, delta_prev)
    prev_delta = None
    for i, p in enumerate(peaks):
        delta = int(dtab[p])
        if delta == INVALID_DELTA:
            break                                      # pv:127: this and all later peaks are dropped
        s = 0 if i == 0 else p - (p - peaks[i - 1]) // 2
        e = N if i == len(peaks) - 1 else p + -(-(peaks[i + 1] - p) // 2)
        e = min(e, src_end_max)
        ds = s + delta
        overlap = 0 if prev_delta is None else max(0, prev_delta - delta)
        first = max(ds, 0)
        if first < NB and e + delta > 0:
            assert first not in desc, "two regions start at one destination bin"
            desc[first] = (delta, overlap - (first - ds), e, prev_delta)
        prev_delta = delta
    Y = np.zeros(NB, C64)
    cur = None
    loads = []                                         # (destination, source) pairs, for the conflict count
    for d in range(NB):
        if d in desc:
            cur = (d,) + desc[d]
        if cur is None:
            continue
        d0, delta, overlap, e, dprev = cur
        b = d - delta
        if b < e:
            Y[d] = xfull[b]; loads.append((d, b))
        if d - d0 < overlap:
            Y[d] += xfull[d - dprev]; loads.append((d, d - dprev))
    if cf is not None:
        # thread t owns destinations 16 t .. 16 t + 15 (bin M: one more load by the last thread); X as
        # 8 bytes (re, im) per channel and bin, 16-byte slots padded every 16 bins: first / second sources
        for which in (0, 1):
            for e_ in range(16):
                addr = []
                for tt in range(g.TP):
                    srcs = [b for (d, b) in loads if d == 16 * tt + e_]
                    addr.append(xslot(srcs[which]) * 16 if len(srcs) > which else -1)
                cf.gather_log.append(np.array(addr))
    return Y


def gather_shift_lanes(g: Geo, xfull, mask, prev_before, next_after, dtab, contract, cf: Conflicts | None = None):
    """gather_shift restated thread by thread, the way a kernel would do it (blueprint for the CUDA):

    A. every thread walks the peaks of its own SOURCE run (bits of `mask`; the nearest peaks below /
       above its run come from the ballot exchange the kernel already has) and stores one 32-bit
       descriptor per peak at the first destination bin of the peak's region, in a zeroed array
       D[nb]: contracting (delta <= 0, regions overlap, never leave gaps): delta | overlap << 16;
       expanding (no overlaps, gaps): delta | source_end << 16.  Bit 31 marks "present".
    B. every thread owns the DESTINATION run 16 tp .. 16 tp + 15 (the last thread also bin M): it
       reads its 16 descriptors, gets the latest descriptor below its run (word and start bin) from
       the nearest lower thread that has one (ballot + two shuffles), forward-fills, and gathers
       Y[d] = X[d - delta] (+ X[d - delta - overlap] on the first `overlap` bins of a region).
    """
    N, M, NB, TP = g.N, g.M, g.NB, g.TP
    PRESENT = 1 << 31
    D = np.zeros(NB + 15, np.int64)                    # zeroed every call (8 x 128-bit stores per lane and pair)
    n_src = xfull.shape[0]
    desc_st = []
    for L in range(TP):                                # ---- A: descriptors --------------------------------
        pp = int(prev_before[L]); has_prev = pp >= 0
        bits = [e for e in range(16) if (int(mask[L]) >> e) & 1]
        for j, e in enumerate(bits):
            p = 16 * L + e
            nxt = 16 * L + bits[j + 1] if j + 1 < len(bits) else int(next_after[L])
            has_next = nxt < 20000
            delta = int(dtab[p])
            if delta != INVALID_DELTA:
                s = p - ((p - pp) >> 1) if has_prev else 0
                e_end = min(p + ((nxt - p + 1) >> 1) if has_next else N, n_src)
                dprev = int(dtab[pp]) if has_prev else delta
                ds = s + delta
                first = max(ds, 0)
                if first < NB and e_end + delta > 0:
                    if contract:
                        ovl = max(0, dprev - delta)
                        assert ds >= 0 or ovl == 0
                        word = PRESENT | ((delta + 4096) & 0x1FFF) | (ovl << 16)
                    else:
                        word = PRESENT | ((delta + 4096) & 0x1FFF) | (e_end << 16)
                    assert D[first] == 0, "two regions start at one destination bin"
                    D[first] = word
                    desc_st.append((L, first * 4))
            pp, has_prev = p, True
    if ROWS:
        # ---- B': destinations in ROW order: lane l of the pair handles d = TP r + l, r = 0 .. 15 (and bin
        # M in one more row).  Consecutive lanes read consecutive sources (delta is piecewise constant), so
        # the 64-bit gathers are conflict free; the latest descriptor at or below d comes from a ballot of
        # the row (nearest set lane at or below) or, if the row has none below, from the carry of the
        # rows before.  Y[TP r + l] is exactly what the unsplit of thread l needs for bins l + 64 j
        # (r = 2 j at frame 1024), and bins M - l - 64 j sit in lane TP - l of row 15 - 2 j: one shuffle
        # per value, the shifted spectrum never goes through shared memory.
        Y = np.zeros(NB, C64)
        carry, cstart = 0, 0
        for r in range((NB + TP - 1) // TP):
            drow = TP * r + np.arange(TP)
            words = np.array([int(D[d]) if d < NB else 0 for d in drow])
            present = words != 0
            addr = [np.full(TP, -1), np.full(TP, -1)]
            for l in range(TP):
                d = int(drow[l])
                if d >= NB:
                    continue
                below = [k for k in range(l + 1) if present[k]]          # ballot & ((2 << l) - 1), then clz
                cur, dstart = (int(words[below[-1]]), TP * r + below[-1]) if below else (carry, cstart)
                if not cur:
                    continue
                delta = (cur & 0x1FFF) - 4096
                hi = (cur >> 16) & 0x7FFF
                b = d - delta
                if contract:
                    Y[d] = xfull[b]; addr[0][l] = 8 * (b + 2 * (b >> 4))
                    if d - dstart < hi:
                        Y[d] += xfull[b - hi]; addr[1][l] = 8 * (b - hi + 2 * ((b - hi) >> 4))
                elif b < hi:
                    Y[d] = xfull[b]; addr[0][l] = 8 * (b + 2 * (b >> 4))
            if present.any():
                k = int(np.nonzero(present)[0][-1])
                carry, cstart = int(words[k]), TP * r + k
            if cf is not None and cf.gather_log is not None:
                cf.gather_log.extend(addr)
        return Y
    # ---- B: forward fill + gather ---------------------------------------------------------------
    own = D[:16 * TP].reshape(TP, 16)
    last_pos = np.array([max([e for e in range(16) if own[L, e]], default=-1) for L in range(TP)])
    Y = np.zeros(NB, C64)
    loads = []
    for L in range(TP):
        lower = [l for l in range(L) if last_pos[l] >= 0]          # ballot + shuffle in the kernel
        cur, dstart = (int(own[lower[-1], last_pos[lower[-1]]]), 16 * lower[-1] + int(last_pos[lower[-1]])) if lower else (0, 0)
        dests = list(range(16 * L, 16 * L + 16)) + ([M] if L == TP - 1 else [])
        for d in dests:
            w = int(D[d])
            if w:
                cur, dstart = w, d
            if not cur:
                continue
            delta = (cur & 0x1FFF) - 4096
            hi = (cur >> 16) & 0x7FFF
            b = d - delta
            if contract:
                Y[d] = xfull[b]; loads.append((L, d, b))
                if d - dstart < hi:
                    Y[d] += xfull[b - hi]; loads.append((L, d, b - hi))
            elif b < hi:
                Y[d] = xfull[b]; loads.append((L, d, b))
    if cf is not None and cf.gather_log is not None:
        # X per channel as 8-byte (re, im) slots, one pad slot per 16 bins: byte address 8 (b + (b >> 4))
        for which in (0, 1):
            for e_ in range(16):
                addr = np.full(TP, -1)
                for L in range(TP):
                    srcs = [b for (l, d, b) in loads if l == L and d == 16 * L + e_]
                    if len(srcs) > which:
                        addr[L] = 8 * (srcs[which] + (srcs[which] >> 4))
                cf.gather_log.append(addr)
    return Y


def desc_layout(g: Geo):
    """bit fields of the 32-bit region descriptors of gather_rows_kernel: T' in the low TB bits, the overlap
    c in the next CB bits, -delta (signed) above"""
    TB = int(np.log2(g.N))                 # T' = clamp(T, 0, 2^TB - 2) + 1 <= N - 1 (never zero: marks "present")
    CB = 9                                  # overlap of two regions while contracting: <= (nb / 4) + 1 bins
    return TB, CB


def gather_rows_kernel(g: Geo, xfull, mask, prev_before, next_after, dtab, contract, cf: Conflicts | None = None,
                       stats: dict | None = None):
    """BLUEPRINT of the gather middle of pv_kernel_ring.cuh (PVB_RING_GATHER), thread by thread.

    Regions of influence in DESTINATION space.  Region i (peak p_i, delta_i, sources [s_i, s_{i+1}),
    s_i = p_i - floor((p_i - p_{i-1}) / 2), s_0 = 0) lands on [s_i + delta_i, s_{i+1} + delta_i).  With
    q_i = s_i + min(delta_i, delta_{i-1}) and T_i = s_i + max(delta_i, delta_{i-1}) the destination axis is
    cut at the q_i, and inside [q_i, q_{i+1}):
        d <  T_i :  contracting: Y[d] = X[d - delta_i] + X[d - delta_{i-1}]   (the two regions overlap)
                    expanding:   Y[d] = 0                                      (gap between the regions)
        d >= T_i :  Y[d] = X[d - delta_i]
    (pitch factors >= 0.75: never more than two regions on one bin).  A peak whose shifted position is
    beyond nb (pv:127, only while expanding) has delta = INVALID (huge): the same formulas give q = end of the
    previous image and T = "never", i.e. zeros to the end, and later peaks fall outside [0, nb).

    A. Every thread walks the peaks of its SOURCE run and stores one 32-bit descriptor
       (T' | c << TB | -delta << (TB + CB)) at D[max(q_i, 0)], and copies of it at every bin = 0 or 1
       (mod RW) strictly inside (q_i, q_{i+1}) (RW = lanes of a pair in one warp): then every aligned
       group of RW destination bins starts with a descriptor, at its first AND second bin.
    B. Destinations are taken in the order the Hermitian pre-pass wants them: step j, thread tp holds
       Y[tp + KS j] and Y[M - tp - KS j].  The RW lanes of a warp cover an aligned group (lane 0 of the
       group holds an unrelated bin that is = 0 (mod RW), so its own descriptor is always there): the
       latest descriptor at or below a bin comes from a ballot and one shuffle, nothing is carried from
       step to step, and the shifted spectrum never exists in shared memory.
    """
    N, M, NB, TP, KS = g.N, g.M, g.NB, g.TP, g.KS
    RW = min(TP, 32)
    TB, CB = desc_layout(g)
    TMAX = (1 << TB) - 2
    DSZ = ((NB + RW - 1) // RW) * RW + RW
    D = np.zeros(DSZ, np.int64)
    n_src = xfull.shape[0]
    st_addr = []                                        # descriptor stores: (iteration, lane, word index)
    any_peak = bool((np.asarray(mask) != 0).any())
    big = 1 << 20
    for L in range(TP):                                # ---- A: descriptors --------------------------------
        pp = int(prev_before[L]); has_prev = pp >= 0
        bits = [e for e in range(16) if (int(mask[L]) >> e) & 1]
        for it, e in enumerate(bits):
            p = 16 * L + e
            nxt = 16 * L + bits[it + 1] if it + 1 < len(bits) else int(next_after[L])
            has_next = nxt < 20000
            delta = int(dtab[p])
            dprev = int(dtab[pp]) if has_prev else (delta if contract else 0)
            s = p - ((p - pp) >> 1) if has_prev else 0
            q = s + min(delta, dprev)
            T = s + max(delta, dprev)
            c = abs(dprev - delta) if contract else 0
            if has_next:
                dn = int(dtab[nxt])
                qn = (nxt - ((nxt - p) >> 1)) + min(dn, delta)
            else:
                qn = big
            if contract:
                assert delta <= dprev and c < (1 << CB), (delta, dprev)
                assert q >= 0 or c == 0
                # never three regions on one bin: the next image starts after this overlap zone
                assert qn >= T, "three regions overlap (pitch factor < 0.75?)"
            qs = max(q, 0)
            if qs >= NB:
                continue
            Tp = min(max(T, 0), TMAX) + 1
            ndl = 0 if delta == INVALID_DELTA else -delta   # (an invalid peak's descriptor is all "zero zone")
            word = Tp | (c << TB) | ((ndl & ((1 << (32 - TB - CB)) - 1)) << (TB + CB))
            assert word != 0 and D[qs] == 0, "two regions start at one destination bin"
            D[qs] = word
            st_addr.append((it, L, qs))
            lim = min(qn, NB)
            r = qs // RW
            while RW * r < lim:
                for x in (RW * r, RW * r + 1):
                    if qs < x < lim:
                        assert D[x] == 0
                        D[x] = word
                        st_addr.append((it + 100 * (x - RW * r + 1), L, x))
                r += 1
            pp, has_prev = p, True
    if not any_peak:
        zero_forever = (TMAX + 1)
        for r in range(0, DSZ // RW):
            D[RW * r] = zero_forever; D[RW * r + 1] = zero_forever
    if stats is not None:
        stats.setdefault("desc_stores", []).append(len(st_addr))
        stats.setdefault("peak_iters", []).append(max((bin(int(m)).count("1") for m in mask), default=0))

    # ---- B: gather in the order of the Hermitian pre-pass ---------------------------------------------
    sh = TB + CB
    def fields(w):
        Tp = w & ((1 << TB) - 1)
        c = (w >> TB) & ((1 << CB) - 1)
        nd = w >> sh
        if nd >= 1 << (31 - sh):
            nd -= 1 << (32 - sh)
        return Tp, c, -nd
    Y = np.zeros(NB, C64)
    T_ = np.arange(TP)
    gathers = []
    for j in range(8):
        dk = np.where(T_ == 0, (KS // 2 + KS * j) if j < 4 else KS * (j - 4), T_ + KS * j)
        gathers.append(("k", dk)); gathers.append(("m", M - dk))
    gathers.append(("k", np.full(TP, M // 2)))
    for kind, dd in gathers:
        words = D[dd]
        present = words != 0
        addr0 = np.full(TP, -1); addr1 = np.full(TP, -1)
        for seg in range(0, TP, RW):                   # one warp (or the pair's part of a warp) at a time
            for l in range(RW):
                L = seg + l
                d = int(dd[L])
                if kind == "k":
                    cand = [k for k in range(l + 1) if present[seg + k]]
                    assert cand, "no descriptor at or below (k)"
                    src = cand[-1]
                else:
                    cand = [k for k in range(l, RW) if present[seg + k]]
                    assert cand, "no descriptor at or below (m)"
                    src = cand[0]
                if L != seg + src and l != 0:
                    # the lane that serves must hold a bin at or below ours, in our group
                    assert int(dd[seg + src]) <= d and d - int(dd[seg + src]) < RW, (kind, L, src)
                Tp, c, delta = fields(int(words[seg + src]))
                zone = d + 1 < Tp
                if contract:
                    b = d - delta
                    v = xfull[b]; addr0[L] = b
                    if zone:
                        v = v + xfull[b - c]; addr1[L] = b - c
                else:
                    v = 0
                    if not zone:
                        b = d - delta
                        assert 0 <= b < n_src
                        v = xfull[b]; addr0[L] = b
                # reference semantics: a region never reads beyond the sources that exist
                if 0 <= d < NB:
                    Y[d] = v
        if cf is not None:
            cf.note_sparse("gather_ld", addr0)
            if contract:
                cf.note_sparse("gather_ld2", addr1)
            cf.note_sparse("desc_ld", dd)
    if cf is not None:
        its = sorted(set(i for i, _, _ in st_addr))
        for it in its:
            a = np.full(TP, -1)
            for i, L, x in st_addr:
                if i == it:
                    a[L] = x
            cf.note_sparse("desc_st", a)
    return Y


def walk_block_tree(n, pos):
    """(L, r, s, o): slot `pos` of fft.js's realTransform output holds DFT_L(xw[r m + s])[o] (bundle:394-438 writes
    only outputs 0 .. L/2 of every length-L block of the radix-4 recursion)"""
    l0 = 4 if int(np.log2(n)) % 2 == 0 else 2
    L, r, s, o = n, 1, 0, pos
    while L > l0 and o > L // 2:
        q = L // 4
        sb = o // q
        o -= sb * q
        s += r * sb
        r *= 4
        L = q
    return L, r, s, o


def deep_shift(g: Geo, Xc, ring, win, tw, t, mask, prev_before, next_after, dtab, d_last, pf):
    """The middle of the kernel's DEEP instances for one channel.  Xc: spectrum slots (2x scaled, ring order);
    ring: the channel's history ring [N] (sample n of the frame at (n + t) mod N).  Returns Y[0 .. nb)."""
    N, M, NB, TP, T = g.N, g.M, g.NB, g.TP, g.T
    Xv = Xc[xslot(np.arange(M + 1))]
    full = np.concatenate([Xv, np.conj(Xv[1:M][::-1])]).astype(C64)              # Hermitian extension, k < N
    n = np.arange(N)
    xw = (ring[(n + t) % N].astype(F32) * win).astype(F32)                       # windowed frame, frame order

    def stale(q):                                                                # slot N/2 + q in the kernel's units
        if q == 0:
            return Xv[M]
        L, r, s, o = walk_block_tree(N, M + q)
        if r == 4:                                                               # quarters s = 2 (q <= N/8) and s = 3 (q >= N/4)
            idx = o + L * np.arange(4)
            return C64(0.25 * np.sum(full[idx] * np.conj(tw[(s * idx) % N])))
        if r == 16:                                                              # second level: 16 terms of the spectrum
            idx = o + L * np.arange(16)
            return C64(np.sum(full[idx] * np.conj(tw[(s * idx) % N])) / 16)
        m = np.arange(L)                                                         # deeper: L <= N/64 frame samples
        acc = np.sum(xw[r * m + s].astype(C64) * tw[(o * r * m) % N])
        return C64(2 * acc * tw[((M + q) * t) % N])                              # frame order -> ring order, 2x scale

    # owner (ordinal of the peak) and destination of every bin of every run
    peaks = np.array([int(16 * L + e) for L in range(TP) for e in range(16) if (int(mask[L]) >> e) & 1])
    b = np.arange(16 * TP)
    hi = np.searchsorted(peaks, b, side="right")                                 # first peak above b
    lo = hi - 1
    take_next = (lo < 0) | ((hi < len(peaks)) & (peaks[np.minimum(hi, len(peaks) - 1)] - b <= b - peaks[np.maximum(lo, 0)]))
    ordinal = np.where(take_next, hi, lo)
    dest = b + dtab[peaks[ordinal]]
    Y = np.zeros(NB + 1, C64)
    ok = (dest >= 0) & (dest < NB)
    if pf >= 0.5:
        for colour in range(3):                                                  # pairwise disjoint inside a sub-step
            sel = ok & (ordinal % 3 == colour)
            assert len(set(dest[sel].tolist())) == int(sel.sum()), "two writers for one bin inside a coloured sub-step"
            Y[dest[sel]] += Xv[b[sel]]
    else:
        np.add.at(Y, dest[ok], Xv[b[ok]])
    for q in range(0, N // 3 + 2):                                               # the last region beyond bin N/2 - 1
        d = M + q + d_last
        if 0 <= d < NB:
            Y[d] += stale(q)
    return Y[:NB]


def step(g: Geo, hist2, acc2, inblk, t, pf32, hop, cf: Conflicts | None = None, capture: dict | None = None):
    """One process() call of one channel pair.  hist2/acc2: [N][2] float32 rings (modified in
    place), inblk: [2][hop] or None (paused), t = timeCursor (multiple of hop).  Returns out[2][hop]."""
    N, M, NB, TP, R1, KS, NJ, T = g.N, g.M, g.NB, g.TP, g.R1, g.KS, g.NJ, g.T
    # frames 256 / 512 also take hops that are odd multiples of 64 (whether a register is in the first
    # or the second half of its 128-sample block is a compile-time fact there: nl >= 32 <=> h >= 32 / TP)
    assert (hop % 128 == 0 or (hop % 64 == 0 and TP <= 16)) and hop <= N // 2
    R = N // hop
    win, win_out, tw = tables(g, R)
    w64s = N // 64                                  # W_64^x = tw[w64s * x]
    dtab = delta_table(g, pf32)
    contract = bool(pf32 < 1.0)
    t = int(t) % N
    nblk = hop // 128
    ex = np.zeros((g.EX_SLOTS, 2), C64)            # exchange: slot -> (ch0, ch1) complex
    nls = [T + TP * h for h in range(16 // R1)]   # pass-1 butterflies of a thread (n mod 64)

    # ---- load, window, forward pass 1 (DFT over the 128-sample blocks, stride 64) ---------------
    # Registers are indexed by FRAME block f (so the role of every register -- history, new input,
    # emitted head, zero tail -- is static); it sits in ring block j = (f + toff) mod NJ.  The DFT over
    # ring blocks of the rotated register array is the DFT over f times W_R1^{toff k1} (shift theorem),
    # which is folded into the pass-1 twiddle: W_M^{(n + 64 toff) k1}.
    toff = (t // 128) % NJ
    half = (t // 64) & 1                            # rings rotated by half a block (hop 64 only)
    w32s, w128s = N // 32, N // 128                 # W_32^x = tw[w32s * x], W_128^x = tw[w128s * x]
    sg = -1.0 if (toff & 1) else 1.0
    if R1 == 32:
        # frame 4096: a thread holds 16 of the 32 frame blocks of its column n, those of parity s
        # (f = 2 f' + s), and does a 16-point DFT over f'; the radix-2 step that completes the
        # 32-point DFT over f is done by the READER in pass 2, which holds rows k' and k' + 16 anyway:
        #   row (s, k') = E_s[k'] W_M^{(n + 64 toff) k'} W_32^{s k'}
        #   X[k'] = row(0) + row(1),  X[k' + 16] = (row(0) - row(1)) W_128^n (-1)^toff
        assert nblk % 2 == 0
        n_ = T & 63; s_ = T >> 6
        z = np.zeros((16, TP, 2), C64)
        for fp in range(16):
            f = 2 * fp + s_
            j = (f + toff) % NJ
            i = 2 * n_ + 128 * j
            if fp >= 16 - nblk // 2:                # new block (both parities)
                si = 2 * n_ + 128 * (f - (NJ - nblk))
                if inblk is None:
                    x0 = np.zeros((TP, 2), F32); x1 = np.zeros((TP, 2), F32)
                else:
                    x0 = inblk[:, si].T.astype(F32); x1 = inblk[:, si + 1].T.astype(F32)
                hist2[i] = x0; hist2[i + 1] = x1
            else:
                x0 = hist2[i]; x1 = hist2[i + 1]
            fi = 2 * n_ + 128 * f
            w0 = win[fi][:, None]; w1 = win[fi + 1][:, None]
            z[fp] = (x0 * w0).astype(F32) + 1j * (x1 * w1).astype(F32)
        z = dft(z)
        for kp in range(16):
            twd = (tw[(2 * (n_ + 64 * toff) * kp) % N] * tw[(w32s * s_ * kp) % N]).astype(C64)
            slot = g.exs(16 * s_ + kp, n_)
            ex[slot] = z[kp] * twd[:, None]
            if cf: cf.note("p1_st", slot * 16, 16)
    for nl in nls:
        # ring column and block carry of this butterfly (half-block rotation: hop 64)
        cc = (nl + 32 * half) & 63
        carry = (nl + 32 * half) >> 6
        z = np.zeros((R1, TP, 2), C64)
        for f in range(R1):
            j = (f + toff + carry) % NJ
            i = 2 * cc + 128 * j                    # ring index of sample pair (i, i + 1)
            fi = 2 * nl + 128 * f                   # frame position of the sample pair
            isnew = fi >= N - hop                   # new block (ola:91-108); static per register
            assert isnew.all() or not isnew.any()
            if isnew.all():
                s = fi - (N - hop)
                if inblk is None:
                    x0 = np.zeros((TP, 2), F32); x1 = np.zeros((TP, 2), F32)
                else:
                    x0 = inblk[:, s].T.astype(F32); x1 = inblk[:, s + 1].T.astype(F32)
                hist2[i] = x0; hist2[i + 1] = x1
            else:
                x0 = hist2[i]; x1 = hist2[i + 1]
            w0 = win[fi][:, None]; w1 = win[fi + 1][:, None]
            z[f] = (x0 * w0).astype(F32) + 1j * (x1 * w1).astype(F32)
        z = dft(z)
        for k1 in range(R1):
            # W_M^{c k1} W_R1^{(toff + carry) k1}: the DFT over ring blocks of registers indexed by frame block
            v = z[k1] * (tw[(2 * cc * k1) % N] * tw[((N // R1) * (toff + carry) * k1) % N]).astype(C64)[:, None]
            slot = g.exs(k1, cc)
            ex[slot] = v
            if cf: cf.note("p1_st", slot * 16, 16)

    # ---- forward pass 2 (DFT over m2), in place -------------------------------------------------
    m3 = T & 7
    if R1 == 32:
        kp = T >> 3
        A0 = np.stack([ex[g.exs(kp, m3 + 8 * m2)] for m2 in range(8)])
        A1 = np.stack([ex[g.exs(16 + kp, m3 + 8 * m2)] for m2 in range(8)])
        if cf:
            for m2 in range(8):
                cf.note("p2_ld", g.exs(kp, m3 + 8 * m2) * 16, 16); cf.note("p2_ld", g.exs(16 + kp, m3 + 8 * m2) * 16, 16)
        wn = np.stack([tw[(w128s * (m3 + 8 * m2)) % N] * sg for m2 in range(8)]).astype(C64)
        x0 = dft((A0 + A1).astype(C64)); x1 = dft(((A0 - A1).astype(C64) * wn[:, :, None]).astype(C64))
        for k2 in range(8):
            w = tw[(w64s * m3 * k2) % N][:, None]
            ex[g.exs(kp, m3 + 8 * k2)] = x0[k2] * w
            ex[g.exs(16 + kp, m3 + 8 * k2)] = x1[k2] * w
    for h in range(2 if R1 < 32 else 0):
        k1 = (T >> 3) + (R1 // 2) * h
        x = np.stack([ex[g.exs(k1, m3 + 8 * m2)] for m2 in range(8)])
        if cf:
            for m2 in range(8): cf.note("p2_ld", g.exs(k1, m3 + 8 * m2) * 16, 16)
        x = dft(x)
        for k2 in range(8):
            ex[g.exs(k1, m3 + 8 * k2)] = x[k2] * tw[(w64s * m3 * k2) % N][:, None]       # W_64^{m3 k2}

    # ---- forward pass 3 (DFT over m3): outputs stay in registers -------------------------------
    kA = T.copy()
    kB = KS - T
    kB[0] = KS // 2                                 # thread 0 owns the two self-paired butterflies
    sA = g.exs(kA % R1, 8 * (kA // R1))
    sB = g.exs(kB % R1, 8 * (kB // R1))
    a = dft(np.stack([ex[sA + c] for c in range(8)]))        # a[j] = Z[kA + KS j]
    b = dft(np.stack([ex[sB + c] for c in range(8)]))        # b[j] = Z[kB + KS j]
    if cf:
        for c in range(8):
            cf.note("p3_ldA", (sA + c) * 16, 16); cf.note("p3_ldB", (sB + c) * 16, 16)

    # ---- real split (registers) -> X (2x scaled), one 16-byte slot per bin ---------------------
    # general threads: slot j pairs (a[j], b[7-j]) at k = tp + KS j.
    # thread 0: j < 4: (b[j], b[7-j]) at k = KS/2 + KS j; j >= 4: (a[j-4], a[(12-j) & 7]) at
    #           k = KS (j-4); plus the self pair k = M/2 (a[4]).
    l0 = (T == 0)[:, None]
    X = np.zeros((2, g.XSLOTS), C64)

    def split(za, zb, k, active=None):
        e = za + np.conj(zb)
        o = -1j * (za - np.conj(zb))
        tt = o * tw[k][:, None]
        xk = e + tt                                  # 2 X[k]
        xm = np.conj(e - tt)                         # 2 X[M - k]
        s1, s2 = xslot(k), xslot(M - k)
        for L in range(TP):
            if active is not None and not active[L]: continue
            X[:, s1[L]] = xk[L]; X[:, s2[L]] = xm[L]
        if cf and active is None:
            cf.note("split_st_k", s1 * 16, 16); cf.note("split_st_mk", s2 * 16, 16)

    def split_k(j):
        if j < 4: return np.where(T == 0, KS // 2 + KS * j, T + KS * j)
        return np.where(T == 0, KS * (j - 4), T + KS * j)

    for j in range(8):
        if j < 4:
            za = np.where(l0, b[j], a[j]); zb = b[7 - j]
        else:
            za = np.where(l0, a[j - 4], a[j]); zb = np.where(l0, a[(12 - j) & 7], b[7 - j])
        split(za, zb, split_k(j))
    split(a[4], a[4], np.full(TP, M // 2), active=(T == 0))

    if capture is not None:                          # the float32 spectrum (2x scaled, ring order) per channel
        capture["X"] = X[:, xslot(np.arange(NB))].copy()
    # ---- per channel: peaks, owners, shift ------------------------------------------------------
    for ch in range(2):
        Xc = X[ch]
        b0 = 16 * T
        run = np.stack([Xc[xslot(np.clip(b0 + e, 0, M + 1))] for e in range(-2, 18)])     # [20][TP]
        if cf:
            for e in range(16): cf.note("run_ld", xslot(b0 + e) * 16, 16)
        mag = (run.real.astype(F32) ** 2 + run.imag.astype(F32) ** 2).astype(F32)
        mask = np.zeros(TP, np.int64)
        for e in range(16):
            c = mag[e + 2]
            pk = (c > mag[e]) & (c > mag[e + 1]) & (c > mag[e + 3]) & (c > mag[e + 4])
            bb = b0 + e
            pk &= (bb >= 2) & (bb <= NB - 3)          # pv:98-99
            mask |= pk.astype(np.int64) << e
        xv = run[2:18].copy()                         # own 16 bins
        nz = mask != 0
        if not nz.any():
            X[ch, :] = 0                              # no peaks: shifted spectrum is zero (pv:121)
            continue
        own_last = np.where(nz, b0 + np.floor(np.log2(np.maximum(mask, 1))).astype(np.int64), -1)
        own_first = np.where(nz, b0 + np.array([(int(m) & -int(m)).bit_length() - 1 for m in mask]), -1)
        prev_before = np.full(TP, -30000); next_after = np.full(TP, 30000)
        for L in range(TP):
            lo = [l for l in range(L) if nz[l]]
            hi = [l for l in range(L + 1, TP) if nz[l]]
            if lo: prev_before[L] = own_last[lo[-1]]
            if hi: next_after[L] = own_first[hi[0]]
        p_last = own_last[np.nonzero(nz)[0][-1]]
        d_last = int(dtab[p_last])

        if float(pf32) < 0.75:
            # DEEP instances (pitch factors in [0.33, 0.75)): every stale slot the last region reaches, and the
            # scatter in three coloured sub-steps (>= 0.5) or with atomic adds (below)
            Yd = deep_shift(g, Xc.copy(), hist2[:, ch], win, tw, t, mask, prev_before, next_after, dtab, d_last, float(pf32))
            X[ch, :] = 0
            X[ch, xslot(np.arange(NB))] = Yd
            continue

        # extension: bin M and the first stale level (bundle:394-438), owned by the last peak
        ext = np.zeros((4, TP), C64)
        for i in range(4):
            q = T + TP * i
            if i == 0 or contract:
                qq = np.maximum(q, 1)
                A = Xc[xslot(qq)]; B = Xc[xslot(N // 4 + qq)]; Cc = Xc[xslot(M - qq)]; D = Xc[xslot(N // 4 - qq)]
                s = (A - B) + np.conj(Cc - D)
                v = (0.25 * s * np.conj(tw[(2 * qq) % N])).astype(C64)
                v = np.where(q == 0, Xc[xslot(M)], v)
                if not contract: v = np.where(q == 0, v, 0)
                ext[i] = v
                if cf and contract:
                    cf.note("stale_ld", xslot(qq) * 16, 16); cf.note("stale_ld_m", xslot(M - qq) * 16, 16)

        if MIDDLE in ("gather", "gather_lanes", "gather_rows2"):
            xfull = np.zeros(M + N // 8 + 1, C64)
            xfull[:M + 1] = Xc[xslot(np.arange(M + 1))]
            for i in range(4):
                q = T + TP * i
                sel_ = (q > 0) if contract else np.zeros(TP, bool)
                xfull[M + q[sel_]] = ext[i][sel_]
            peaks = [int(b0[L]) + e for L in range(TP) for e in range(16) if (int(mask[L]) >> e) & 1]
            if MIDDLE == "gather_rows2":
                Yc = gather_rows_kernel(g, xfull if contract else xfull[:M + 1], mask, prev_before, next_after,
                                        dtab, contract, cf, capture)
            elif MIDDLE == "gather_lanes":
                Yc = gather_shift_lanes(g, xfull if contract else xfull[:M + 1], mask, prev_before, next_after,
                                        dtab, contract, cf)
            else:
                Yc = gather_shift(g, xfull if contract else xfull[:M + 1], peaks, dtab, contract,
                                  cf if cf is not None and cf.gather_log is not None else None)
            X[ch, :] = 0
            X[ch, xslot(np.arange(NB))] = Yc
            continue

        # owner of every bin of the run: nearest peak, ties to the higher one (pv:132-141)
        nextv = np.zeros((16, TP), np.int64)
        Q = next_after.copy()
        for e in range(15, -1, -1):
            nextv[e] = Q
            Q = np.where((mask >> e) & 1, b0 + e, Q)
        P = prev_before.copy()
        dest = np.zeros((16, TP), np.int64); first = np.zeros((16, TP), bool)
        for e in range(16):
            bb = b0 + e
            P = np.where((mask >> e) & 1, bb, P)
            take_next = (nextv[e] - bb) <= (bb - P)
            owner = np.where(take_next, nextv[e], P)
            dest[e] = bb + dtab[owner]
            first[e] = ~take_next | (not contract)    # right half of its region, or expanding
            if EXACT and contract:
                # only the left-half bins that land on the previous region's right half are added on
                # top: the first c = delta_prev - delta_next bins of the left half.  With
                # T = 4 b + 2 - 2 next - 2 prev (even, >= 2 on a left half): collide <=> T <= 4 c
                Tq = 4 * bb + 2 - 2 * nextv[e] - 2 * P
                cq = dtab[np.maximum(P, 0)] - dtab[np.minimum(nextv[e], NB)]
                collide = take_next & (P >= 0) & (nextv[e] < 20000) & (Tq <= 4 * cq)
                first[e] = ~collide

        # every thread holds its sources in registers; zero fill, then two ordered sub-steps
        # (right halves are pairwise disjoint after the shift, and so are left halves: checked here)
        X[ch, :] = np.nan if (EXACT and contract) else 0      # EXACT: full coverage, no zero fill
        written = np.zeros(g.XSLOTS, np.int64)
        for e in range(16):                           # first sub-step: plain stores
            ok = (dest[e] >= 0) & (dest[e] < NB) & first[e]
            if cf is not None and cf.scatter_log is not None: cf.scatter_log.append((dest[e].copy(), ok.copy()))
            for L in np.nonzero(ok)[0]:
                written[xslot(dest[e][L])] += 1
                X[ch, xslot(dest[e][L])] = xv[e][L]
        for i in range(4):
            d = M + T + TP * i + d_last
            ok = (d >= 0) & (d < NB)
            for L in np.nonzero(ok)[0]:
                written[xslot(d[L])] += 1
                X[ch, xslot(d[L])] = ext[i][L]
        assert written.max() <= 1, "two first writers for one bin"
        if EXACT and contract:
            assert (written[xslot(np.arange(NB))] == 1).all(), "a bin of the shifted spectrum was never written"
            X[ch, ~np.isin(np.arange(g.XSLOTS), xslot(np.arange(NB)))] = 0
        written[:] = 0
        if contract:
            for e in range(16):                       # second sub-step: left halves add on top
                ok = (dest[e] >= 0) & (dest[e] < NB) & ~first[e]
                if cf is not None and cf.scatter_log is not None: cf.scatter_log.append((dest[e].copy(), ok.copy()))
                for L in np.nonzero(ok)[0]:
                    written[xslot(dest[e][L])] += 1
                    X[ch, xslot(dest[e][L])] += xv[e][L]
            assert written.max() <= 1, "two second writers for one bin"

    # ---- Hermitian C2R pre-pass (mirror of the split) -------------------------------------------
    def unsplit(k):
        yk = X[:, xslot(k)].T.copy(); ym = X[:, xslot(M - k)].T.copy()       # [TP][2]
        if cf and len(set(k.tolist())) > 1:          # planes of 32-bit words: word = slot
            cf.note("unsplit_ld_k", xslot(k) * 4, 4); cf.note("unsplit_ld_mk", xslot(M - k) * 4, 4)
        z0 = (k == 0)[:, None]
        yk = np.where(z0, yk.real + 0j, yk); ym = np.where(z0, ym.real + 0j, ym)
        e = yk + np.conj(ym)
        d = yk - np.conj(ym)
        pp = d * np.conj(tw[k])[:, None]
        zk = e + 1j * pp
        zmk = np.conj(e - 1j * pp)
        return zk.astype(C64), zmk.astype(C64)

    a = np.zeros((8, TP, 2), C64); b = np.zeros((8, TP, 2), C64)
    for j in range(8):
        zk, zmk = unsplit(split_k(j))
        if j < 4:
            a[j] = np.where(l0, a[j], zk); b[7 - j] = zmk
            b[j] = np.where(l0, zk, b[j])
        else:
            a[j] = np.where(l0, a[j], zk)
            b[7 - j] = np.where(l0, b[7 - j], zmk)
            a[j - 4] = np.where(l0, zk, a[j - 4])    # thread 0: (a[j-4], a[(12-j)&7])
            if j > 4: a[(12 - j) & 7] = np.where(l0, zmk, a[(12 - j) & 7])
    zh, _ = unsplit(np.full(TP, M // 2))
    a[4] = np.where(l0, zh, a[4])

    # ---- inverse pass 1 (DFT over k3 -> m3), twiddle conj(W_64^{k2 m3}) ------------------------
    a = dft(a, inv=True); b = dft(b, inv=True)
    for c in range(8):
        ex[sA + c] = a[c] * np.conj(tw[(w64s * (kA // R1) * c) % N])[:, None]
        ex[sB + c] = b[c] * np.conj(tw[(w64s * (kB // R1) * c) % N])[:, None]
    # ---- inverse pass 2 (DFT over k2 -> m2), twiddle conj(W_M^{k1 (m3 + 8 m2 + 64 toff)}) -------
    if R1 == 32:
        kp = T >> 3
        d0 = dft(np.stack([ex[g.exs(kp, m3 + 8 * k2)] for k2 in range(8)]), inv=True)
        d1 = dft(np.stack([ex[g.exs(16 + kp, m3 + 8 * k2)] for k2 in range(8)]), inv=True)
        for m2 in range(8):
            nn = m3 + 8 * m2
            tc = np.conj(tw[(2 * kp * (nn + 64 * toff)) % N])
            cw = np.conj(tw[(w128s * nn) % N]) * sg
            v1 = (d1[m2] * cw[:, None]).astype(C64)
            ex[g.exs(kp, nn)] = ((d0[m2] + v1).astype(C64) * tc[:, None]).astype(C64)
            ex[g.exs(16 + kp, nn)] = ((d0[m2] - v1).astype(C64)
                                      * (tc * np.conj(tw[(w32s * kp) % N])).astype(C64)[:, None]).astype(C64)
    for h in range(2 if R1 < 32 else 0):
        k1 = (T >> 3) + (R1 // 2) * h
        x = dft(np.stack([ex[g.exs(k1, m3 + 8 * k2)] for k2 in range(8)]), inv=True)
        for m2 in range(8):
            c2 = m3 + 8 * m2
            cr = toff + (half & (c2 < 32))           # block carry of ring column c2
            ex[g.exs(k1, c2)] = x[m2] * np.conj(tw[(2 * k1 * c2) % N] * tw[((N // R1) * cr * k1) % N]).astype(C64)[:, None]
    # ---- inverse pass 3 (DFT over k1 -> frame block f); window, overlap-add, emit ---------------
    out = np.zeros((2, hop), F32)
    if R1 == 32:
        n_ = T & 63; s_ = T >> 6
        x = dft(np.stack([ex[g.exs(16 * s_ + kp, n_)] for kp in range(16)]), inv=True)
        for fp in range(16):
            f = 2 * fp + s_
            j = (f + toff) % NJ
            i = 2 * n_ + 128 * j
            fi = 2 * n_ + 128 * f
            w0 = win_out[fi][:, None]; w1 = win_out[fi + 1][:, None]
            y0 = (x[fp].real.astype(F32) * w0).astype(F32); y1 = (x[fp].imag.astype(F32) * w1).astype(F32)
            tail = fp >= 16 - nblk // 2
            q0 = np.zeros((TP, 2), F32) if tail else acc2[i]
            q1 = np.zeros((TP, 2), F32) if tail else acc2[i + 1]
            y0 = y0 + q0; y1 = y1 + q1
            if fp < nblk // 2:
                si = 2 * n_ + 128 * f
                out[:, si] = y0.T; out[:, si + 1] = y1.T
            else:
                acc2[i] = y0; acc2[i + 1] = y1
    for nl in nls:
        cc = (nl + 32 * half) & 63
        carry = (nl + 32 * half) >> 6
        x = dft(np.stack([ex[g.exs(k1, cc)] for k1 in range(R1)]), inv=True)
        for f in range(R1):
            j = (f + toff + carry) % NJ
            i = 2 * cc + 128 * j
            fi = 2 * nl + 128 * f
            w0 = win_out[fi][:, None]; w1 = win_out[fi + 1][:, None]
            y0 = (x[f].real.astype(F32) * w0).astype(F32); y1 = (x[f].imag.astype(F32) * w1).astype(F32)
            tail = bool((fi >= N - hop).all())          # starts from zero (ola:134)
            q0 = np.zeros((TP, 2), F32) if tail else acc2[i]
            q1 = np.zeros((TP, 2), F32) if tail else acc2[i + 1]
            y0 = y0 + q0; y1 = y1 + q1
            if bool((fi < hop).all()):                  # head: emit (ola:111-118)
                s = fi
                out[:, s] = y0.T; out[:, s + 1] = y1.T
            else:
                acc2[i] = y0; acc2[i + 1] = y1
    return out


def run(signal: np.ndarray, pf: float, hop: int, cf: Conflicts | None = None, start_calls: int = 0,
        frame: int = 1024):
    """signal [2][T*hop] -> output [2][T*hop]"""
    g = Geo(frame)
    pf32 = np.float32(pf)
    hist2 = np.zeros((frame, 2), F32); acc2 = np.zeros((frame, 2), F32)
    nsteps = signal.shape[1] // hop
    out = np.zeros_like(signal, dtype=F32)
    for m in range(nsteps):
        out[:, m * hop:(m + 1) * hop] = step(g, hist2, acc2, signal[:, m * hop:(m + 1) * hop],
                                             (start_calls + m) * hop, pf32, hop, cf)
    return out

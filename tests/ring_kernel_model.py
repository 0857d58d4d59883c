"""Lane-level numpy model of pv_kernel_ring.cuh (frame 1024, one warp per channel pair).

TEST INFRASTRUCTURE: the model restates, array-of-32-lanes style, exactly the index
arithmetic of the CUDA kernel (ring-order FFT, padded exchange slots, run-based peak /
owner scan, two-pass in-place shift, stale upper bins) so that the design can be checked
against the CPU oracle without a GPU (tests/test_ring_model.py), including the shared-memory
bank-conflict degree of every access pattern.

Key identity the kernel relies on: with both rings aligned to the time
cursor t (frame sample n lives at ring index (n + t) mod N), the FFT of the ring-ordered
windowed frame is U[k] = X[k] exp(-j 2 pi k t / N), so the per-region rotation
exp(j 2 pi (bs - b) t / N) of shiftPeaks (phase-vocoder.js:155-170) becomes the identity:
V[b + delta] += U[b], and the inverse FFT comes out in ring order as well.
"""
from __future__ import annotations

import numpy as np

N, M, NB = 1024, 512, 513
LANES = np.arange(32)
F32 = np.float32
C64 = np.complex64

INVALID_DELTA = 0x3000
EX_SLOTS = 65 * 7 + 64            # exchange slots of 16 bytes
XSLOTS = 546                      # 16-byte slots (re0, re1, im0, im1) of X / Y: bin k at k + (k >> 4)
XW = 16                           # bytes per slot


def xslot(k):
    k = np.asarray(k)
    return k + (k >> 4)


def tables(overlaps: int):
    i = np.arange(N)
    win = (0.5 * (1 - np.cos(2 * np.pi * i / N))).astype(F32)
    win_out = (win * F32(1.0 / (2 * N)) * F32(1.0 / overlaps)).astype(F32)
    tw = np.exp(-2j * np.pi * i / N).astype(C64)
    for q in range(4):
        tw[q * N // 4] = [1, -1j, -1, 1j][q]
    return win, win_out, tw


def delta_table(pf32: np.float32) -> np.ndarray:
    """round(p * pitchFactor) - p (pv:125, round half up of the exact f64 product), or INVALID
    when the shifted peak is beyond nb (pv:127 `break`: this and all later peaks are dropped)."""
    p = np.arange(NB + 1, dtype=np.float64)
    ps = np.floor(p * np.float64(pf32) + 0.5)
    d = (ps - p).astype(np.int64)
    d[ps > NB] = INVALID_DELTA
    return d


class Conflicts:
    """worst wavefront count per warp access, per pattern name"""

    def __init__(self):
        self.worst = {}

    def note(self, name, byte_addr, width):
        byte_addr = np.asarray(byte_addr).reshape(-1)
        assert byte_addr.size == 32
        group = {4: 32, 8: 16, 16: 8}[width]
        total = 0
        for g in range(0, 32, group):
            a = byte_addr[g:g + group]
            a = a[a >= 0]
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


W8 = np.exp(-2j * np.pi * np.outer(np.arange(8), np.arange(8)) / 8).astype(np.complex128)


def dft8(x, inv=False):
    """x[8, ...] complex -> DFT over axis 0 (forward e^-, inverse e^+, unnormalised)"""
    w = np.conj(W8) if inv else W8
    return np.tensordot(w, x.astype(np.complex128), axes=(1, 0)).astype(C64)


def lane_butterflies():
    """(kA, kB) natural base bins of each lane's two last-pass butterflies; lane 0 owns the
    self-paired ones (0 and 32)."""
    kA = LANES.copy()
    kB = 64 - LANES
    kB[0] = 32
    return kA, kB


def step(hist2, acc2, inblk, t, pf32, hop, cf: Conflicts | None = None):
    """One process() call of one channel pair.  hist2/acc2: [N][2] float32 rings (modified in
    place), inblk: [2][hop] or None (paused), t = timeCursor (multiple of hop).  Returns out[2][hop]."""
    assert hop % 128 == 0
    R = N // hop
    win, win_out, tw = tables(R)
    dtab = delta_table(pf32)
    contract = bool(pf32 < 1.0)
    t = int(t) % N
    jb = ((t - hop) // 128) & 7
    nblk = hop // 128
    ex = np.zeros((EX_SLOTS, 2), C64)              # exchange: slot -> (ch0, ch1) complex

    # ---- load, window, forward pass 1 (DFT over j, stride 64) ----------------------------------
    for h in range(2):
        nl = LANES + 32 * h
        z = np.zeros((8, 32, 2), C64)
        for j in range(8):
            i = 2 * nl + 128 * j                    # ring index of sample pair (i, i + 1)
            jj = (j - jb) & 7
            if jj < nblk:                           # new block (ola:91-108)
                s = 2 * nl + 128 * jj
                if inblk is None:
                    x0 = np.zeros((32, 2), F32); x1 = np.zeros((32, 2), F32)
                else:
                    x0 = inblk[:, s].T.astype(F32); x1 = inblk[:, s + 1].T.astype(F32)
                hist2[i] = x0; hist2[i + 1] = x1
            else:
                x0 = hist2[i]; x1 = hist2[i + 1]
            w0 = win[(i - t) % N][:, None]; w1 = win[(i + 1 - t) % N][:, None]
            z[j] = (x0 * w0).astype(F32) + 1j * (x1 * w1).astype(F32)
        z = dft8(z)
        for k1 in range(8):
            v = z[k1] * tw[(2 * nl * k1) % N][:, None]
            slot = 65 * k1 + nl
            ex[slot] = v
            if cf: cf.note("p1_st", slot * 16, 16)

    # ---- forward pass 2 (DFT over m2), in place -------------------------------------------------
    m3 = LANES & 7
    for h in range(2):
        k1 = (LANES >> 3) + 4 * h
        base = 65 * k1 + m3
        x = np.stack([ex[base + 8 * m2] for m2 in range(8)])
        if cf:
            for m2 in range(8): cf.note("p2_ld", (base + 8 * m2) * 16, 16)
        x = dft8(x)
        for k2 in range(8):
            ex[base + 8 * k2] = x[k2] * tw[(16 * m3 * k2) % N][:, None]

    # ---- forward pass 3 (DFT over m3): outputs stay in registers -------------------------------
    kA, kB = lane_butterflies()
    sA = 65 * (kA & 7) + 8 * (kA >> 3)
    sB = 65 * (kB & 7) + 8 * (kB >> 3)
    a = dft8(np.stack([ex[sA + c] for c in range(8)]))       # a[j] = Z[kA + 64 j]
    b = dft8(np.stack([ex[sB + c] for c in range(8)]))       # b[j] = Z[kB + 64 j]
    if cf:
        for c in range(8):
            cf.note("p3_ldA", (sA + c) * 16, 16); cf.note("p3_ldB", (sB + c) * 16, 16)

    # ---- real split (registers) -> X (2x scaled) in the padded per-channel layout --------------
    # general lanes: slot j pairs (a[j], b[7-j]) at k = lane + 64 j.
    # lane 0: j < 4: (b[j], b[7-j]) at k = 32 + 64 j; j >= 4: (a[j-4], a[(12-j) & 7]) at k = 64 (j-4);
    #         plus the self pair k = 256 (a[4]).
    l0 = (LANES == 0)[:, None]
    X = np.zeros((2, XSLOTS), C64)

    def split(za, zb, k, active=None):
        e = za + np.conj(zb)
        o = -1j * (za - np.conj(zb))
        tt = o * tw[k][:, None]
        xk = e + tt                                  # 2 X[k]
        xm = np.conj(e - tt)                         # 2 X[M - k]
        s1, s2 = xslot(k), xslot(M - k)
        for L in range(32):
            if active is not None and not active[L]: continue
            X[:, s1[L]] = xk[L]; X[:, s2[L]] = xm[L]
        if cf and active is None:
            cf.note("split_st_k", s1 * XW, 16); cf.note("split_st_mk", s2 * XW, 16)

    for j in range(8):
        if j < 4:
            za = np.where(l0, b[j], a[j]); zb = np.where(l0, b[7 - j], b[7 - j])
            k = np.where(LANES == 0, 32 + 64 * j, LANES + 64 * j)
        else:
            za = np.where(l0, a[j - 4], a[j]); zb = np.where(l0, a[(12 - j) & 7], b[7 - j])
            k = np.where(LANES == 0, 64 * (j - 4), LANES + 64 * j)
        split(za, zb, k)
    split(a[4], a[4], np.full(32, 256), active=(LANES == 0))

    # ---- per channel: peaks, owners, in-place shift ---------------------------------------------
    for ch in range(2):
        Xc = X[ch]
        b0 = 16 * LANES
        # run read: bins b0-2 .. b0+17 (halo for the 5-point stencil), as 16-byte loads of bin pairs
        run = np.stack([Xc[xslot(np.clip(b0 + e, 0, 513))] for e in range(-2, 18)])     # [20][32]
        if cf:
            for e in range(16): cf.note("run_ld", xslot(b0 + e) * XW, 16)
        mag = (run.real.astype(F32) ** 2 + run.imag.astype(F32) ** 2).astype(F32)
        bins = b0[None, :] + np.arange(-2, 18)[:, None]
        mag[(bins < 0) | (bins > M)] = F32(-1.0)      # outside the spectrum: never a neighbour that wins
        mask = np.zeros(32, np.int64)
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
        prev_before = np.full(32, -30000); next_after = np.full(32, 30000)
        for L in range(32):
            lo = [l for l in range(L) if nz[l]]
            hi = [l for l in range(L + 1, 32) if nz[l]]
            if lo: prev_before[L] = own_last[lo[-1]]
            if hi: next_after[L] = own_first[hi[0]]
        p_last = own_last[np.nonzero(nz)[0][-1]]
        d_last = int(dtab[p_last])

        # extension: bin 512 and the first stale level (bundle:394-438), owned by the last peak
        ext = np.zeros((4, 32), C64)
        for i in range(4):
            q = LANES + 32 * i
            if i == 0 or contract:
                qq = np.maximum(q, 1)
                A = Xc[xslot(qq)]; B = Xc[xslot(256 + qq)]; Cc = Xc[xslot(512 - qq)]; D = Xc[xslot(256 - qq)]
                s = (A - B) + np.conj(Cc - D)
                v = (0.25 * s * np.conj(tw[(2 * qq) % N])).astype(C64)
                v = np.where(q == 0, Xc[xslot(M)], v)
                if not contract: v = np.where(q == 0, v, 0)
                ext[i] = v
                if cf and contract:
                    cf.note("stale_ld", xslot(qq) * XW, 16); cf.note("stale_ld_m", xslot(512 - qq) * XW, 16)

        # owner of every bin of the run: nearest peak, ties to the higher one (pv:132-141)
        nextv = np.zeros((16, 32), np.int64)
        Q = next_after.copy()
        for e in range(15, -1, -1):
            nextv[e] = Q
            Q = np.where((mask >> e) & 1, b0 + e, Q)
        P = prev_before.copy()
        dest = np.zeros((16, 32), np.int64); first = np.zeros((16, 32), bool)
        for e in range(16):
            bb = b0 + e
            P = np.where((mask >> e) & 1, bb, P)
            take_next = (nextv[e] - bb) <= (bb - P)
            owner = np.where(take_next, nextv[e], P)
            dest[e] = bb + dtab[owner]
            first[e] = ~take_next | (not contract)    # right half of its region, or expanding

        # in place: every lane holds its sources in registers; zero fill, then two ordered sub-steps
        # (right halves are pairwise disjoint after the shift, and so are left halves: checked here)
        X[ch, :] = 0
        written = np.zeros(XSLOTS, np.int64)
        for e in range(16):                           # first sub-step: plain stores
            ok = (dest[e] >= 0) & (dest[e] < NB) & first[e]
            for L in np.nonzero(ok)[0]:
                written[xslot(dest[e][L])] += 1
                X[ch, xslot(dest[e][L])] = xv[e][L]
        for i in range(4):
            d = 512 + LANES + 32 * i + d_last
            ok = (d >= 0) & (d < NB)
            for L in np.nonzero(ok)[0]:
                written[xslot(d[L])] += 1
                X[ch, xslot(d[L])] = ext[i][L]
        assert written.max() <= 1, "two first writers for one bin"
        written[:] = 0
        if contract:
            for e in range(16):                       # second sub-step: left halves add on top
                ok = (dest[e] >= 0) & (dest[e] < NB) & ~first[e]
                for L in np.nonzero(ok)[0]:
                    written[xslot(dest[e][L])] += 1
                    X[ch, xslot(dest[e][L])] += xv[e][L]
            assert written.max() <= 1, "two second writers for one bin"

    # ---- Hermitian C2R pre-pass (mirror of the split) -------------------------------------------
    def unsplit(k):
        yk = X[:, xslot(k)].T.copy(); ym = X[:, xslot(M - k)].T.copy()       # [32][2]
        z0 = (k == 0)[:, None]
        yk = np.where(z0, yk.real + 0j, yk); ym = np.where(z0, ym.real + 0j, ym)
        e = yk + np.conj(ym)
        d = yk - np.conj(ym)
        pp = d * np.conj(tw[k])[:, None]
        zk = e + 1j * pp
        zmk = np.conj(e - 1j * pp)
        return zk.astype(C64), zmk.astype(C64)

    a = np.zeros((8, 32, 2), C64); b = np.zeros((8, 32, 2), C64)
    for j in range(8):
        if j < 4: k = np.where(LANES == 0, 32 + 64 * j, LANES + 64 * j)
        else: k = np.where(LANES == 0, 64 * (j - 4), LANES + 64 * j)
        zk, zmk = unsplit(k)
        if j < 4:
            a[j] = np.where(l0, a[j], zk); b[7 - j] = np.where(l0, zmk, zmk)
            b[j] = np.where(l0, zk, b[j])
        else:
            a[j] = np.where(l0, a[j], zk)
            b[7 - j] = np.where(l0, b[7 - j], zmk)
            # lane 0: (a[j-4], a[(12-j)&7])
            a[j - 4] = np.where(l0, zk, a[j - 4])
            if j > 4: a[(12 - j) & 7] = np.where(l0, zmk, a[(12 - j) & 7])
    z256, _ = unsplit(np.full(32, 256))
    a[4] = np.where(l0, z256, a[4])

    # ---- inverse pass 1 (DFT over k3 -> m3), twiddle conj(W_64^{k2 m3}) ------------------------
    a = dft8(a, inv=True); b = dft8(b, inv=True)
    for c in range(8):
        ex[sA + c] = a[c] * np.conj(tw[(16 * (kA >> 3) * c) % N])[:, None]
        ex[sB + c] = b[c] * np.conj(tw[(16 * (kB >> 3) * c) % N])[:, None]
    # ---- inverse pass 2 (DFT over k2 -> m2), twiddle conj(W_512^{k1 (m3 + 8 m2)}) ---------------
    for h in range(2):
        k1 = (LANES >> 3) + 4 * h
        base = 65 * k1 + m3
        x = dft8(np.stack([ex[base + 8 * k2] for k2 in range(8)]), inv=True)
        for m2 in range(8):
            ex[base + 8 * m2] = x[m2] * np.conj(tw[(2 * k1 * (m3 + 8 * m2)) % N])[:, None]
    # ---- inverse pass 3 (DFT over k1 -> j); window, overlap-add, emit ---------------------------
    out = np.zeros((2, hop), F32)
    je = (t // 128) & 7                                 # ring block of frame sample 0 (emitted)
    for h in range(2):
        nl = LANES + 32 * h
        x = dft8(np.stack([ex[65 * k1 + nl] for k1 in range(8)]), inv=True)
        for j in range(8):
            i = 2 * nl + 128 * j
            w0 = win_out[(i - t) % N][:, None]; w1 = win_out[(i + 1 - t) % N][:, None]
            y0 = (x[j].real.astype(F32) * w0).astype(F32); y1 = (x[j].imag.astype(F32) * w1).astype(F32)
            tail = ((j - jb) & 7) < nblk                # starts from zero (ola:134)
            q0 = np.zeros((32, 2), F32) if tail else acc2[i]
            q1 = np.zeros((32, 2), F32) if tail else acc2[i + 1]
            y0 = y0 + q0; y1 = y1 + q1
            jj = (j - je) & 7
            if jj < nblk:                               # head: emit (ola:111-118)
                s = 2 * nl + 128 * jj
                out[:, s] = y0.T; out[:, s + 1] = y1.T
            else:
                acc2[i] = y0; acc2[i + 1] = y1
    return out


def run(signal: np.ndarray, pf: float, hop: int, cf: Conflicts | None = None, start_calls: int = 0):
    """signal [2][T*hop] -> output [2][T*hop]"""
    pf32 = np.float32(pf)
    hist2 = np.zeros((N, 2), F32); acc2 = np.zeros((N, 2), F32)
    T = signal.shape[1] // hop
    out = np.zeros_like(signal, dtype=F32)
    for m in range(T):
        out[:, m * hop:(m + 1) * hop] = step(hist2, acc2, signal[:, m * hop:(m + 1) * hop],
                                             (start_calls + m) * hop, pf32, hop, cf)
    return out

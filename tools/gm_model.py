"""Plain-numpy model of the in-place iterative Gao-Mateer additive FFT used by the CUDA
kernels (development aid and documentation; not imported by the product).

Layout contract shared with bch_fft_b200/csrc:
  * forward phase, depth d = 0..K-1: subproblem (d, s) keeps coefficient i at position
    i*2^d + s; twist multiplies position p by tw[d][p >> d]; Taylor levels b = K-2-d..0
    act on position bits (d+b+1, d+b).
  * backward phase, depth d = K-1..0: butterfly on position bit d with constant
    gam[d][p >> (d+1)]:  lo' = lo + G*hi ; hi' = lo' + hi.
  * after the last butterfly position p holds P(bitrev_m(p)) when K = m; for K < m the
    2^K values are first replicated over the 2^(m-K) cosets (positions p + c*2^K).
"""
import numpy as np


class Field:
    POLY = {2: 0x7, 3: 0xB, 4: 0x13, 5: 0x25, 6: 0x43, 7: 0x83, 8: 0x171, 9: 0x211, 10: 0x409,
            11: 0x805, 12: 0x1099, 13: 0x201B, 14: 0x5803, 15: 0x8003, 16: 0x1002D,
            17: 0x20009, 18: 0x40081, 19: 0x80063, 20: 0x100009}

    def __init__(self, m):
        self.m, self.n = m, 1 << m
        self.order = self.n - 1
        exp = np.zeros(2 * self.n, np.int64)
        log = np.zeros(self.n, np.int64)
        v = 1
        for i in range(self.order):
            exp[i] = v
            log[v] = i
            v <<= 1
            if v & self.n:
                v ^= self.POLY[m]
        exp[self.order:2 * self.order] = exp[:self.order]
        self.exp, self.log = exp, log

    def mul(self, a, b):
        a = np.asarray(a, np.int64)
        b = np.asarray(b, np.int64)
        r = self.exp[self.log[a] + self.log[b]]
        return np.where((a == 0) | (b == 0), 0, r)

    def inv(self, a):
        return int(self.exp[(self.order - self.log[a]) % self.order])

    def pw(self, a, e):
        if a == 0:
            return 0 if e else 1
        return int(self.exp[(int(self.log[a]) * e) % self.order])


def bitrev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def gm_constants(F):
    """tw[d][i] (i < 2^(m-d)), gam[d][q] (q < 2^(m-d-1), indexed by position bits above d)."""
    m = F.m
    basis = [1 << i for i in range(m)]          # beta_1..beta_m = 1, 2, 4, ...
    tw, gam = [], []
    for d in range(m):
        k = m - d
        tau = basis[-1]
        tw.append(np.array([F.pw(tau, i) for i in range(1 << k)], np.int64))
        it = F.inv(tau)
        g = [int(F.mul(b, it)) for b in basis[:-1]]      # gamma_1..gamma_{k-1}
        # twiddle for child output index j = XOR_b bit_b(j) gamma_{b+1}; position bits above d
        # hold j bit-reversed over k-1 bits
        G = np.zeros(1 << (k - 1), np.int64)
        for q in range(1 << (k - 1)):
            j = bitrev(q, k - 1)
            v = 0
            for b in range(k - 1):
                if (j >> b) & 1:
                    v ^= g[b]
            G[q] = v
        gam.append(G)
        basis = [int(F.mul(x, x)) ^ x for x in g]
    return tw, gam


def gm_fft(F, coef, K=None, consts=None):
    """coef: 2^K coefficients (K <= m).  Returns a[p] = P(bitrev_m(p)) over all 2^m positions."""
    m = F.m
    K = m if K is None else K
    tw, gam = consts or gm_constants(F)
    a = np.array(coef, np.int64)
    assert a.size == 1 << K
    pos = np.arange(1 << K)
    for d in range(K):
        a = F.mul(a, tw[d][pos >> d])
        for b in range(K - d - 2, -1, -1):
            lo, hi = d + b, d + b + 1
            q = 1 << lo
            sel = (((pos >> hi) & 1) == 1) & (((pos >> lo) & 1) == 0)
            a[pos[sel]] ^= a[pos[sel] + q]
            sel = (((pos >> hi) & 1) == 0) & (((pos >> lo) & 1) == 1)
            a[pos[sel]] ^= a[pos[sel] + q]
    a = np.tile(a, 1 << (m - K))                  # constants broadcast over the cosets
    pos = np.arange(1 << m)
    for d in range(K - 1, -1, -1):
        lo = pos[((pos >> d) & 1) == 0]
        G = gam[d][lo >> (d + 1)]
        a[lo] ^= F.mul(G, a[lo + (1 << d)])
        a[lo + (1 << d)] ^= a[lo]
    return a


def taylor_forward_xor(a, m):
    """forward phase of a fully twist-free field (XOR only): m(m-1)/2 Taylor levels"""
    a = a.copy()
    pos = np.arange(1 << m)
    for d in range(m):
        for lo in range(m - 2, d - 1, -1):
            q = 1 << lo
            sel = (((pos >> (lo + 1)) & 1) == 1) & (((pos >> lo) & 1) == 0)
            a[pos[sel]] ^= a[pos[sel] + q]
            sel = (((pos >> (lo + 1)) & 1) == 0) & (((pos >> lo) & 1) == 1)
            a[pos[sel]] ^= a[pos[sel] + q]
    return a


def tx_level(a, m, lo, h):
    """OP_TX(lo, h) of gm_plan.h: generalised Taylor level at y^(2^h) + y, blocks of 2^lo"""
    t = 1 << h
    pos = np.arange(1 << m)
    blk = (pos >> lo) & (2 * t - 1)
    off = (t - 1) << lo
    sel = blk == t
    a[pos[sel]] ^= a[pos[sel] + off]
    sel = (blk >= 1) & (blk < t)
    a[pos[sel]] ^= a[pos[sel] + off]


def cantor_conversion(a, m, b0=0, w=None):
    """same coefficients as taylor_forward_xor in (m/2) log2(m) levels (m a power of two)"""
    w = m if w is None else w
    if w < 2:
        return
    h = w // 2
    for k in range(h - 1, -1, -1):
        tx_level(a, m, b0 + k, h)
    cantor_conversion(a, m, b0, h)
    cantor_conversion(a, m, b0 + h, h)


def natural(F, a):
    out = np.zeros_like(a)
    for p in range(F.n):
        out[bitrev(p, F.m)] = a[p]
    return out


if __name__ == "__main__":
    import sys
    sys.path.insert(0, ".")
    from oracle.pyoracle import Port
    P = Port()
    rng = np.random.default_rng(3)
    for m in (2, 4, 8, 16):
        a = rng.integers(0, 1 << 30, 1 << m)
        b = a.copy()
        cantor_conversion(b, m)
        assert (taylor_forward_xor(a, m) == b).all(), m
    print("cantor conversion == taylor forward (m = 2, 4, 8, 16)")
    for m in (3, 4, 5, 8, 10, 13):
        F = Field(m)
        c = gm_constants(F)
        for K in sorted({1, 2, 3, m // 2, m - 1, m}):
            if K < 1 or K > m:
                continue
            coef = rng.integers(0, F.n, 1 << K)
            full = np.zeros(F.n, np.uint32)
            full[:1 << K] = coef
            _, nat = P.additive_fft(m, full, F.n - 1)
            got = natural(F, gm_fft(F, coef, K, c))
            assert (got == nat).all(), (m, K)
        print("m", m, "ok")


def taylor_group_xor(a, K, d0, depths):
    """Taylor levels of the depths d0 .. d0 + depths - 1 (no twist between them) on 2^K coefficients"""
    a = a.copy()
    pos = np.arange(1 << K)
    for d in range(d0, d0 + depths):
        for lo in range(K - 2, d - 1, -1):
            q = 1 << lo
            sel = (((pos >> (lo + 1)) & 1) == 1) & (((pos >> lo) & 1) == 0)
            a[pos[sel]] ^= a[pos[sel] + q]
            sel = (((pos >> (lo + 1)) & 1) == 0) & (((pos >> lo) & 1) == 1)
            a[pos[sel]] ^= a[pos[sel] + q]
    return a


def grouped_conversion(a, K, d0, h):
    """what gm_plan.h emits for such a group (h a power of two): K - d0 - h levels of base x^(2^h) + x on the
    index bits d0 .. K-1, then the Cantor conversion of bits d0 .. d0 + h - 1"""
    a = a.copy()
    for k in range(K - h - 1, d0 - 1, -1):
        tx_level(a, K, k, h)
    cantor_conversion(a, K, d0, h)
    return a


def check_grouped_conversion():
    rng = np.random.default_rng(5)
    for K, d0, h in ((6, 0, 2), (9, 0, 4), (13, 0, 4), (14, 0, 8), (12, 4, 4), (12, 8, 4), (13, 4, 4),
                     (14, 2, 2), (14, 4, 2), (10, 6, 4), (8, 4, 4)):
        a = rng.integers(0, 1 << 30, 1 << K)
        assert (taylor_group_xor(a, K, d0, h) == grouped_conversion(a, K, d0, h)).all(), (K, d0, h)
    print("grouped Cantor conversion == Taylor levels of the group (every offset)")


if __name__ == "__main__":
    check_grouped_conversion()

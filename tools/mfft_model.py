"""Model of the multiplicative FFT over GF(2^16) used by k_mfft16_* (numpy, exact):
Good-Thomas 65535 = 255 x 257 (no twiddles), the 257-point transforms by the cyclotomic method
(linearised polynomials over the cyclotomic cosets of 2 mod 257, coordinates in a normal basis), the
255-point transforms as a 3 x 5 x 17 prime-factor transform.  Checked against the oracle's
multiplicative_FFT (oracle/bchfft_oracle.c).   python tools/mfft_model.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402

M, N, N1, N2 = 16, 65535, 255, 257
P = pyoracle.Port()
EXP, LOG = P.tables(M)
EXP = EXP.astype(np.int64)
LOG = LOG.astype(np.int64)


def mul(a, b):
    a = np.asarray(a, np.int64)
    b = np.asarray(b, np.int64)
    r = EXP[(LOG[a] + LOG[b]) % N]
    return np.where((a == 0) | (b == 0), 0, r)


def normal_basis():
    """smallest e such that the 16 conjugates of x^e are linearly independent over GF(2)"""
    for e in range(1, N):
        g = int(EXP[e])
        conj = [g]
        for _ in range(15):
            conj.append(int(mul(conj[-1], conj[-1])))
        rows = conj[:]
        rank = 0
        for bit in range(15, -1, -1):
            piv = next((i for i in range(rank, 16) if (rows[i] >> bit) & 1), None)
            if piv is None:
                continue
            rows[rank], rows[piv] = rows[piv], rows[rank]
            for i in range(16):
                if i != rank and (rows[i] >> bit) & 1:
                    rows[i] ^= rows[rank]
            rank += 1
        if rank == 16:
            return e, conj
    raise RuntimeError


def coords(conj, z):
    """bits c_b with z = sum_b c_b conj[b]"""
    # solve by brute-force Gaussian elimination on the augmented system
    rows = [(conj[b], 1 << b) for b in range(16)]
    # reduce to a basis in echelon form keeping track of combinations
    ech = {}
    for v, comb in rows:
        for bit in range(15, -1, -1):
            if not (v >> bit) & 1:
                continue
            if bit in ech:
                v ^= ech[bit][0]
                comb ^= ech[bit][1]
            else:
                ech[bit] = (v, comb)
                break
    out = 0
    for bit in range(15, -1, -1):
        if (z >> bit) & 1:
            z ^= ech[bit][0]
            out ^= ech[bit][1]
    assert z == 0
    return out


def dft257(a, conj, ctab, cosets):
    """a[257] -> A(w2^j), j < 257, by the cyclotomic method"""
    w = len(cosets)
    V = np.zeros((w, 16), np.int64)
    for s, ks in enumerate(cosets):
        for b in range(16):
            acc = 0
            for i in range(16):
                acc ^= int(mul(a[(ks << i) % N2], conj[(b + i) % 16]))
            V[s, b] = acc
    out = np.zeros(N2, np.int64)
    for j in range(N2):
        acc = int(a[0])
        for s, ks in enumerate(cosets):
            cw = ctab[(j * ks) % N2]
            for b in range(16):
                if (cw >> b) & 1:
                    acc ^= int(V[s, b])
        out[j] = acc
    return out


def dft_direct(a, root_log, p):
    """p-point DFT with the root x^root_log"""
    out = np.zeros(p, np.int64)
    for j in range(p):
        acc = 0
        for k in range(p):
            acc ^= int(mul(a[k], EXP[(root_log * j * k) % N]))
        out[j] = acc
    return out


def main():
    rng = np.random.default_rng(1)
    a = rng.integers(0, 1 << M, N).astype(np.int64)
    want = P.mult_fft(M, a.astype(np.uint32))
    e, conj = normal_basis()
    print("normal basis generator x^%d = 0x%x" % (e, conj[0]))
    w2 = 255                                   # log of the 257th root of unity x^255
    ctab = [coords(conj, int(EXP[(w2 * k) % N])) for k in range(N2)]
    seen, cosets = {0}, []
    for k in range(1, N2):
        if k in seen:
            continue
        cosets.append(k)
        for i in range(16):
            seen.add((k << i) % N2)
    assert len(cosets) == 16
    # pass 1: 257-point transforms for every k1
    T = np.zeros((N1, N2), np.int64)
    for k1 in range(N1):
        src = [(N2 * k1 + N1 * k2) % N for k2 in range(N2)]
        T[k1] = dft257(a[src], conj, ctab, cosets) if k1 < 3 else 0
    # check pass 1 on the first rows against the direct transform
    for k1 in range(3):
        src = [(N2 * k1 + N1 * k2) % N for k2 in range(N2)]
        assert (T[k1] == dft_direct(a[src], w2, N2)).all()
    print("cyclotomic 257-point transform ok")
    # full transform with direct small DFTs (checks the index maps): vectorised over rows
    L = LOG
    def dft_axis(X, root_log, p, axis):
        X = np.moveaxis(X, axis, 0)
        out = np.zeros_like(X)
        for j in range(p):
            acc = np.zeros(X.shape[1:], np.int64)
            for k in range(p):
                acc ^= mul(X[k], EXP[(root_log * j * k) % N])
            out[j] = acc
        return np.moveaxis(out, 0, axis)
    k1 = np.arange(N1)[:, None]
    k2 = np.arange(N2)[None, :]
    A2 = a[(N2 * k1 + N1 * k2) % N]
    T = dft_axis(A2, 255, N2, 1)                       # T[k1][i2]
    # 255 = 3 x 5 x 17 prime-factor transform along k1: k1 = (85 a + 51 b + 15 c) % 255
    ia, ib, ic = np.meshgrid(np.arange(3), np.arange(5), np.arange(17), indexing="ij")
    kk = (85 * ia + 51 * ib + 15 * ic) % 255
    X = T[kk]                                          # [3][5][17][257]
    w1 = 257                                           # log of the 255th root of unity
    X = dft_axis(X, w1 * 85, 3, 0)
    X = dft_axis(X, w1 * 51, 5, 1)
    X = dft_axis(X, w1 * 15, 17, 2)
    inv17 = pow(15, -1, 17)
    i1 = (85 * ia + 51 * ib + 15 * inv17 * ic) % 255   # output index of (a', b', c')
    c1 = N2 * pow(N2, -1, N1)
    c2 = N1 * pow(N1, -1, N2)
    out = np.zeros(N, np.int64)
    i2 = np.arange(N2)
    for aa in range(3):
        for bb in range(5):
            for cc in range(17):
                out[(i1[aa, bb, cc] * c1 + i2 * c2) % N] = X[aa, bb, cc]
    assert (out == want).all()
    print("Good-Thomas 255 x 257 with 3 x 5 x 17 prime-factor transform == oracle multiplicative_FFT")
    print("c1 = %d, c2 = %d, inv17 = %d" % (c1 % N, c2 % N, inv17))


if __name__ == "__main__":
    main()

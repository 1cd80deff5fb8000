"""Prototype of the matrix-representation engine's planner (numpy; the C++ planner in csrc/device/dense_matrix.cu
follows it).  G(p,q) -> real matrices through n pairwise anticommuting signed Pauli strings
P(x,z) = X^x Z^z over m' qubits:  P(x,z)[i,j] = [i == j^x] (-1)^(z.j),  P(x,z)^2 = (-1)^(x.z),
P(x,z) P(x',z') = (-1)^(z.x') P(x^x', z^z')."""
import itertools, sys
import numpy as np

def par(v): return bin(v).count("1") & 1

def find_strings(signs, mp):
    """n vectors (x,z) in F2^mp x F2^mp, pairwise anticommuting, Q = x.z = [sign<0], linearly independent."""
    n = len(signs)
    K = 1 << mp
    cand = [[(x, z) for x in range(K) for z in range(K) if par(x & z) == (1 if s < 0 else 0) and (x or z)] for s in signs]
    chosen = []
    def indep(vs):
        basis = []
        for x, z in vs:
            v = x << mp | z
            for b in basis: v = min(v, v ^ b)
            if v == 0: return False
            basis.append(v)
        return True
    def rec(i):
        if i == n: return True
        for (x, z) in cand[i]:
            if all(par(x & z2) ^ par(z & x2) for x2, z2 in chosen):
                chosen.append((x, z))
                if indep(chosen) and rec(i + 1): return True
                chosen.pop()
        return False
    return list(chosen) if rec(0) else None

def gf2_rank_basis(vs, width):
    basis = []
    for v in vs:
        for b in basis: v = min(v, v ^ b)
        if v: basis.append(v)
    return basis

def matvec(M, v, rows):
    # M: list of row bitmasks; returns bits r = parity(M[r] & v)
    return sum(par(M[r] & v) << r for r in range(rows))

def gf2_inv(M, m):
    A = [M[r] | (1 << (m + r)) for r in range(m)]
    for c in range(m):
        p = next(r for r in range(c, m) if A[r] >> c & 1)
        A[c], A[p] = A[p], A[c]
        for r in range(m):
            if r != c and A[r] >> c & 1: A[r] ^= A[c]
    return [a >> m for a in A]

def transpose(M, m):
    return [sum(((M[r] >> c) & 1) << r for r in range(m)) for c in range(m)]

def plan(signs):
    n = len(signs)
    for mp in range((n + 1) // 2, (n + 1) // 2 + 3):
        vs = find_strings(signs, mp)
        if vs is None: continue
        # subgroup G spanned by the strings; Z0 = {z : (0,z) in G}; need x-projection onto
        # enumerate G
        K = 1 << mp
        blade_str = {}
        for S in range(1 << n):
            x = z = 0; sg = 0
            for i in range(n):
                if S >> i & 1:
                    # (X^x Z^z)(X^xi Z^zi) = (-1)^(z.xi) X^(x^xi) Z^(z^zi)
                    sg ^= par(z & vs[i][0]); x ^= vs[i][0]; z ^= vs[i][1]
            blade_str[S] = (x, z, sg)
        xs = {v[0] for v in blade_str.values()}
        if len(xs) != K:
            continue
        Z0 = sorted({v[1] for v in blade_str.values() if v[0] == 0})
        d0 = n - mp
        assert len(Z0) == 1 << d0
        zb = gf2_rank_basis(Z0, mp)
        assert len(zb) == d0
        # M invertible with M zb[k] = e_k: complete zb to a basis, M = inverse of the matrix with those columns
        full = list(zb)
        for e in range(mp):
            if len(gf2_rank_basis(full + [1 << e], mp)) > len(full): full.append(1 << e)
        cols = full  # column c = full[c]
        Cm = [sum(((cols[c] >> r) & 1) << c for c in range(mp)) for r in range(mp)]  # rows
        M = gf2_inv(Cm, mp)
        MinvT = transpose(Cm, mp)  # (M^-1)^T = Cm^T
        vs2 = [(matvec(MinvT, x, mp), matvec(M, z, mp)) for x, z in vs]
        for (x, z), s in zip(vs2, signs): assert par(x & z) == (1 if s < 0 else 0)
        for a, b in itertools.combinations(vs2, 2): assert par(a[0] & b[1]) ^ par(a[1] & b[0]) == 1
        return mp, d0, vs2
    return None

def tables(signs):
    n = len(signs)
    mp, d0, vs = plan(signs)
    K, NJ = 1 << mp, 1 << d0
    # blade -> (x, z, sign)
    info = {}
    for S in range(1 << n):
        x = z = 0; sg = 0
        for i in range(n):
            if S >> i & 1:
                sg ^= par(z & vs[i][0]); x ^= vs[i][0]; z ^= vs[i][1]
        info[S] = (x, z, sg)
    # section L: x -> z with low d0 bits zero
    Lx = {}
    src = {}
    for S, (x, z, sg) in info.items():
        hi = z >> d0 << d0
        if x in Lx: assert Lx[x] == hi, "Z0 is not the low subspace"
        Lx[x] = hi
        src[(x, z & (NJ - 1))] = (S, sg)
    assert len(src) == 1 << n
    return n, mp, d0, vs, Lx, src

def blade_product_tables(signs):
    n = len(signs)
    def coeff(a, b):
        s = 1.0
        # canonical reordering sign
        t = a >> 1; c = 0
        while t: c += bin(t & b).count("1"); t >>= 1
        if c & 1: s = -s
        for i in range(n):
            if (a & b) >> i & 1: s *= signs[i]
        return s
    return coeff

def check(signs, seed=0):
    n, mp, d0, vs, Lx, src = tables(signs)
    K, NJ = 1 << mp, 1 << d0
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, 1 << n); b = rng.uniform(-1, 1, 1 << n)
    # direct product in the blade basis
    coeff = blade_product_tables(signs)
    want = np.zeros(1 << n)
    if n <= 10:
        for s in range(1 << n):
            for t in range(1 << n):
                want[s ^ t] += coeff(s, t) * a[s] * b[t]
    had = np.array([[(-1) ** par(t & j) for t in range(NJ)] for j in range(NJ)], dtype=float)
    def W(v):
        w = np.zeros((K, NJ))
        for x in range(K):
            col = np.array([(-1) ** src[(x, t)][1] * v[src[(x, t)][0]] for t in range(NJ)])
            w[x] = had @ col
        return w
    WA, WB = W(a), W(b)
    MA = np.zeros((K, K)); MB = np.zeros((K, NJ))
    for i in range(K):
        for l in range(K):
            x = i ^ l
            MA[i, l] = (-1) ** par(Lx[x] & l) * WA[x, l & (NJ - 1)]
    for l in range(K):
        for j in range(NJ):
            MB[l, j] = WB[l ^ j, j]
    C = MA @ MB
    WC = np.zeros((K, NJ))
    for i in range(K):
        for j in range(NJ):
            WC[i ^ j, j] = C[i, j]
    got = np.zeros(1 << n)
    for x in range(K):
        col = had @ WC[x] / NJ
        for t in range(NJ):
            S, sg = src[(x, t)]
            got[S] = (-1) ** sg * col[t]
    err = np.abs(got - want).max() if n <= 10 else float("nan")
    print(f"signs p={sum(s>0 for s in signs)} q={sum(s<0 for s in signs)}: m'={mp} d0={d0} K'={K} NJ={NJ} "
          f"fma={K*K*NJ} direct={4**n} ratio={4**n/(K*K*NJ):.0f} err={err:.2e} Lx nonzero={any(Lx.values())}")
    return err

if __name__ == "__main__":
    for p, q in [(2,0),(3,0),(4,0),(6,0),(7,0),(8,0),(4,1),(3,3),(4,4),(5,3),(9,0),(5,4),(0,6),(1,7)]:
        check([1.0] * p + [-1.0] * q)

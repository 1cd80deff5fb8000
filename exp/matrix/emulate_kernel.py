"""Lane-level emulation of dense_matrix_kernel.h (index logic only) against the host mirror of the library."""
import ctypes as C, numpy as np, sys
sys.path.insert(0, "/root/repo")
from gaast_b200 import _lib as L
import exp.matrix.proto as P

def slot_ab(NT, x, jc): return ((((x >> 2) * (NT // 4)) + (jc >> 2)) << 4) | ((x & 3) << 2) | (jc & 3)
def slot_c(NT, x, jc): return ((((x >> 2) * (NT // 4)) + (((jc >> 3) << 1) | (jc & 1))) << 4) | ((x & 3) << 2) | ((jc >> 1) & 3)
def par(v): return bin(v).count("1") & 1

def run(p, q, RC=1):
    n = p + q; neg = ((1 << q) - 1) << p
    shape = (C.c_int32 * 4)()
    NB = 1 << n
    rng = np.random.default_rng(1)
    a = rng.uniform(-1, 1, NB); b = rng.uniform(-1, 1, NB); c = np.zeros(NB)
    dp = lambda v: v.ctypes.data_as(C.POINTER(C.c_double))
    assert L.lib.gaast_diag_matrix_rep(n, neg, shape, dp(a), dp(b), dp(c)) == 0
    MX, DB, DL, has_lx = list(shape)
    D0 = DB + DL; NT = 1 << D0; K = 1 << MX; NJB = 1 << DL; NTI = max(1, NJB // 8); RI = K // RC; MTI = RI // 8
    # recover entry/lx tables by probing the mirror with unit vectors?  simpler: rebuild through unit products
    # -> instead check the slot maps are bijections and emulate the matmul fragment logic on W arrays from random data
    assert sorted(slot_ab(NT, x, jc) for x in range(K) for jc in range(NT)) == list(range(K * NT))
    assert sorted(slot_c(NT, x, jc) for x in range(K) for jc in range(NT)) == list(range(K * NT))
    WAv = rng.uniform(-1, 1, (K, NT)); WBv = rng.uniform(-1, 1, (K, NT)); lx = [0] * K
    WA = np.zeros(K * NT); WB = np.zeros(K * NT)
    for x in range(K):
        for jc in range(NT):
            WA[slot_ab(NT, x, jc)] = WAv[x, jc]; WB[slot_ab(NT, x, jc)] = WBv[x, jc]
    WC = np.full(K * NT, np.nan)
    for blk in range(1 << DB):
        for rc in range(RC):
            rowbase = rc * RI
            acc = np.zeros((MTI, NTI, 8, 8))
            for kt in range(K // 4):
                for mt in range(MTI):
                    for nt in range(NTI):
                        Af = np.zeros((8, 4)); Bf = np.zeros((4, 8))
                        for lane in range(32):
                            fr, fc = lane >> 2, lane & 3
                            l = 4 * kt + fc
                            j = 8 * nt + fr
                            Bf[fc, fr] = WB[slot_ab(NT, l ^ j, j | (blk << DL))] if (NJB >= 8 or j < NJB) else 0.0
                            xa = (rowbase + 8 * mt + fr) ^ l
                            Af[fr, fc] = WA[slot_ab(NT, xa, (l & (NJB - 1)) | (blk << DL))]
                        acc[mt, nt] += Af @ Bf
            for mt in range(MTI):
                for nt in range(NTI):
                    for lane in range(32):
                        fr, fc = lane >> 2, lane & 3
                        for u in range(2):
                            i = rowbase + 8 * mt + fr; j = 8 * nt + 2 * fc + u
                            if NJB >= 8 or j < NJB:
                                WC[slot_c(NT, i ^ j, j | (blk << DL))] = acc[mt, nt, fr, 2 * fc + u]
    assert not np.isnan(WC).any()
    # reference
    for blk in range(1 << DB):
        for i in range(K):
            for j in range(NJB):
                s = sum(WAv[i ^ l, (l & (NJB - 1)) | blk << DL] * WBv[l ^ j, j | blk << DL] for l in range(K))
                got = WC[slot_c(NT, i ^ j, j | blk << DL)]
                assert abs(s - got) < 1e-12, (i, j, s, got)
    # bank conflicts of the fragment reads (pitch odd: bank pair = slot mod 16 within a half warp)
    worst = 0
    for kt in range(K // 4):
        for half in range(2):
            for kind in range(3):
                banks = []
                for lane in range(16 * half, 16 * half + 16):
                    fr, fc = lane >> 2, lane & 3
                    l = 4 * kt + fc
                    if kind == 0: s = slot_ab(NT, fr ^ l, (l & (NJB - 1)))
                    elif kind == 1:
                        j = fr
                        if not (NJB >= 8 or j < NJB): continue
                        s = slot_ab(NT, l ^ j, j)
                    else: s = slot_c(NT, fr ^ (2 * fc), 2 * fc)
                    banks.append(s % 16)
                worst = max(worst, max(banks.count(v) for v in set(banks)))
    print(f"G({p},{q}) MX={MX} DB={DB} DL={DL} RC={RC}: fragment logic ok, worst bank multiplicity {worst}")

for (p, q, rc) in [(8,0,1),(7,0,1),(9,0,1),(10,0,2),(6,1,1),(11,0,4)]:
    run(p, q, rc)

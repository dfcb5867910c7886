"""Generate the per-lane term tables of the lane-parallel backward Riccati stage of the
kinematic kernels (csrc/kin_riccati_tables.cuh) and check them with a 32-lane simulator against
the plain formulas.

One entry of the value function V(x,w) = 1/2 [x;w]'[P W;W' Q][x;w] + [px;pw]'[x;w] per lane
(27 lanes).  A stage is three rounds of "gather a value from another lane, multiply by one or two
stage scalars, accumulate" (tables below) and a 2x2 pivot (hand-written in the kernel).
Run:  python scripts/gen_kin_riccati_tables.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# stage-record field offsets (doubles) - must match KinLayout in mpcb_kernel.cuh
NF = 45
CDEF, JAC, HXX, HUX, GX, HUU, EE, GU, TK = 0, 4, 10, 16, 17, 21, 23, 25, 27
ONE, TEE, ZERO = 41, 42, 43
A02, A03, A12, A13, A23, B2 = JAC + 0, JAC + 1, JAC + 2, JAC + 3, JAC + 4, JAC + 5
NEXT = NF  # offset of the next stage's record: its CDEF slots hold b = -(defect into stage k+1)

# ---- lane layout of the value function / G entries
PIDX = {}
q = 0
for i in range(4):
    for j in range(i, 4):
        PIDX[(i, j)] = q
        PIDX[(j, i)] = q
        q += 1
LP = lambda i, j: PIDX[(i, j)]            # P(i,j) / G_xx(i,j)
LW = lambda i, c: 10 + 2 * i + c          # W(i,c) / G_ux(c,i)
LQ = {(0, 0): 18, (0, 1): 19, (1, 0): 19, (1, 1): 20}
LPX = lambda i: 21 + i
LPW = lambda c: 25 + c
# ---- lane layout of the phase-1 intermediates
LS = lambda a, j: (0 if j == 2 else 4) + a    # S[a][j] = (P A)[a][j], j in {2,3}
LR = lambda c, j: (8 if j == 2 else 10) + c   # R[c][j] = (W' A)[c][j]
LPB = lambda a: 12 + a                        # Pb[a] = (P b + px)[a]
LX = lambda q: 16 + q                         # X0 = b2 P22, X1 = T P23, X2 = T P33 (for G_uu)

T1, T2A, T2B, ND = 5, 5, 4, 2


def term(src, f1=ONE):
    return (src, f1)


EMPTY = (0, ZERO)


def build():
    ph1 = [[] for _ in range(32)]
    for a in range(4):
        ph1[LS(a, 2)] = [term(LP(a, 0), A02), term(LP(a, 1), A12), term(LP(a, 2))]
        ph1[LS(a, 3)] = [term(LP(a, 0), A03), term(LP(a, 1), A13), term(LP(a, 2), A23), term(LP(a, 3))]
        ph1[LPB(a)] = [term(LPX(a))] + [term(LP(a, b), NEXT + CDEF + b) for b in range(4)]
    for c in range(2):
        ph1[LR(c, 2)] = [term(LW(0, c), A02), term(LW(1, c), A12), term(LW(2, c))]
        ph1[LR(c, 3)] = [term(LW(0, c), A03), term(LW(1, c), A13), term(LW(2, c), A23), term(LW(3, c))]
    ph1[LX(0)] = [term(LP(2, 2), B2)]
    ph1[LX(1)] = [term(LP(2, 3), TEE)]
    ph1[LX(2)] = [term(LP(3, 3), TEE)]
    pa = [[] for _ in range(32)]   # phase 2A: sources are value-function entries
    pb = [[] for _ in range(32)]   # phase 2B: sources are phase-1 intermediates
    dr = [[] for _ in range(32)]   # direct loads from the stage record
    dr[LP(0, 0)], pa[LP(0, 0)] = [HXX + 0], [term(LP(0, 0))]
    dr[LP(0, 1)], pa[LP(0, 1)] = [HXX + 1], [term(LP(0, 1))]
    dr[LP(1, 1)], pa[LP(1, 1)] = [HXX + 2], [term(LP(1, 1))]
    pb[LP(0, 2)] = [term(LS(0, 2))]
    pb[LP(0, 3)] = [term(LS(0, 3))]
    pb[LP(1, 2)] = [term(LS(1, 2))]
    pb[LP(1, 3)] = [term(LS(1, 3))]
    dr[LP(2, 2)], pb[LP(2, 2)] = [HXX + 3], [term(LS(2, 2)), term(LS(0, 2), A02), term(LS(1, 2), A12)]
    dr[LP(2, 3)], pb[LP(2, 3)] = [HXX + 4], [term(LS(2, 3)), term(LS(0, 3), A02), term(LS(1, 3), A12)]
    dr[LP(3, 3)], pb[LP(3, 3)] = [HXX + 5], [term(LS(3, 3)), term(LS(0, 3), A03), term(LS(1, 3), A13), term(LS(2, 3), A23)]
    # G_ux(c,j) on lane LW(j,c):  B' P A + W' A + Hux;   B[2][d] = b2, B[3][a] = T
    for c, (row, bf) in enumerate(((2, B2), (3, TEE))):
        pa[LW(0, c)] = [term(LP(row, 0), bf), term(LW(0, c))]
        pa[LW(1, c)] = [term(LP(row, 1), bf), term(LW(1, c))]
        pb[LW(2, c)] = [term(LS(row, 2), bf), term(LR(c, 2))]
        pb[LW(3, c)] = [term(LS(row, 3), bf), term(LR(c, 3))]
    dr[LW(3, 0)] = [HUX]
    # G_uu
    dr[18], pa[18], pb[18] = [HUU + 0, EE + 0], [term(18), term(LW(2, 0), B2), term(LW(2, 0), B2)], [term(LX(0), B2)]
    pa[19], pb[19] = [term(19), term(LW(2, 1), B2), term(LW(3, 0), TEE)], [term(LX(1), B2)]
    dr[20], pa[20], pb[20] = [HUU + 1, EE + 1], [term(20), term(LW(3, 1), TEE), term(LW(3, 1), TEE)], [term(LX(2), TEE)]
    # g_x = gx + A' Pb
    dr[LPX(0)], pb[LPX(0)] = [GX + 0], [term(LPB(0))]
    dr[LPX(1)], pb[LPX(1)] = [GX + 1], [term(LPB(1))]
    dr[LPX(2)], pb[LPX(2)] = [GX + 2], [term(LPB(2)), term(LPB(0), A02), term(LPB(1), A12)]
    dr[LPX(3)], pb[LPX(3)] = [GX + 3], [term(LPB(3)), term(LPB(0), A03), term(LPB(1), A13), term(LPB(2), A23)]
    # g_u = gu + t + pw + B' Pb + W' b
    for c, (row, bf) in enumerate(((2, B2), (3, TEE))):
        dr[LPW(c)] = [GU + c, TK + c]
        pa[LPW(c)] = [term(LPW(c))] + [term(LW(b, c), NEXT + CDEF + b) for b in range(4)]
        pb[LPW(c)] = [term(LPB(row), bf)]
    pad = lambda lst, n, e: [l + [e] * (n - len(l)) for l in lst]
    assert max(map(len, ph1)) <= T1 and max(map(len, pa)) <= T2A and max(map(len, pb)) <= T2B and max(map(len, dr)) <= ND
    return pad(ph1, T1, EMPTY), pad(pa, T2A, EMPTY), pad(pb, T2B, EMPTY), pad(dr, ND, ZERO)


def simulate(tables, rec, rec_next, v):
    """32-lane emulation of one stage; rec/rec_next = stage records (NF doubles); v = V entries per lane."""
    ph1, pa, pb, dr = tables
    both = np.concatenate([rec, rec_next])

    def run(tab, src, acc):
        out = acc.copy()
        for l in range(32):
            for (s_, f1) in tab[l]:
                out[l] += both[f1] * src[s_]
        return out

    s = run(ph1, v, np.zeros(32))
    g = np.array([sum(both[o] for o in dr[l]) for l in range(32)])
    g = run(pa, v, g)
    g = run(pb, s, g)
    return g


def reference(rec, rec_next, P, W, Q, px, pw):
    a02, a03, a12, a13, a23, b2 = rec[JAC:JAC + 6]
    T = rec[TEE]
    A = np.eye(4)
    A[0, 2], A[0, 3], A[1, 2], A[1, 3], A[2, 3] = a02, a03, a12, a13, a23
    B = np.zeros((4, 2))
    B[2, 0], B[3, 1] = b2, T
    b = rec_next[CDEF:CDEF + 4]
    Hxx = np.zeros((4, 4))
    Hxx[0, 0], Hxx[0, 1], Hxx[1, 1], Hxx[2, 2], Hxx[2, 3], Hxx[3, 3] = rec[HXX:HXX + 6]
    Hxx = np.triu(Hxx) + np.triu(Hxx, 1).T
    Hux = np.zeros((2, 4))
    Hux[0, 3] = rec[HUX]
    E = rec[EE:EE + 2]
    Gxx = Hxx + A.T @ P @ A
    Gux = Hux + B.T @ P @ A + W.T @ A
    Guu = np.diag(rec[HUU:HUU + 2] + E) + Q + B.T @ P @ B + B.T @ W + (B.T @ W).T
    Pb = P @ b + px
    gx = rec[GX:GX + 4] + A.T @ Pb
    gu = rec[GU:GU + 2] + rec[TK:TK + 2] + pw + B.T @ Pb + W.T @ b
    return Gxx, Gux, Guu, gx, gu


def check(tables):
    rng = np.random.default_rng(0)
    for _ in range(20):
        rec, rec_next = rng.standard_normal(NF), rng.standard_normal(NF)
        rec[ONE], rec[TEE], rec[ZERO] = 1.0, 0.1, 0.0
        Pm = rng.standard_normal((4, 4)); Pm = Pm + Pm.T
        W = rng.standard_normal((4, 2))
        Qm = rng.standard_normal((2, 2)); Qm = Qm + Qm.T
        px, pw = rng.standard_normal(4), rng.standard_normal(2)
        v = np.zeros(32)
        for i in range(4):
            for j in range(i, 4):
                v[LP(i, j)] = Pm[i, j]
            for c in range(2):
                v[LW(i, c)] = W[i, c]
            v[LPX(i)] = px[i]
        v[18], v[19], v[20] = Qm[0, 0], Qm[0, 1], Qm[1, 1]
        v[25], v[26] = pw
        g = simulate(tables, rec, rec_next, v)
        Gxx, Gux, Guu, gx, gu = reference(rec, rec_next, Pm, W, Qm, px, pw)
        err = 0.0
        for i in range(4):
            for j in range(i, 4):
                err = max(err, abs(g[LP(i, j)] - Gxx[i, j]))
            for c in range(2):
                err = max(err, abs(g[LW(i, c)] - Gux[c, i]))
            err = max(err, abs(g[LPX(i)] - gx[i]))
        err = max(err, abs(g[18] - Guu[0, 0]), abs(g[19] - Guu[0, 1]), abs(g[20] - Guu[1, 1]), abs(g[25] - gu[0]), abs(g[26] - gu[1]))
        assert err < 1e-12, err
    print("tables verified against the dense formulas")


def emit(tables):
    ph1, pa, pb, dr = tables
    pack = lambda t: t[0] | (t[1] << 5)   # src lane (5 bits) | coefficient offset in the record (may reach into the next record)
    L = ["// GENERATED by scripts/gen_kin_riccati_tables.py - do not edit.",
         "// Per-lane term tables of the lane-parallel backward Riccati stage (kinematic model).",
         "// term = src_lane | f1 << 5 ; coefficient = rec[f1] (offsets >= NF address the next stage's record)",
         "#pragma once", "namespace mpcb {",
         f"constexpr int KR_NF = {NF}, KR_T1 = {T1}, KR_T2A = {T2A}, KR_T2B = {T2B}, KR_ND = {ND};",
         f"constexpr int KR_ONE = {ONE}, KR_TEE = {TEE}, KR_ZERO = {ZERO};"]
    rows = []
    for t in range(T1):
        rows.append([pack(ph1[l][t]) for l in range(32)])
    for t in range(T2A):
        rows.append([pack(pa[l][t]) for l in range(32)])
    for t in range(T2B):
        rows.append([pack(pb[l][t]) for l in range(32)])
    for t in range(ND):
        rows.append([dr[l][t] for l in range(32)])
    L.append(f"__device__ const int kr_table[{len(rows)}][32] = {{")
    for r in rows:
        L.append("  {" + ", ".join(str(x) for x in r) + "},")
    L.append("};")
    L += ["}  // namespace mpcb"]
    out = os.path.join(HERE, "..", "mpc_motion_planning_b200", "csrc", "kin_riccati_tables.cuh")
    open(out, "w").write("\n".join(L) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    tb = build()
    check(tb)
    emit(tb)

"""Product-form triangular solves: explicit sparse inverses of the factors of the sparse pivots.

Why: a Newton solve on the generated pattern is 10 forward levels, a tail
level, a dense mat-vec and 11 backward levels -- 23 dependent levels, each bounded below by a block
barrier plus the ~200-instruction chain of the few warps that have rows in it.  Solves outnumber
factorisations 6:1.  The factors of the 247 sparse pivots are so sparse that their *inverses* are
too (default network: inv(L11) has 1 044 off-diagonal entries for 616 in L11, inv(U11) 1 019 for
476), so after every factorisation the kernel can form

    X = inv(L11)   (unit lower, closure pattern of L11)
    Y = inv(U11)   (upper, diagonal = the stored reciprocal pivots)

and a solve becomes five wide levels instead of 23 narrow ones:

    P1  y1 = b1 + X b1                      (rows < n0)
    P2  b2' = b2 - L21 y1                   (the existing tail program)
    P3  x2 = Tinv b2'                       (dense mat-vec)
    P4  w  = y1 - U12 x2
    P5  x1 = Y w

X and Y are computed by level-scheduled programs of the same kind as the factorisation (11 merged
levels: level k of X next to level k of Y) into a staging buffer -- the flux array, which is dead
outside the RHS -- and then copied into the Newton-matrix storage: entries that exist in L11 / U11
overwrite them in place (the originals are not needed any more), fill entries go to `nfill` extra
slots after the regular storage.

This module builds the patterns and programs from a :class:`~uclchem_b200.symbolic.Symbolic` and
provides numpy executors (used by the CPU tests and `tools/study_inverse_solve.py`).  The generator
emits the tables; the device side is compiled in with -DUCLGPU_PRODUCT_FORM (see engine_la.cuh).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .symbolic import Symbolic


@dataclass
class ProductForm:
    n0: int
    m: int
    nval: int                 # regular storage size (sym.nval)
    nval_pf: int              # storage size with the fill slots appended
    # staging indices: [0, nx) X entries, [nx, nx+ny) Y off-diagonal entries, then n0 diagonal copies, then ONE
    nx: int = 0
    ny: int = 0
    stg_diag0: int = 0
    stg_one: int = 0
    nstg: int = 0
    x_rc: np.ndarray = field(default=None)      # [nx, 2] (row, col) of X entries, new indexing
    y_rc: np.ndarray = field(default=None)      # [ny, 2]
    final_pos: np.ndarray = field(default=None)  # [nx+ny] storage position each staged entry is copied to
    nfill: int = 0
    # inverse program: list of levels; each level = list of items (stg target, stg index of the scale,
    # [(val pos, stg index), ...]);   stg[target] = -stg[scale] * sum(val[pos] * stg[idx])
    inv_levels: list = field(default=None)
    # solve programs, items (vector index, [(val pos, vector index), ...]) over final storage positions
    p1: list = field(default=None)   # tmpv[i] = xs[i] + sum X_ij xs[j]
    p4: list = field(default=None)   # xs[i] = tmpv[i] - sum u_ij tmpv[n0 + j']        (U12 part)
    p5: list = field(default=None)   # tmpv[i] = sum Y_ij xs[j]   (diagonal included as a term)
    stats: dict = field(default_factory=dict)


def build(sym: Symbolic) -> ProductForm:
    n0, m = sym.n0, sym.m
    Lrow = [[] for _ in range(n0)]
    U11 = [[] for _ in range(n0)]
    U12 = [[] for _ in range(n0)]
    for (i, j) in sym.ent_pos:
        if i < n0:
            if j < i:
                Lrow[i].append(j)
            elif j > i:
                (U11 if j < n0 else U12)[i].append(j)
    for r in (Lrow, U11, U12):
        for lst in r:
            lst.sort()
    pos = sym.ent_pos

    # ---- closure patterns -------------------------------------------------------------------
    CL = [set() for _ in range(n0)]          # X_ij != 0 for j in CL[i] (j < i)
    for i in range(n0):
        for k in Lrow[i]:
            CL[i].add(k)
            CL[i] |= CL[k]
    CU = [set() for _ in range(n0)]          # Y_ij != 0 for j in CU[i] (j > i)
    for i in range(n0 - 1, -1, -1):
        for k in U11[i]:
            CU[i].add(k)
            CU[i] |= CU[k]
    x_rc = [(i, j) for i in range(n0) for j in sorted(CL[i])]
    y_rc = [(i, j) for i in range(n0) for j in sorted(CU[i])]
    nx, ny = len(x_rc), len(y_rc)
    x_idx = {rc: k for k, rc in enumerate(x_rc)}
    y_idx = {rc: nx + k for k, rc in enumerate(y_rc)}
    stg_diag0 = nx + ny
    stg_one = stg_diag0 + n0
    nstg = stg_one + 1

    # ---- final storage positions: in place where the factor has an entry, fill slots otherwise ----
    final_pos, nfill = [], 0
    for rc in x_rc + y_rc:
        if rc in pos:
            final_pos.append(pos[rc])
        else:
            final_pos.append(sym.nval + nfill)
            nfill += 1
    fpos = {rc: p for rc, p in zip(x_rc + y_rc, final_pos)}

    # ---- inverse program ----------------------------------------------------------------------
    flev = np.zeros(n0, np.int64)
    for i in range(n0):
        flev[i] = 1 + max([flev[k] for k in Lrow[i]], default=-1)
    blev = np.zeros(n0, np.int64)
    for i in range(n0 - 1, -1, -1):
        blev[i] = 1 + max([blev[k] for k in U11[i]], default=-1)
    nlev = int(max(flev.max(), blev.max())) + 1
    inv_levels = [[] for _ in range(nlev)]
    for (i, j) in x_rc:
        # X_ij = -( l_ij [j in L(i)] + sum_{k in L(i), k > j, j in CL[k]} l_ik X_kj )
        terms = []
        for k in Lrow[i]:
            if k == j:
                terms.append((pos[(i, k)], stg_one))
            elif k > j and j in CL[k]:
                terms.append((pos[(i, k)], x_idx[(k, j)]))
        inv_levels[flev[i]].append((x_idx[(i, j)], stg_one, terms))
    for (i, j) in y_rc:
        # Y_ij = -d_i * sum_{k in U11(i), k <= j, (k == j or j in CU[k])} u_ik Y_kj ,  Y_kk = d_k
        terms = []
        for k in U11[i]:
            if k == j:
                terms.append((pos[(i, k)], stg_diag0 + k))
            elif k < j and j in CU[k]:
                terms.append((pos[(i, k)], y_idx[(k, j)]))
        inv_levels[blev[i]].append((y_idx[(i, j)], stg_diag0 + i, terms))
    inv_levels = [lv for lv in inv_levels if lv]

    # ---- solve programs over the final positions ----------------------------------------------------
    p1 = [(i, [(fpos[(i, j)], j) for j in sorted(CL[i])]) for i in range(n0)]
    p4 = [(i, [(pos[(i, j)], j) for j in U12[i]]) for i in range(n0)]
    p5 = [(i, [(int(sym.diag_pos[i]), i)] + [(fpos[(i, j)], j) for j in sorted(CU[i])]) for i in range(n0)]

    pf = ProductForm(n0=n0, m=m, nval=sym.nval, nval_pf=sym.nval + nfill, nx=nx, ny=ny, stg_diag0=stg_diag0,
                     stg_one=stg_one, nstg=nstg, x_rc=np.asarray(x_rc, np.int32).reshape(-1, 2),
                     y_rc=np.asarray(y_rc, np.int32).reshape(-1, 2), final_pos=np.asarray(final_pos, np.int32),
                     nfill=nfill, inv_levels=inv_levels, p1=p1, p4=p4, p5=p5)
    pf.stats = dict(nnz_L11=sum(map(len, Lrow)), nnz_U11=sum(map(len, U11)), nnz_U12=sum(map(len, U12)),
                    nnz_X=nx, nnz_Y=ny, nfill=nfill, inv_levels=len(inv_levels),
                    inv_terms=sum(len(t) for lv in inv_levels for _, _, t in lv),
                    p1_terms=sum(len(t) for _, t in p1), p4_terms=sum(len(t) for _, t in p4),
                    p5_terms=sum(len(t) for _, t in p5))
    return pf


# ------------------------------------------------------------------------------------------------
# numpy executors (what the device does, sequentially)
# ------------------------------------------------------------------------------------------------
def invert(pf: ProductForm, sym: Symbolic, val: np.ndarray) -> np.ndarray:
    """val: factored storage (TableEngine.factor).  Returns storage of size nval_pf with X / Y in
    their final positions (L11 / U11 overwritten, fill slots appended)."""
    stg = np.zeros(pf.nstg)
    stg[pf.stg_one] = 1.0
    stg[pf.stg_diag0: pf.stg_diag0 + pf.n0] = val[sym.diag_pos[: pf.n0]]
    for lv in pf.inv_levels:
        new = []
        for target, scale, terms in lv:        # all items of a level read the state before the level
            acc = 0.0
            for p, k in terms:
                acc += val[p] * stg[k]
            new.append((target, -stg[scale] * acc))
        for target, v in new:
            stg[target] = v
    out = np.zeros(pf.nval_pf)
    out[: pf.nval] = val
    out[pf.final_pos] = stg[: pf.nx + pf.ny]
    return out


def solve(pf: ProductForm, sym: Symbolic, valpf: np.ndarray, b_old: np.ndarray) -> np.ndarray:
    """Solve P x = b with the product form; b_old / result in OLD augmented indexing."""
    n0, m = pf.n0, pf.m
    xs = np.asarray(b_old, float)[sym.perm].copy()
    tmpv = np.zeros(sym.naug)
    for i, terms in pf.p1:
        tmpv[i] = xs[i] + sum(valpf[p] * xs[j] for p, j in terms)
    for t in range(m):
        a, b = sym.tail_l_ptr[t], sym.tail_l_ptr[t + 1]
        xs[n0 + t] -= np.dot(valpf[sym.tail_l_pos[a:b]], tmpv[sym.tail_l_col[a:b]])
    Tinv = valpf[sym.off_dense: sym.off_dense + m * m].reshape(m, m)
    tmpv[n0:] = Tinv @ xs[n0:]
    for i, terms in pf.p4:
        xs[i] = tmpv[i] - sum(valpf[p] * tmpv[j] for p, j in terms)
    for i, terms in pf.p5:
        tmpv[i] = sum(valpf[p] * xs[j] for p, j in terms)
    out = np.empty(sym.naug)
    out[sym.perm] = tmpv
    return out

"""Network-generation-time symbolic work for the B200 engine.

Everything here runs once per MakeRates network (never at run time):

* the *flux table* -- every reaction as ``rate_r * prod_k yext[f_rk]``
  (rules of reference ``reaction.py:779-819``);
* the *gather table* -- ydot_i as a signed sum of fluxes
  (``io_functions.py:562-581``), chunked for balanced parallel reduction;
* the *analytic Jacobian* of the three-phase RHS (``odes.f90``) including the
  dense couplings through BULK/SURFACE, totalSwap and the surface-growth term
  ``YDOT(SURFACE)`` (``odes.f90:4815-5153``), expressed as a SPARSE bordered
  system with two auxiliary unknowns (S = uncorrected surface growth,
  tau = totalSwap/safeMantle);
* a fill-reducing elimination order (Markowitz, diagonal pivots) and the
  *fixed-sparsity symbolic LU*: a level-scheduled gather program for the sparse
  rows, a dense trailing block that is inverted in place, and level-scheduled
  triangular-solve programs.

The same tables drive the CUDA kernels (emitted by :mod:`makerates_cuda`) and the
numpy emulator in :mod:`uclchem_b200.table_emulator` that the CPU tests use to
check them against the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .network import EXT_BLR, EXT_INV_SM, EXT_ONE, EXT_SWAP_SM, N_EXT, TYPE_ID, Network

# column kinds of a Jacobian term (what the differentiated factor was)
KIND_PLAIN, KIND_BLR, KIND_ISM = 0, 1, 2


@dataclass
class Symbolic:
    net: Network
    neq: int
    naug: int                 # neq + 2 (S, tau)
    iB: int
    iS: int
    iD: int
    iSg: int                  # aux unknown: uncorrected surface growth
    iTau: int                 # aux unknown: totalSwap/safeMantle
    fwidth: int
    flux_f: np.ndarray        # [nreac, fwidth] ext-state indices
    swap_reacs: np.ndarray    # BULKSWAP reaction ids (totalSwap terms)
    # ydot gather: species-major term list
    g_ptr: np.ndarray         # [nspec+1] CSR pointers over real species rows 0..nspec-1
    g_reac: np.ndarray        # reaction id per term
    g_sign: np.ndarray        # +1 gain / -1 loss
    # Jacobian terms (m-part), entry-major
    perm: np.ndarray = field(default=None)     # elimination order: new -> old
    iperm: np.ndarray = field(default=None)    # old -> new
    n0: int = 0               # first row/col of the dense trailing block (new indexing)
    m: int = 0                # dense block size
    # pattern after fill, new indexing, for rows/cols outside the dense block
    ent_row: np.ndarray = field(default=None)  # [nent] new row index
    ent_col: np.ndarray = field(default=None)
    ent_pos: dict = field(default=None)        # (newrow,newcol) -> storage index
    nval: int = 0             # total storage (sparse entries + m*m dense block + 1 zero slot)
    off_dense: int = 0
    zero_slot: int = 0
    # J assembly program: per storage position, list of (reaction, k, sign, kind, scale_is_gamma)
    j_ptr: np.ndarray = field(default=None)
    j_term: np.ndarray = field(default=None)   # packed uint32
    j_pos: np.ndarray = field(default=None)    # storage position of each assembled entry
    j_gamma: np.ndarray = field(default=None)  # 1: multiply by -gamma ; 0: multiply by -1 (aux rows)
    const_pos: np.ndarray = field(default=None)  # positions with constant values (+1/-1 rows)
    const_val: np.ndarray = field(default=None)
    diag_pos: np.ndarray = field(default=None)   # [naug] storage position of (n,n) in new indexing
    # positions touched by the transfer (c-part) terms, per surface/bulk pair
    tr_pos: np.ndarray = field(default=None)     # [nsurf, 10]
    tau_pos_b: np.ndarray = field(default=None)  # [nswap] position of (tau, b)
    tau_pos_B: int = 0
    tau_pos_S: int = 0
    dd_pos: int = 0
    # factor program (levels of entries)
    f_levels: list = field(default=None)       # list of dict(target[], kind[], diag[], lptr[], terms[])
    # solve programs
    fwd_levels: list = field(default=None)
    bwd_levels: list = field(default=None)
    tail_l_ptr: np.ndarray = field(default=None)
    tail_l_col: np.ndarray = field(default=None)
    tail_l_pos: np.ndarray = field(default=None)
    stats: dict = field(default_factory=dict)


def pack_jterm(r, k, neg, kind):
    return (int(r) << 8) | (int(k) << 4) | (int(kind) << 1) | int(neg)


def unpack_jterm(t):
    t = np.asarray(t, dtype=np.int64)
    return t >> 8, (t >> 4) & 0xF, (t >> 1) & 0x7, t & 1


def build(net: Network, dense_threshold: float = 0.90) -> Symbolic:
    nspec, neq = net.nspec, net.neq
    iB, iS, iD = net.species_idx["nbulk"], net.species_idx["nsurface"], neq - 1
    iSg, iTau = neq, neq + 1
    naug = neq + 2
    F5 = net.flux_factors(5)
    one = neq + EXT_ONE
    fwidth = int((F5 != one).sum(axis=1).max())
    flux_f = F5[:, :fwidth].copy()
    ls, lr, gs, gr = net.stoichiometry()

    # ---- ydot gather table (species-major) --------------------------------
    terms = [[] for _ in range(nspec)]
    for s, r in zip(ls, lr):
        terms[s].append((int(r), -1))
    for s, r in zip(gs, gr):
        terms[s].append((int(r), +1))
    g_ptr = np.zeros(nspec + 1, np.int64)
    g_reac, g_sign = [], []
    for i in range(nspec):
        # net stoichiometric coefficient per reaction: A + A -> ... appears twice, keep multiplicity
        for r, sg in terms[i]:
            g_reac.append(r)
            g_sign.append(sg)
        g_ptr[i + 1] = len(g_reac)
    g_reac = np.asarray(g_reac, np.int32)
    g_sign = np.asarray(g_sign, np.int8)

    lo, hi = net.type_ranges["BULKSWAP"]
    swap_reacs = np.arange(lo, hi + 1, dtype=np.int32)

    sym = Symbolic(net=net, neq=neq, naug=naug, iB=iB, iS=iS, iD=iD, iSg=iSg, iTau=iTau, fwidth=fwidth,
                   flux_f=flux_f, swap_reacs=swap_reacs, g_ptr=g_ptr, g_reac=g_reac, g_sign=g_sign)

    # ---- Jacobian m-part terms, keyed by (row, col) in OLD augmented indexing ----
    jt: dict = {}

    def add(i, j, r, k, sign, kind):
        jt.setdefault((i, j), []).append((r, k, sign, kind))

    is_surf = np.zeros(naug, bool)
    is_surf[net.surface_list] = True
    for i in range(nspec):
        if i in (iB, iS):
            continue
        for r, sg in terms[i]:
            for k in range(fwidth):
                f = int(flux_f[r, k])
                if f == one:
                    continue
                if f < neq:
                    col, kind = f, KIND_PLAIN
                elif f == neq + EXT_BLR:
                    col, kind = iB, KIND_BLR
                elif f == neq + EXT_INV_SM:
                    col, kind = iS, KIND_ISM
                elif f == neq + EXT_SWAP_SM:
                    col, kind = iTau, KIND_PLAIN
                else:
                    raise AssertionError
                add(i, col, r, k, sg, kind)
                if is_surf[i]:
                    add(iSg, col, r, k, sg, kind)

    # structural pattern (old indexing)
    pat = [set([i]) for i in range(naug)]
    for (i, j) in jt:
        pat[i].add(j)
    for r in swap_reacs:
        pat[iTau].add(int(net.re[r, 0]))
    pat[iTau] |= {iB, iS}
    for s, b in zip(net.surface_list, net.bulk_list):
        s, b = int(s), int(b)
        for i in (s, b):
            pat[i] |= {iSg, b, s, iB, iS}
    for b in net.bulk_list:
        pat[iB].add(int(b))
    for s in net.surface_list:
        pat[iS].add(int(s))
    sym.stats["nnz_P"] = sum(len(p) for p in pat)

    # ---- ordering: Markowitz with diagonal pivots; border forced last --------
    last = [iB, iS, iSg, iTau]
    perm = _markowitz(pat, last)
    iperm = np.empty(naug, np.int64)
    iperm[perm] = np.arange(naug)
    sym.perm, sym.iperm = np.asarray(perm, np.int64), iperm

    # ---- symbolic factorisation (new indexing) -------------------------------
    rows = [set(int(iperm[c]) for c in pat[perm[n]]) for n in range(naug)]
    for i in range(naug):
        done = set()
        while True:
            ks = [k for k in rows[i] if k < i and k not in done]
            if not ks:
                break
            k = min(ks)
            done.add(k)
            rows[i] |= {j for j in rows[k] if j > k}
    # choose the dense tail: smallest n0 whose trailing block is >= threshold dense
    n0 = naug
    for cand in range(naug - 8, 0, -1):
        mm = naug - cand
        cnt = sum(len([j for j in rows[i] if j >= cand]) for i in range(cand, naug))
        if cnt / (mm * mm) >= dense_threshold:
            n0 = cand
        else:
            break
    m = naug - n0
    sym.n0, sym.m = n0, m
    for i in range(n0, naug):
        rows[i] |= set(range(n0, naug))

    # storage: sparse entries (row<n0 or col<n0) in row-major order, then dense block, then zero slot
    ent_pos = {}
    ent_row, ent_col = [], []
    for i in range(naug):
        for j in sorted(rows[i]):
            if i >= n0 and j >= n0:
                continue
            ent_pos[(i, j)] = len(ent_row)
            ent_row.append(i)
            ent_col.append(j)
    off_dense = len(ent_row)
    for i in range(n0, naug):
        for j in range(n0, naug):
            ent_pos[(i, j)] = off_dense + (i - n0) * m + (j - n0)
    zero_slot = off_dense + m * m
    sym.ent_row, sym.ent_col = np.asarray(ent_row, np.int32), np.asarray(ent_col, np.int32)
    sym.ent_pos, sym.off_dense, sym.zero_slot, sym.nval = ent_pos, off_dense, zero_slot, zero_slot + 1
    sym.diag_pos = np.asarray([ent_pos[(n, n)] for n in range(naug)], np.int32)
    Lrow = [sorted(j for j in rows[i] if j < i) for i in range(naug)]
    Urow = [sorted(j for j in rows[i] if j > i) for i in range(naug)]
    sym.stats.update(nnz_L=sum(map(len, Lrow)), nnz_U=sum(map(len, Urow)), n0=n0, m=m, nsparse=off_dense)

    def pos_old(i_old, j_old):
        return ent_pos[(int(iperm[i_old]), int(iperm[j_old]))]

    # ---- J assembly program ---------------------------------------------------
    j_ptr, j_term, j_pos, j_gamma = [0], [], [], []
    for (i, j), lst in sorted(jt.items(), key=lambda kv: pos_old(*kv[0])):
        j_pos.append(pos_old(i, j))
        j_gamma.append(0 if i == iSg else 1)
        for (r, k, sg, kind) in lst:
            j_term.append(pack_jterm(r, k, 1 if sg < 0 else 0, kind))
        j_ptr.append(len(j_term))
    sym.j_ptr = np.asarray(j_ptr, np.int64)
    sym.j_term = np.asarray(j_term, np.uint32)
    sym.j_pos = np.asarray(j_pos, np.int32)
    sym.j_gamma = np.asarray(j_gamma, np.int8)
    cpos, cval = [], []
    for b in net.bulk_list:
        cpos.append(pos_old(iB, int(b)))
        cval.append(-1.0)
    for s in net.surface_list:
        cpos.append(pos_old(iS, int(s)))
        cval.append(-1.0)
    sym.const_pos, sym.const_val = np.asarray(cpos, np.int32), np.asarray(cval)
    tr = []
    for s, b in zip(net.surface_list, net.bulk_list):
        s, b = int(s), int(b)
        tr.append([pos_old(s, iSg), pos_old(b, iSg), pos_old(s, b), pos_old(b, b), pos_old(s, s), pos_old(b, s),
                   pos_old(s, iB), pos_old(b, iB), pos_old(s, iS), pos_old(b, iS)])
    sym.tr_pos = np.asarray(tr, np.int32)
    sym.tau_pos_b = np.asarray([pos_old(iTau, int(net.re[r, 0])) for r in swap_reacs], np.int32)
    sym.tau_pos_B, sym.tau_pos_S = pos_old(iTau, iB), pos_old(iTau, iS)
    sym.dd_pos = pos_old(iD, iD)
    sym.stats["j_terms"] = len(j_term)
    sym.stats["j_entries"] = len(j_pos)

    # ---- factor program: entry gather, levelled ---------------------------------
    # entry (i,j) with min(i,j) < n0:  a_ij -= sum_{k<min(i,j), k in L(i), j in U(k)} l_ik u_kj ; L entries scaled by 1/u_jj
    # dense-block entries: Schur update with k < n0 only.
    Uset = [set(u) for u in Urow]
    lev = {}
    entries = []
    for i in range(naug):
        for j in sorted(rows[i]):
            kmax = min(i, j, n0)
            ks = [k for k in Lrow[i] if k < kmax and (j in Uset[k])]
            entries.append((i, j, ks))
    # level by dependency (process in an order where dependencies come first: row-major works since
    # (i,k) is earlier in row i and (k,j) is in an earlier row)
    for (i, j, ks) in entries:
        lv = 0
        for k in ks:
            lv = max(lv, lev[(i, k)] + 1, lev[(k, j)] + 1)
        if j < i and j < n0:
            lv = max(lv, lev[(j, j)] + 1)
        lev[(i, j)] = lv
    nlev = max(lev.values()) + 1
    f_levels = []
    tot_terms = 0
    for lv in range(nlev):
        # every sparse pivot (i == j < n0) is an item even without terms: it is stored as its reciprocal
        es = [(i, j, ks) for (i, j, ks) in entries
              if lev[(i, j)] == lv and (ks or (j < i and j < n0) or (i == j and i < n0))]
        if not es:
            continue
        es.sort(key=lambda e: -len(e[2]))
        target = [ent_pos[(i, j)] for (i, j, _) in es]
        diag = [ent_pos[(j, j)] if (j < i and j < n0) else (-2 if (i == j and i < n0) else -1) for (i, j, _) in es]
        lptr = [0]
        tl, tu = [], []
        for (i, j, ks) in es:
            for k in ks:
                tl.append(ent_pos[(i, k)])
                tu.append(ent_pos[(k, j)])
            lptr.append(len(tl))
        tot_terms += len(tl)
        f_levels.append(dict(target=np.asarray(target, np.int32), diag=np.asarray(diag, np.int32),
                             ptr=np.asarray(lptr, np.int64), tl=np.asarray(tl, np.int32),
                             tu=np.asarray(tu, np.int32)))
    sym.f_levels = f_levels
    sym.stats.update(factor_levels=len(f_levels), factor_terms=tot_terms)

    # ---- solve programs -------------------------------------------------------------
    # forward (unit L): rows n < n0 in levels; x_n = b_n - sum_{k in L(n)} l_nk x_k
    flev = np.zeros(n0, np.int64)
    for n in range(n0):
        flev[n] = 1 + max([flev[k] for k in Lrow[n]], default=-1)
    fwd_levels = []
    for lv in range(int(flev.max()) + 1 if n0 else 0):
        rws = [n for n in range(n0) if flev[n] == lv and Lrow[n]]
        if not rws:
            continue
        rws.sort(key=lambda n: -len(Lrow[n]))
        ptr, cols, poss = [0], [], []
        for n in rws:
            for k in Lrow[n]:
                cols.append(k)
                poss.append(ent_pos[(n, k)])
            ptr.append(len(cols))
        fwd_levels.append(dict(rows=np.asarray(rws, np.int32), ptr=np.asarray(ptr, np.int64),
                               cols=np.asarray(cols, np.int32), pos=np.asarray(poss, np.int32)))
    sym.fwd_levels = fwd_levels
    # tail rows: b_T -= L21 x_sparse
    tptr, tcol, tpos = [0], [], []
    for n in range(n0, naug):
        for k in Lrow[n]:
            if k < n0:
                tcol.append(k)
                tpos.append(ent_pos[(n, k)])
        tptr.append(len(tcol))
    sym.tail_l_ptr, sym.tail_l_col, sym.tail_l_pos = (np.asarray(tptr, np.int64), np.asarray(tcol, np.int32),
                                                      np.asarray(tpos, np.int32))
    # backward: rows n < n0 in reverse levels; x_n = (b_n - sum_{j in U(n)} u_nj x_j) / u_nn
    blev = np.zeros(n0, np.int64)
    for n in range(n0 - 1, -1, -1):
        blev[n] = 1 + max([blev[j] for j in Urow[n] if j < n0], default=-1)
    bwd_levels = []
    for lv in range(int(blev.max()) + 1 if n0 else 0):
        rws = [n for n in range(n0) if blev[n] == lv]
        rws.sort(key=lambda n: -len(Urow[n]))
        ptr, cols, poss = [0], [], []
        for n in rws:
            for j in Urow[n]:
                cols.append(j)
                poss.append(ent_pos[(n, j)])
            ptr.append(len(cols))
        bwd_levels.append(dict(rows=np.asarray(rws, np.int32), ptr=np.asarray(ptr, np.int64),
                               cols=np.asarray(cols, np.int32), pos=np.asarray(poss, np.int32)))
    sym.bwd_levels = bwd_levels
    sym.stats.update(fwd_levels=len(fwd_levels), bwd_levels=len(bwd_levels),
                     fwd_terms=int(sum(len(l["cols"]) for l in fwd_levels)),
                     bwd_terms=int(sum(len(l["cols"]) for l in bwd_levels)), tail_terms=len(tcol))
    return sym


def _markowitz(pat, last):
    """Greedy Markowitz ordering with diagonal pivots on the structural pattern;
    variables in `last` are eliminated at the end in the given order."""
    n = len(pat)
    rows = [set(p) for p in pat]
    cols = [set() for _ in range(n)]
    for i in range(n):
        for c in rows[i]:
            cols[c].add(i)
    remaining = set(range(n)) - set(last)
    order = []

    def eliminate(k):
        rk = rows[k] - {k}
        ck = cols[k] - {k}
        for i in ck:
            new = rk - rows[i]
            rows[i] |= rk
            for j in new:
                cols[j].add(i)
            rows[i].discard(k)
        for j in rk:
            cols[j].discard(k)
        rows[k] = set()
        cols[k] = set()

    while remaining:
        k = min(remaining, key=lambda q: ((len(rows[q]) - 1) * (len(cols[q]) - 1), q))
        remaining.discard(k)
        order.append(k)
        eliminate(k)
    for k in last:
        order.append(k)
        eliminate(k)
    return order

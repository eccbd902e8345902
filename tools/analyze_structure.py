"""Exploration: Jacobian pattern, ordering, fill and level structure."""
import sys
import numpy as np
from uclchem_b200.network import *

net = load_default()
neq = net.neq
NS = net.nspec
iB, iS, iD = net.species_idx["nbulk"], net.species_idx["nsurface"], neq - 1
iSg, iTau = neq, neq + 1  # aux unknowns
N = neq + 2
F = net.flux_factors()
ls, lr, gs, gr = net.stoichiometry()
rows_of_reac = [[] for _ in range(net.nreac)]
for s, r in zip(ls, lr):
    rows_of_reac[r].append(s)
for s, r in zip(gs, gr):
    rows_of_reac[r].append(s)

pat = [set() for _ in range(N)]  # pat[i] = columns
for i in range(N):
    pat[i].add(i)
is_surf = np.zeros(N, bool); is_surf[net.surface_list] = True
is_bulk = np.zeros(N, bool); is_bulk[net.bulk_list] = True
ext0 = neq
for r in range(net.nreac):
    cols = set()
    for f in F[r]:
        if f < neq:
            cols.add(int(f))
        elif f == ext0 + EXT_BLR:
            cols.add(iB)
        elif f == ext0 + EXT_INV_SM:
            cols.add(iS)
        elif f == ext0 + EXT_SWAP_SM:
            cols.add(iTau)
    for i in set(rows_of_reac[r]):
        pat[i] |= cols
        if is_surf[i]:
            pat[iSg] |= cols
# tau row: depends on bulk species with swap reactions, BULK, SURF
lo, hi = net.type_ranges["BULKSWAP"]
for r in range(lo, hi + 1):
    pat[iTau].add(int(net.re[r, 0]))
pat[iTau] |= {iB, iS}
# transfer terms
for s, b in zip(net.surface_list, net.bulk_list):
    s = int(s); b = int(b)
    for i in (s, b):
        pat[i] |= {iSg, b, s, iB, iS}
# BULK / SURF rows = sums
for b in net.bulk_list:
    pat[iB] |= pat[int(b)]
for s in net.surface_list:
    pat[iS] |= pat[int(s)]
nnz = sum(len(p) for p in pat)
print("N", N, "nnz(P)", nnz)
core = [i for i in range(N) if i not in (iB, iS, iD, iSg, iTau)]
print("core nnz", sum(len([c for c in pat[i] if c in set(core)]) for i in core))
colcount = np.zeros(N, int)
for i in range(N):
    for c in pat[i]:
        colcount[c] += 1
print("densest cols", sorted([(colcount[c], c) for c in range(N)], reverse=True)[:12])
print("densest rows", sorted([(len(pat[i]), i) for i in range(N)], reverse=True)[:12])


def markowitz(pat, last):
    N = len(pat)
    rows = [set(p) for p in pat]
    cols = [set() for _ in range(N)]
    for i in range(N):
        for c in rows[i]:
            cols[c].add(i)
    remaining = set(range(N)) - set(last)
    order = []
    for phase in (0, 1):
        cand = remaining if phase == 0 else list(last)
        while cand:
            if phase == 0:
                best = min(cand, key=lambda k: ((len(rows[k]) - 1) * (len(cols[k]) - 1), k))
                cand.discard(best)
            else:
                best = cand.pop(0)
            k = best
            order.append(k)
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
            # freeze row k / col k as final L/U
            rows[k] = rk | {k}
            cols[k] = ck | {k}
    return order


def symbolic(pat, order):
    N = len(pat)
    pos = {k: n for n, k in enumerate(order)}
    rows = [set(pos[c] for c in pat[order[n]]) for n in range(N)]
    # row-wise up-looking symbolic: row i pattern = union over k in L(i) of U(k)
    for i in range(N):
        done = set()
        while True:
            ks = sorted(k for k in rows[i] if k < i and k not in done)
            if not ks:
                break
            k = ks[0]
            done.add(k)
            rows[i] |= {j for j in rows[k] if j > k}
    return rows


last = [iB, iS, iSg, iTau, iD]
order = markowitz(pat, last)
rows = symbolic(pat, order)
nL = sum(len([j for j in rows[i] if j < i]) for i in range(N))
nU = sum(len([j for j in rows[i] if j > i]) for i in range(N))
print("nnz L", nL, "U", nU, "total", nL + nU + N)
# flops for factorization: for each i, for k in L(i): 1 div + |U(k)| fma
Urow = [sorted(j for j in rows[i] if j > i) for i in range(N)]
Lrow = [sorted(j for j in rows[i] if j < i) for i in range(N)]
fma = sum(len(Urow[k]) for i in range(N) for k in Lrow[i])
print("factor fma", fma, "divs", nL)
# levels: row i depends on rows k in L(i)
lev = np.zeros(N, int)
for i in range(N):
    lev[i] = 1 + max([lev[k] for k in Lrow[i]], default=-1)
print("LU/fwd levels", lev.max() + 1, np.bincount(lev)[:40])
# backward levels: x_i depends on x_j for j in U(i)
blev = np.zeros(N, int)
for i in range(N - 1, -1, -1):
    blev[i] = 1 + max([blev[j] for j in Urow[i]], default=-1)
print("bwd levels", blev.max() + 1, np.bincount(blev)[:40])
# dense tail: find smallest n0 such that trailing block is >70% dense
for n0 in range(N - 150, N - 5, 5):
    m = N - n0
    cnt = sum(len([j for j in rows[i] if j >= n0]) for i in range(n0, N))
    print(n0, m, cnt / (m * m))
np.save("/tmp/order.npy", np.array(order))

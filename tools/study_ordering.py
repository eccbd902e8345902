"""CPU study: elimination orderings that trade fill for depth.

The sparse factorisation is latency bound: every one of its 18 levels costs a block barrier plus a
dependent chain, whatever its width.  Greedy Markowitz minimises fill, not depth.
This script replaces it by a multi-elimination ordering (each round eliminates a maximal independent
set of cheap pivots, so a round is one level of pivots) and reports storage, level counts and term
counts of the resulting programs for a few acceptance thresholds, next to the Markowitz baseline.
Nothing in the library uses it yet: no-pivot LU stability under a different order has to be checked on
the GPU parity tests first.
usage: study_ordering.py [network-tag]
"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
from uclchem_b200 import product_form, symbolic
from uclchem_b200.network import Network

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
net = Network.from_json(ROOT / "uclchem_b200" / "networks" / f"{tag}.json")


def multi_eliminate(slack, cap):
    def order(pat, last):
        n = len(pat)
        rows = [set(p) for p in pat]
        cols = [set() for _ in range(n)]
        for i in range(n):
            for c in rows[i]:
                cols[c].add(i)
        remaining = set(range(n)) - set(last)
        out = []

        def eliminate(k):
            rk, ck = rows[k] - {k}, cols[k] - {k}
            for i in ck:
                new = rk - rows[i]
                rows[i] |= rk
                for j in new:
                    cols[j].add(i)
                rows[i].discard(k)
            for j in rk:
                cols[j].discard(k)
            rows[k], cols[k] = set(), set()

        while remaining:
            cost = {q: (len(rows[q]) - 1) * (len(cols[q]) - 1) for q in remaining}
            cmin = min(cost.values())
            bound = max(cap, slack * cmin)
            chosen, blocked = [], set()
            for q in sorted(remaining, key=lambda q: (cost[q], q)):
                if cost[q] > bound:
                    break
                if q in blocked:
                    continue
                chosen.append(q)
                blocked |= rows[q] | cols[q]
            for q in chosen:
                remaining.discard(q)
                out.append(q)
            for q in chosen:
                eliminate(q)
        for k in last:
            out.append(k)
            eliminate(k)
        return out
    return order


def report(name):
    for thr in (0.9,):
        s = symbolic.build(net, thr)
        pf = product_form.build(s)
        st = s.stats
        print(f"{name:28s} thr {thr}: n0 {s.n0:3d} m {s.m:3d} nval {s.nval:6d} ({s.nval * 8 / 1024:6.1f} KB) "
              f"factor levels {st['factor_levels']:2d} terms {st['factor_terms']:6d} | fwd {st['fwd_levels']:2d} bwd {st['bwd_levels']:2d} "
              f"| product form: inverse levels {pf.stats['inv_levels']:2d} X {pf.nx} Y {pf.ny} fill {pf.nfill}")


base = symbolic._markowitz
report("markowitz (library)")
for slack, cap in ((1.0, 0), (1.5, 2), (2.0, 4), (3.0, 8), (4.0, 16)):
    symbolic._markowitz = multi_eliminate(slack, cap)
    report(f"multi-elim slack {slack} cap {cap}")
symbolic._markowitz = base

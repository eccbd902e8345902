"""Product-form solves (uclchem_b200/product_form.py): explicit sparse inverses of the factors of the
sparse pivots, five wide levels per Newton solve instead of 23 narrow ones.  Everything the device
will execute is checked here on the CPU: the patterns, the numpy executors, and the generated team
programs run through the TeamProgram interpreter with the device's buffer discipline."""
import numpy as np
import pytest

from uclchem_b200 import product_form
from uclchem_b200.makerates_cuda import Generated
from uclchem_b200.table_emulator import TableEngine

from conftest import GOLDEN


@pytest.fixture(scope="module")
def gen(net):
    return Generated(net)


@pytest.fixture(scope="module")
def eng(gen):
    return TableEngine(gen.sym)


def _state(gen, case):
    sym, net = gen.sym, gen.net
    g = np.load(GOLDEN / "getydot_cases.npz")
    y, rate = g[f"y_{case}"].copy(), g[f"rate_{case}"]
    y[sym.iB] = y[net.bulk_list].sum()
    y[sym.iS] = y[net.surface_list].sum()
    return y, rate


def _rel(x, ref):
    big = np.abs(ref) > 1e-12 * np.abs(ref).max()
    return float((np.abs(x - ref) / (np.abs(ref) + 1e-300))[big].max())


def test_patterns(gen):
    pf, sym = gen.pf, gen.sym
    s = pf.stats
    assert pf.nx == s["nnz_X"] >= s["nnz_L11"] and pf.ny == s["nnz_Y"] >= s["nnz_U11"]
    assert pf.nfill == (pf.nx - s["nnz_L11"]) + (pf.ny - s["nnz_U11"])       # closures contain the factors
    assert pf.nval_pf == sym.nval + pf.nfill and pf.nstg == pf.nx + pf.ny + sym.n0 + 1
    # final positions: distinct, in the sparse part or in the fill slots, never the dense block / zero slot
    fp = pf.final_pos
    assert len(set(fp.tolist())) == len(fp)
    assert ((fp < sym.off_dense) | (fp >= sym.nval)).all() and fp.max() == pf.nval_pf - 1
    # the Newton matrix with the fill slots must still fit next to the rest of the CTA's shared memory
    assert pf.nfill * 8 <= 11 * 1024
    # default network: the inverses are barely denser than the factors
    assert pf.nx < 2 * s["nnz_L11"] and pf.ny < 3 * s["nnz_U11"]


@pytest.mark.parametrize("case,gamma", [(0, 1e3), (1, 1e6), (2, 1e9), (3, 1e12), (5, 1e4)])
def test_numpy_executors_match_substitution(gen, eng, case, gamma):
    pf, sym = gen.pf, gen.sym
    y, rate = _state(gen, case)
    val = eng.factor(eng.assemble(y, rate, gamma))
    vpf = product_form.invert(pf, sym, val)
    rng = np.random.default_rng(case)
    b = np.zeros(sym.naug)
    b[: sym.neq] = rng.standard_normal(sym.neq) * (np.abs(y[: sym.neq]) + 1e-20)
    b[sym.iB] = b[sym.iS] = 0.0
    assert _rel(product_form.solve(pf, sym, vpf, b), eng.solve(val, b)) < 1e-5


def _run(prog, term_value, finalize):
    prog.run(term_value, finalize)


@pytest.mark.parametrize("case,gamma", [(0, 1e3), (3, 1e10)])
def test_generated_programs_with_device_buffer_discipline(gen, eng, case, gamma):
    """Interpret the emitted team programs exactly as engine_la.cuh does under UCLGPU_PRODUCT_FORM:
    inverse levels into the staging buffer, copy to the final positions, then P1..P5 on xs / tmpv."""
    pf, sym = gen.pf, gen.sym
    n0, m = sym.n0, sym.m
    y, rate = _state(gen, case)
    val0 = eng.factor(eng.assemble(y, rate, gamma))
    val = np.zeros(pf.nval_pf)
    val[: sym.nval] = val0
    stg = np.zeros(pf.nstg)
    stg[pf.stg_one] = 1.0
    stg[pf.stg_diag0: pf.stg_diag0 + n0] = val[sym.diag_pos[:n0]]
    for prog in gen.pf_inv:                      # one level = reads before writes (block barrier after it)
        out = {}
        prog.run(lambda t: val[t >> 16] * stg[t & 0xFFFF],
                 lambda tg, acc: out.__setitem__(tg, -stg[gen.pf_inv_scale[tg]] * acc))
        for tg, v in out.items():
            stg[tg] = v
    val[pf.final_pos] = stg[: pf.nx + pf.ny]
    ref = product_form.invert(pf, sym, val0)         # same recurrences, sequential summation order
    assert np.allclose(val, ref, rtol=1e-8, atol=1e-12 * np.abs(ref[pf.final_pos]).max())

    rng = np.random.default_rng(7)
    b = np.zeros(sym.naug)
    b[: sym.neq] = rng.standard_normal(sym.neq) * (np.abs(y[: sym.neq]) + 1e-20)
    b[sym.iB] = b[sym.iS] = 0.0
    xs = b[sym.perm].copy()
    tmpv = np.full(sym.naug, np.nan)
    gen.pf_p1.run(lambda t: val[t >> 16] * xs[t & 0xFFFF], lambda tg, acc: tmpv.__setitem__(tg, xs[tg] + acc))
    gen.pf_tail.run(lambda t: val[t >> 16] * tmpv[t & 0xFFFF], lambda tg, acc: xs.__setitem__(tg, xs[tg] - acc))
    Tinv = val[sym.off_dense: sym.off_dense + m * m].reshape(m, m)
    tmpv[n0:] = Tinv @ xs[n0:]
    gen.pf_p4.run(lambda t: val[t >> 16] * tmpv[t & 0xFFFF], lambda tg, acc: xs.__setitem__(tg, tmpv[tg] - acc))
    gen.pf_p5.run(lambda t: val[t >> 16] * xs[t & 0xFFFF], lambda tg, acc: tmpv.__setitem__(tg, acc))
    x = np.empty(sym.naug)
    x[sym.perm] = tmpv
    assert np.isfinite(x).all()
    assert _rel(x, eng.solve(val0, b)) < 1e-6
    # and it really solves the system: residual against the dense matrix
    A = eng.to_dense(eng.assemble(y, rate, gamma))
    r = A @ x[sym.perm] - b[sym.perm]
    assert np.abs(r).max() <= 1e-6 * np.abs(b).max()


def _items(prog):
    """(target, [terms]) of every real team of a TeamProgram, decoded from its descriptor table."""
    d = prog.desc
    out = []
    for s in range(prog.nslots):
        begin, w = int(d[s, 0]), int(d[s, 1])
        target, n, tl = w & 0xFFFF, (w >> 16) & 0xFFF, (w >> 28) & 7
        if target != 0xFFFF and (s & ((1 << tl) - 1)) == 0:
            out.append((target, [int(t) for t in prog.terms[begin: begin + n]]))
    return out


def test_levels_are_race_free(gen):
    """Inside one level (everything between two block barriers) no item may read a location another
    item of the same level writes.  Checked on the emitted programs for the buffer roles the device
    uses: factor (val -> val in place), forward / backward substitution (xs in place), the inverse
    program (val, stg -> stg) and the five product-form solve levels (xs <-> tmpv)."""
    sym, pf = gen.sym, gen.pf
    # factorisation: reads val[l], val[u], val[diag]; writes val[target]
    diag = gen.factor_diag_table()
    for prog, _ in gen.factor:
        items = _items(prog)
        writes = {t for t, _ in items}
        for t, terms in items:
            reads = {x >> 16 for x in terms} | {x & 0xFFFF for x in terms}
            if diag[t] < 0xFFFE:
                reads.add(int(diag[t]))
            assert not (reads & (writes - {t}))
    # substitutions: reads xs[col]; writes xs[target]
    for prog in gen.fwd + [gen.tail] + gen.bwd:
        items = _items(prog)
        writes = {t for t, _ in items}
        for t, terms in items:
            assert not ({x & 0xFFFF for x in terms} & (writes - {t}))
    # inverse program: reads val (never written here) and stg; writes stg[target]; scale read from stg
    written_before = set(range(pf.stg_diag0, pf.nstg))            # diagonal copies + ONE, staged up front
    for prog in gen.pf_inv:
        items = _items(prog)
        writes = {t for t, _ in items}
        for t, terms in items:
            reads = {x & 0xFFFF for x in terms} | {int(gen.pf_inv_scale[t])}
            assert not (reads & writes)                          # nothing of this level is read in it
            assert reads <= written_before                       # everything read was produced earlier
            assert all((x >> 16) < sym.nval for x in terms)      # factor entries only, no fill slot
        written_before |= writes
    assert written_before >= set(range(pf.nx + pf.ny))           # every X / Y entry is produced
    # product-form solve: each level reads one vector and writes the other
    n0 = sym.n0
    for prog, rd_lo, rd_hi, wr_lo, wr_hi in ((gen.pf_p1, 0, n0, 0, n0), (gen.pf_tail, 0, n0, n0, sym.naug),
                                             (gen.pf_p4, n0, sym.naug, 0, n0), (gen.pf_p5, 0, n0, 0, n0)):
        items = _items(prog)
        assert sorted(t for t, _ in items) == list(range(wr_lo, wr_hi))     # every row of the block is written once
        for t, terms in items:
            assert all(rd_lo <= (x & 0xFFFF) < rd_hi for x in terms)
            assert all((x >> 16) < pf.nval_pf for x in terms)

"""Generator tables (RHS gather, analytic Jacobian, symbolic LU, solve programs) validated on
the CPU with the numpy emulator against the oracle and dense linear algebra."""
import numpy as np
import pytest
from conftest import GOLDEN

from uclchem_b200.makerates_cuda import Generated
from uclchem_b200.table_emulator import COV0, TableEngine, _gauss_jordan_inverse


@pytest.fixture(scope="module")
def gen(net):
    return Generated(net)


@pytest.fixture(scope="module")
def eng(gen):
    return TableEngine(gen.sym)


def test_network_sizes(net, gen):
    assert (net.nspec, net.nreac) == (335, 3203)           # f2py_constants.f90:3
    assert net.type_ranges["TWOBODY"] == (1235, 3202)      # network.f90:5169 (0-based here)
    assert net.type_ranges["ER"] == (492, 492)             # single ER reaction: block skipped (Q13)
    s = gen.sym.stats
    assert gen.sym.naug == 338 and s["m"] + s["n0"] == 338
    assert s["nnz_L"] + s["nnz_U"] < 20000                 # vs 112 896 for the reference's dense LU


def test_table_rhs_matches_oracle(oracle, net, eng):
    g = np.load(GOLDEN / "getydot_cases.npz")
    for i in range(6):
        y, rate = g[f"y_{i}"], g[f"rate_{i}"]
        _, e = eng.ext_state(y, rate)
        ref, s_ref = oracle.getydot(rate, y, e["blr"], COV0, e["sm"], e["sb"], y[-1])
        got, s_got = eng.rhs(y, rate)
        assert np.abs(got[:335] - ref[:335]).max() <= 1e-14 * np.abs(ref).max()
        assert abs(s_got - s_ref) <= 1e-14 * abs(s_ref)


def test_gather_program_matches_csr(gen, eng):
    sym = gen.sym
    g = np.load(GOLDEN / "getydot_cases.npz")
    y, rate = g["y_1"], g["rate_1"]
    ye, _ = eng.ext_state(y, rate)
    flux = rate * np.prod(ye[sym.flux_f], axis=1)
    out = np.zeros(sym.neq)
    gen.gather.run(lambda t: -flux[t & 0x7FFF] if (t >> 15) & 1 else flux[t & 0x7FFF],
                   lambda tg, s: out.__setitem__(tg, s))
    # the two state-dependent photo reactions are not in the gather program: the RHS adds them afterwards
    assert sorted(gen.deferred) == sorted(int(net_r) for net_r in (gen.net.reaction_idx["nR_H2_hv"],
                                                                    gen.net.reaction_idx["nR_CO_hv"]))
    assert not set(gen.deferred) & set(int(r) for r in gen.flux_order)
    assert len(gen.flux_order) == gen.net.nreac - 2
    for r in gen.deferred:
        for i, sg in gen.deferred_rows[r]:
            out[i] += sg * flux[r]
    rows = np.repeat(np.arange(335), np.diff(sym.g_ptr))
    ref = np.bincount(rows, weights=flux[sym.g_reac] * sym.g_sign, minlength=335)
    assert np.abs(out[:335] - ref).max() <= 1e-14 * np.abs(ref).max()


@pytest.mark.parametrize("case,gamma", [(0, 1e3), (1, 1e3), (3, 1e3)])
def test_analytic_jacobian_and_sparse_lu(net, gen, eng, case, gamma):
    """(I - gamma J) x = b through the bordered sparse system equals a dense solve with a
    finite-difference Jacobian of the same RHS (both transfer branches)."""
    sym = gen.sym
    g = np.load(GOLDEN / "getydot_cases.npz")
    y, rate = g[f"y_{case}"].copy(), g[f"rate_{case}"]
    neq = sym.neq
    f = lambda yy: eng.rhs(yy, rate)[0]
    J = np.zeros((neq, neq))
    for j in range(neq):
        h = max(abs(y[j]) * 1e-6, 1e-30)
        yp, ym = y.copy(), y.copy()
        yp[j] += h
        ym[j] -= h
        J[:, j] = (f(yp) - f(ym)) / (2 * h)
    rng = np.random.default_rng(1)
    b = rng.standard_normal(neq) * np.abs(y)
    x_ref = np.linalg.solve(np.eye(neq) - gamma * J, b)
    ba = np.zeros(sym.naug)
    ba[:neq] = b
    ba[sym.iB] = b[sym.iB] - b[net.bulk_list].sum()
    ba[sym.iS] = b[sym.iS] - b[net.surface_list].sum()
    val = eng.assemble(y, rate, gamma)
    x = eng.solve(eng.factor(val), ba)
    assert np.abs(x[:neq] - x_ref).max() <= 1e-7 * np.abs(x_ref).max()


def test_factor_and_solve_programs(gen, eng):
    """The packed team programs reproduce the level-by-level emulator (same tables the kernels read)."""
    sym = gen.sym
    g = np.load(GOLDEN / "getydot_cases.npz")
    y, rate = g["y_1"], g["rate_1"]
    val = eng.assemble(y, rate, 1e5)
    ref = eng.factor(val)
    v = val.copy()
    v[sym.zero_slot] = 0
    for prog, dmap in gen.factor:
        new = {}

        def fin(tg, s):
            x = v[tg] - s
            d = dmap[tg]
            new[tg] = x * v[d] if d >= 0 else (1.0 / x if d == -2 else x)
        prog.run(lambda t: v[t >> 16] * v[t & 0xFFFF], fin)
        for k, x in new.items():
            v[k] = x
    m, off = sym.m, sym.off_dense
    v[off: off + m * m] = _gauss_jordan_inverse(v[off: off + m * m].reshape(m, m)).ravel()
    assert np.abs(v - ref).max() <= 1e-12 * np.abs(ref).max()
    b = np.random.default_rng(0).standard_normal(sym.naug)
    xref = eng.solve(ref, b)
    x = b[sym.perm].copy()
    for prog in list(gen.fwd) + [gen.tail]:
        new = {}
        prog.run(lambda t: v[t >> 16] * x[t & 0xFFFF], lambda tg, s: new.__setitem__(tg, x[tg] - s))
        for k, q in new.items():
            x[k] = q
    x[sym.n0:] = v[off: off + m * m].reshape(m, m) @ x[sym.n0:]
    for prog in gen.bwd:
        new = {}
        prog.run(lambda t: v[t >> 16] * x[t & 0xFFFF],
                 lambda tg, s: new.__setitem__(tg, (x[tg] - s) * v[sym.diag_pos[tg]]))
        for k, q in new.items():
            x[k] = q
    out = np.empty(sym.naug)
    out[sym.perm] = x
    assert np.abs(out - xref).max() <= 1e-10 * np.abs(xref).max()


def test_generated_header_is_current(gen, tmp_path):
    """The committed net_tables.cuh is what the generator emits for the committed network."""
    from pathlib import Path

    from uclchem_b200.makerates_cuda import emit
    p = emit(gen, tmp_path, "default")
    committed = Path(__file__).resolve().parents[1] / "uclchem_b200/csrc/generated/default/net_tables.cuh"
    assert p.read_text() == committed.read_text()


def test_network_json_matches_reference_network_f90(net):
    from pathlib import Path
    ref = Path("/root/reference/src/fortran_src/network.f90")
    if not ref.exists():
        pytest.skip("reference checkout not present (GPU box)")
    from uclchem_b200.network import Network
    fresh = Network.from_network_f90(ref)
    for k in ("alpha", "beta", "gama", "re", "pr", "mass", "binding_energy", "min_temps", "max_temps", "rtype"):
        assert np.array_equal(getattr(fresh, k), getattr(net, k)), k
    assert fresh.names == net.names


def _emit_units(name, programs, **kw):
    """Run the emitter on a list of level programs and parse the unit table back."""
    import re
    from uclchem_b200.makerates_cuda import _emit_program
    out = []
    _emit_program(out.append, name, programs, **kw)
    text = "".join(out)
    body = re.search(name + r"_units\[\d+\] = \{([^}]*)\}", text).group(1)
    u = np.array([int(x.strip().rstrip("u")) for x in body.split(",") if x.strip()], dtype=np.uint64).reshape(-1, 2)
    nterms = int(re.search(name + r"_terms\[(\d+)\]", text).group(1))
    return u, nterms, text


def test_unit_tables_sync_flags_and_padding(gen):
    """Unit tables: block barrier after every level, except a warp-level sync where this level and the
    next both fit into the 32 slots of warp 0; term tables padded for the unconditional batch loads."""
    from uclchem_b200.makerates_cuda import TERM_PAD, NTHREADS
    progs = gen.fwd + [gen.tail]
    u, nterms, text = _emit_units("net_fwd", progs, warp_chains=True)
    assert "__constant__" in text.split("net_fwd_units")[0].splitlines()[-1]
    assert nterms == sum(len(p.terms) for p in progs) + TERM_PAD and TERM_PAD >= 32 * 8
    k = 0
    for lev, p in enumerate(progs):
        npass = max(1, -(-p.nslots // NTHREADS))
        for i in range(npass):
            first, word = int(u[k, 0]), int(u[k, 1])
            end, sync = word & 0x3FFFFFFF, word >> 30
            assert end - first <= p.nslots and first % 32 == 0
            if i < npass - 1:
                assert sync == 0
            else:
                local = lev + 1 < len(progs) and p.nslots <= 32 and progs[lev + 1].nslots <= 32
                assert sync == (1 if local else 2)
            k += 1
    assert k == len(u)
    # without the option every level ends with a block barrier
    u2, _, _ = _emit_units("net_x", progs)
    assert all((int(w) >> 30) in (0, 2) for w in u2[:, 1])
    # the gather runs on the 14 worker warps of rhs_eval: passes of NTHREADS - 64 slots
    ug, _, _ = _emit_units("net_gather", [gen.gather], term16=True, nthreads=NTHREADS - 64)
    assert list(ug[:, 0]) == list(range(0, gen.gather.nslots, NTHREADS - 64))


def test_deferred_photo_reactions(gen, net):
    """H2 + hv and CO + hv: unimolecular, gas-phase rows only, absent from flux table and gather terms."""
    assert len(gen.deferred) == 2
    used = {int(t) & 0x7FFF for t in gen.gather.terms}
    for r in gen.deferred:
        assert r not in used and r not in set(int(x) for x in gen.flux_tab[:, 0] & 0xFFFF)
        rows = gen.deferred_rows[r]
        assert 1 <= len(rows) <= 4 and sum(sg for _, sg in rows) == 1   # A -> B + C
        assert all(i < net.nspec and net.names[i][0] not in "#@" for i, _ in rows)


def test_unsupported_networks_are_refused_by_the_generator(net):
    """Two-phase networks and refractory lists have reference code paths the engine does not implement
    (hotcore.f90:92-107, rates.f90:217, surfacereactions.f90:112, chemistry.f90:198): generation refuses them."""
    import copy
    from uclchem_b200.makerates_cuda import Generated
    two = copy.copy(net)
    two.three_phase = False
    with pytest.raises(NotImplementedError, match="two-phase"):
        Generated(two)
    refr = copy.copy(net)
    refr.refractory_list = np.array([200], dtype=np.int64)
    with pytest.raises(NotImplementedError, match="refractory"):
        Generated(refr)

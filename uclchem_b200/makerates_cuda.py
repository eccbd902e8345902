"""MakeRates CUDA back-end: network.f90 -> generated CUDA tables for the B200 engine.

The reference's MakeRates ends in ``io_functions.write_outputs``
(``src/uclchem/makerates/io_functions.py:231-282``), which writes ``odes.f90``
(GETYDOT) and ``network.f90``.  This module is the additional writer the
north-star asks for: from the same network it emits, once, at
network-generation time,

* the ODE right-hand side as a flux table + a gather program,
* the analytic Jacobian as an assembly program over a fixed sparse pattern,
* the fixed-sparsity symbolic LU factorisation (elimination order, fill pattern,
  level-scheduled factor program, dense trailing block, triangular-solve programs),
* the rate-coefficient tables (alpha/beta/gamma, type ranges, ice positions,
  chemical-desorption fractions, diffusion constants),

as one CUDA header ``net_tables.cuh`` of ``__device__ const`` arrays that
``uclchem_b200/csrc/uclgpu.cu`` includes.  All parallel work is expressed as
*team programs*: lists of lane slots, each slot belonging to a power-of-two team
of lanes that strides over one item's term list and reduces with warp shuffles.

Usage (the hook a maintainer adds after ``write_network_file`` in
``write_outputs``)::

    python -m uclchem_b200.makerates_cuda <network.f90> <outdir> [--tag default]
"""
from __future__ import annotations

import argparse
import hashlib
from pathlib import Path

import numpy as np

from . import product_form, symbolic
from .network import N_EXT, TYPE_ID, TYPE_NAMES, Network

NTHREADS = 512
NULL_TARGET = 0xFFFF
SOLVE_TERMS_PER_LANE = 4  # solve levels are latency chains: short per-lane chains, wider teams
TERM_PAD = 256  # >= 32 lanes x 8 terms per batch


# --------------------------------------------------------------------------
# team programs
# --------------------------------------------------------------------------
class TeamProgram:
    """items: list of (target, [term,...]); terms are ints (already packed)."""

    def __init__(self, items, terms_per_lane=12, max_team=32, term_dtype=np.uint32):
        teams = []
        for target, terms in items:
            n = len(terms)
            t = 1
            while t < max_team and t * terms_per_lane < n:
                t *= 2
            teams.append((t, n, target, terms))
        teams.sort(key=lambda x: (-x[0], -x[1]))
        begin, desc, allterms = 0, [], []
        for (t, n, target, terms) in teams:
            assert len(desc) % t == 0
            assert n < (1 << 12) and target < NULL_TARGET
            tl = t.bit_length() - 1
            for _ in range(t):
                desc.append((begin, target | (n << 16) | (tl << 28)))
            allterms.extend(terms)
            begin += n
        while len(desc) % 32:
            desc.append((0, NULL_TARGET))
        self.desc = np.asarray(desc, dtype=np.uint32).reshape(-1, 2)
        self.terms = np.asarray(allterms, dtype=term_dtype)
        self.nslots = len(desc)

    def run(self, term_value, finalize):
        """Interpreter (tests): term_value(term)->float, finalize(target, sum)."""
        d = self.desc
        for s0 in range(0, self.nslots, 32):
            lane_sum = np.zeros(32)
            info = []
            for l in range(32):
                begin, w = int(d[s0 + l, 0]), int(d[s0 + l, 1])
                target, n, tl = w & 0xFFFF, (w >> 16) & 0xFFF, (w >> 28) & 0x7
                t = 1 << tl
                info.append((target, t))
                if target == NULL_TARGET:
                    continue
                lane = (s0 + l) & (t - 1)
                acc = 0.0
                for q in range(lane, n, t):
                    acc += term_value(int(self.terms[begin + q]))
                lane_sum[l] = acc
            for o in (1, 2, 4, 8, 16):
                nxt = lane_sum.copy()
                for l in range(32):
                    if o < info[l][1]:
                        nxt[l] = lane_sum[l] + lane_sum[l ^ o]
                lane_sum = nxt
            for l in range(32):
                target, t = info[l]
                if target != NULL_TARGET and ((s0 + l) & (t - 1)) == 0:
                    finalize(target, lane_sum[l])


class Generated:
    """All tables of one network, in memory (also used by the CPU tests)."""

    def __init__(self, net: Network, dense_threshold: float = 0.9, factor_terms_per_lane: int = 8):
        # refused, not ignored: the engine implements the three-phase paths only -- the
        # two-phase branches of hotcore.f90:92-107 (instant sublimation / thermal evaporation), rates.f90:217 and
        # surfacereactions.f90:112 are not built -- and not the refractory subtraction of chemistry.f90:198
        if not net.three_phase:
            raise NotImplementedError("two-phase networks (THREE_PHASE = .False.) are not supported by the CUDA back-end")
        if len(net.refractory_list):
            raise NotImplementedError("networks with refractory species (refractoryList) are not supported by the CUDA "
                                      "back-end: safeBulk of chemistry.f90:198 is not implemented on the device")
        self.net = net
        self.sym = sym = symbolic.build(net, dense_threshold)
        neq = sym.neq
        # ydot gather: term = reaction | neg<<15
        assert net.nreac < (1 << 15), "gather term packing needs nreac < 32768"
        # Deferred reactions: the two photo rates that follow the state (H2 self-shielding, CO
        # shielding) are long single-lane chains.  Two warps of the CTA evaluate them while the other
        # warps do the fluxes and the gather, so their fluxes are not part of the flux table / gather
        # program: the RHS adds them to ydot afterwards (deferred_spec / deferred_sign per reaction).
        self.deferred = [int(net.reaction_idx["nR_H2_hv"]), int(net.reaction_idx["nR_CO_hv"])]
        self.deferred_rows = {r: [] for r in self.deferred}
        items = []
        for i in range(net.nspec):
            a, b = sym.g_ptr[i], sym.g_ptr[i + 1]
            if i in (sym.iB, sym.iS):
                continue  # BULK / SURFACE totals are written by the transfer phase of the RHS
            terms = []
            for r, sg in zip(sym.g_reac[a:b], sym.g_sign[a:b]):
                if int(r) in self.deferred_rows:
                    self.deferred_rows[int(r)].append((i, int(sg)))
                    continue
                t = int(r) | ((1 << 15) if sg < 0 else 0)
                terms.append(t)
            items.append((i, terms))
        for r in self.deferred:
            # one plain reactant, no ext factors: flux = rate * y[reactant]
            assert (sym.flux_f[r][1:] >= neq).all() and sym.flux_f[r][0] < net.nspec, "deferred reactions are unimolecular"
            assert all(i < net.nspec and i not in net.surface_list and i not in net.bulk_list for i, _ in self.deferred_rows[r])
        self.gather = TeamProgram(items)
        self.flux_f = sym.flux_f.astype(np.int16)
        # reactions whose flux needs an ext factor (blr, 1/safeMantle, tau) or a per-RHS photo rate
        ext = (sym.flux_f > neq).any(axis=1)
        keep = np.ones(net.nreac, bool)
        keep[self.deferred] = False
        self.flux_order = np.concatenate([np.where(~ext & keep)[0], np.where(ext & keep)[0]]).astype(np.int32)
        self.n_plain = int((~ext & keep).sum())
        ff = np.full((net.nreac, 4), neq + 0, np.int64)
        ff[:, : sym.fwidth] = sym.flux_f
        tab = np.zeros((len(self.flux_order), 2), np.uint32)
        for idx, r in enumerate(self.flux_order):
            f = ff[r]
            tab[idx, 0] = int(r) | (int(f[0]) << 16)
            tab[idx, 1] = int(f[1]) | (int(f[2]) << 10) | (int(f[3]) << 20)
        self.flux_tab = tab
        # Jacobian assembly: one item per assembled entry; term = packed jterm; gamma flag folded in target list
        assert sym.fwidth <= 4 and neq + N_EXT < 1024 and net.nreac < (1 << 16), "packed table limits"
        one = neq + 0
        items = []
        r_, k_, kind_, neg_ = symbolic.unpack_jterm(sym.j_term)
        packed = []
        for r, k, kind, neg in zip(r_, k_, kind_, neg_):
            others = [int(f) for kk, f in enumerate(sym.flux_f[r]) if kk != k]
            others += [one] * (3 - len(others))
            x = int(r) | (int(kind) << 28) | (int(neg) << 31)
            y = others[0] | (others[1] << 10) | (others[2] << 20)
            packed.append((x << 32) | y)   # 64-bit term: high word x, low word y
        for e in range(len(sym.j_pos)):
            a, b = sym.j_ptr[e], sym.j_ptr[e + 1]
            items.append((int(sym.j_pos[e]), packed[a:b]))
        self.jac = TeamProgram(items, terms_per_lane=6, term_dtype=np.uint64)
        self.aux_row_pos = sym.j_pos[sym.j_gamma == 0].astype(np.int32)  # entries scaled by -1 instead of -gamma
        # factor levels
        self.factor = []
        for L in sym.f_levels:
            items = []
            for e in range(len(L["target"])):
                a, b = L["ptr"][e], L["ptr"][e + 1]
                items.append((int(L["target"][e]), [(int(l) << 16) | int(u) for l, u in zip(L["tl"][a:b], L["tu"][a:b])]))
            self.factor.append((TeamProgram(items, terms_per_lane=factor_terms_per_lane), {int(t): int(d) for t, d in zip(L["target"], L["diag"])}))
        # solves
        def lvl(Ls):
            out = []
            for L in Ls:
                items = []
                for e, n in enumerate(L["rows"]):
                    a, b = L["ptr"][e], L["ptr"][e + 1]
                    items.append((int(n), [(int(p) << 16) | int(c) for p, c in zip(L["pos"][a:b], L["cols"][a:b])]))
                out.append(TeamProgram(items, terms_per_lane=SOLVE_TERMS_PER_LANE))
            return out
        self.fwd = lvl(sym.fwd_levels)
        self.bwd = lvl(sym.bwd_levels)
        items = []
        for t in range(sym.m):
            a, b = sym.tail_l_ptr[t], sym.tail_l_ptr[t + 1]
            items.append((sym.n0 + t, [(int(p) << 16) | int(c) for p, c in zip(sym.tail_l_pos[a:b], sym.tail_l_col[a:b])]))
        self.tail = TeamProgram(items, terms_per_lane=8)
        # product-form solves (product_form.py): inverse program + the three new solve levels
        self.pf = pf = product_form.build(sym)
        pk = lambda terms: [(int(p) << 16) | int(k) for p, k in terms]
        self.pf_inv = [TeamProgram([(t, pk(terms)) for t, _, terms in lv], terms_per_lane=SOLVE_TERMS_PER_LANE)
                       for lv in pf.inv_levels]
        scale = np.full(pf.nstg, 0xFFFF, np.int64)
        for lv in pf.inv_levels:
            for t, sc, _ in lv:
                scale[t] = sc
        self.pf_inv_scale = scale
        self.pf_p1 = TeamProgram([(i, pk(t)) for i, t in pf.p1], terms_per_lane=SOLVE_TERMS_PER_LANE)
        self.pf_p4 = TeamProgram([(i, pk(t)) for i, t in pf.p4], terms_per_lane=SOLVE_TERMS_PER_LANE)
        self.pf_p5 = TeamProgram([(i, pk(t)) for i, t in pf.p5], terms_per_lane=SOLVE_TERMS_PER_LANE)
        self.pf_tail = TeamProgram(items, terms_per_lane=SOLVE_TERMS_PER_LANE)

    # scaling map for factor targets: diag position or -1
    def factor_diag_table(self):
        """per storage position: pivot position to scale an L entry with, 0xFFFE for a sparse
        pivot (stored as its reciprocal), 0xFFFF for plain entries."""
        sym = self.sym
        tab = np.full(sym.nval, 0xFFFF, np.int64)
        for _, dmap in self.factor:
            for t, d in dmap.items():
                tab[t] = d if d >= 0 else (0xFFFE if d == -2 else 0xFFFF)
        return tab

    def position_codes(self):
        """per storage position: bits 0-1 = base value of P (0: 0, 1: +1, 2: -1), bit 2 = row of an
        auxiliary unknown (Jacobian scaled by 1 instead of gamma)."""
        sym = self.sym
        code = np.zeros(sym.nval, np.uint8)
        code[sym.diag_pos] = 1
        code[sym.const_pos] = 2
        aux_new = {int(sym.iperm[sym.iSg]), int(sym.iperm[sym.iTau])}
        for (i, j), pos in sym.ent_pos.items():
            if i in aux_new:
                code[pos] |= 4
        return code


# --------------------------------------------------------------------------
# rate-table precomputation (state independent pieces of rates.f90 / surfacereactions.f90)
# --------------------------------------------------------------------------
def _f32(x):
    return float(np.float32(x))


def rate_tables(net: Network) -> dict:
    """Per-reaction integer/float side tables used by the rates kernel."""
    K_BOLTZ, AMU, RP = 1.38065040e-16, 1.66053892e-24, 1.054571628e-27
    PI = _f32(3.141592654)
    nreac = net.nreac
    ice_pos = {int(s): i for i, s in enumerate(net.ice_list)}
    gas_pos = {}
    for i, s in enumerate(net.gas_ice_list):
        gas_pos[int(s)] = i  # later (bulk) entries win, like the reference's search loop
    is_bulk = np.zeros(net.nspec + 1, bool)
    is_bulk[net.bulk_list] = True
    is_surf = np.zeros(net.nspec + 1, bool)
    is_surf[net.surface_list] = True
    # vdiff, chemistry.f90:108-112
    vdiff_pref = 2.0 * K_BOLTZ * 1.5e15 / PI / PI / AMU
    vdiff = np.sqrt(vdiff_pref * net.binding_energy / net.mass[net.ice_list])
    ia = np.full(nreac, -1, np.int32)   # ice position of reactant 1
    ib = np.full(nreac, -1, np.int32)   # ice position of reactant 2
    phase = np.zeros(nreac, np.int8)    # 0 gas, 1 surface, 2 bulk (reactant 1)
    desfrac = np.zeros(nreac)
    tunnel = np.zeros(nreac)
    partner = np.full(nreac, -1, np.int32)
    mass1 = np.ones(nreac)
    for r in range(nreac):
        r1, r2 = int(net.re[r, 0]), int(net.re[r, 1])
        ia[r] = ice_pos.get(r1, -1)
        ib[r] = ice_pos.get(r2, -1)
        phase[r] = 2 if is_bulk[r1] else (1 if is_surf[r1] else 0)
        mass1[r] = net.mass[r1] if r1 < net.nspec else 1.0
    from .table_emulator import GAS_DUST_DENSITY_RATIO, NUM_SITES_PER_GRAIN  # same constants

    def desorption_fraction(r):
        # surfacereactions.f90:218-297 incl. the index-space mix-up
        re, pr = net.re[r], net.pr[r]
        r1 = r2 = -1
        prod = [-1] * 4
        for i in range(len(net.ice_list)):
            ice, gas = int(net.ice_list[i]), int(net.gas_ice_list[i])
            if ice == re[0] or gas == re[0]:
                r1 = i
            if ice == re[1] or gas == re[1]:
                r2 = i
            for k in range(4):
                if pr[k] >= 0 and (ice == pr[k]):
                    prod[k] = i
            for k in range(4):
                if pr[k] >= 0 and (gas == pr[k]):
                    prod[k] = i
        max_be = prod_enth = eps = 0.0
        for k in range(4):
            if prod[k] >= 0:
                max_be = max(max_be, net.binding_energy[prod[k]])
                prod_enth = prod_enth + net.formation_enthalpy[prod[k]]
                eps = eps + net.mass[prod[k]]
        q = (eps - 120.0) / (eps + 120.0)
        eps = q * q
        dh = net.formation_enthalpy[r1] + net.formation_enthalpy[r2] - prod_enth
        dh = dh * 4.184e03 / (1.38054e-23 * 6.02214129e23)
        dh = dh + net.gama[r]
        if dh == 0.0:
            dh = _f32(1e-30)
        dof = net.atom_counts[prod[0]]
        for k in range(1, 4):
            if prod[k] >= 0:
                dof = max(dof, net.atom_counts[prod[k]])
        dof = 3 * int(dof)
        with np.errstate(over="ignore", divide="ignore"):
            frac = float(np.exp((-max_be * float(dof)) / (eps * dh))) if eps * dh != 0 else 0.0
        if dh < 0.0:
            frac = 0.0
        frac = frac / 10
        ngn, ngo, ngoh, nh = (net.species_idx[k] for k in ("ngn", "ngo", "ngoh", "nh"))
        if re[0] == ngn and re[1] == ngn:
            frac = _f32(0.5)
        if (re[0] == ngo and re[1] == nh) or (re[0] == nh and re[1] == ngo):
            frac = _f32(0.3)
        if (re[0] == ngoh and re[1] == nh) or (re[0] == nh and re[1] == ngoh):
            frac = _f32(0.25)
        return frac

    for tname, dname in (("LH", "LHDES"), ("ER", "ERDES")):
        rng, drng = net.type_ranges[tname], net.type_ranges[dname]
        if rng is None:
            continue
        for k in range(rng[1] - rng[0] + 1):
            r, d = rng[0] + k, drng[0] + k
            partner[r], partner[d] = d, r
            desfrac[d] = desorption_fraction(d)
            desfrac[r] = desfrac[d]
    rng = net.type_ranges["LH"]
    if rng is not None:
        for r in list(range(rng[0], rng[1] + 1)) + list(range(net.type_ranges["LHDES"][0], net.type_ranges["LHDES"][1] + 1)):
            rm = net.reduced_masses[r]
            if rm == 0.0:
                m1, m2 = net.mass[net.re[r, 0]], net.mass[net.re[r, 1]]
                rm = m1 * m2 / (m1 + m2)
            tunnel[r] = 2.0 * 1.40e-8 / RP * np.sqrt(2.0 * AMU * rm * K_BOLTZ * net.gama[r])
    fpart = np.full(nreac, -1, np.int32)  # for a FREEZE reaction: index k of its desorption partner row, else -1
    for k, fp in enumerate(net.freeze_partners):
        fpart[int(fp)] = k
    return dict(vdiff=vdiff, ia=ia, ib=ib, phase=phase, desfrac=desfrac, tunnel=tunnel, partner=partner,
                mass1=mass1, fpart=fpart, gdr=GAS_DUST_DENSITY_RATIO, nsites=NUM_SITES_PER_GRAIN)


def photo_tables() -> dict:
    """Second-derivative tables of the two NR splines in photoreactions.f90 (:225-271),
    including the (8,6)->(7,6) re-stride of the CO shielding table."""
    lam = np.array([910.0, 950.0, 1000.0, 1050.0, 1110.0, 1180.0, 1250.0, 1390.0, 1490.0, 1600.0, 1700.0,
                    1800.0, 1900.0, 2000.0, 2100.0, 2190.0, 2300.0, 2400.0, 2500.0, 2740.0, 3440.0, 4000.0,
                    4400.0, 5500.0, 7000.0, 9000.0, 12500.0, 22000.0, 34000.0, 1.0e9])
    xl = np.array([5.76, 5.18, 4.65, 4.16, 3.73, 3.40, 3.11, 2.74, 2.63, 2.62, 2.54, 2.50, 2.58, 2.78, 3.01,
                   3.12, 2.86, 2.58, 2.35, 2.00, 1.58, 1.42, 1.32, 1.00, 0.75, 0.48, 0.28, 0.12, 0.05, 0.00])
    sco = np.array([
        0.000e+00, -1.408e-02, -1.099e-01, -4.400e-01, -1.154e+00, -1.888e+00, -2.760e+00, -4.001e+00,
        -8.539e-02, -1.015e-01, -2.104e-01, -5.608e-01, -1.272e+00, -1.973e+00, -2.818e+00, -4.055e+00,
        -1.451e-01, -1.612e-01, -2.708e-01, -6.273e-01, -1.355e+00, -2.057e+00, -2.902e+00, -4.122e+00,
        -4.559e-01, -4.666e-01, -5.432e-01, -8.665e-01, -1.602e+00, -2.303e+00, -3.146e+00, -4.421e+00,
        -1.303e+00, -1.312e+00, -1.367e+00, -1.676e+00, -2.305e+00, -3.034e+00, -3.758e+00, -5.077e+00,
        -3.883e+00, -3.888e+00, -3.936e+00, -4.197e+00, -4.739e+00, -5.165e+00, -5.441e+00, -6.446e+00])
    nh2 = np.array([18.0, 19.0, 20.0, 21.0, 22.0, 23.0])

    def spline(x, y):
        n = len(x)
        y2 = np.zeros(n)
        u = np.zeros(n)
        for i in range(1, n - 1):
            sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1])
            p = sig * y2[i - 1] + 2.0
            y2[i] = (sig - 1.0) / p
            u[i] = (6.0 * ((y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1])) /
                    (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p
        y2[n - 1] = (0.0 - 0.0 * u[n - 2]) / (0.0 * y2[n - 2] + 1.0)
        for k in range(n - 2, -1, -1):
            y2[k] = y2[k] * y2[k + 1] + u[k]
        return y2

    m, n = 7, 6
    sco_rows = np.zeros((m, n))
    sco_d2 = np.zeros((m, n))
    for j in range(m):
        sco_rows[j] = [sco[j + k * m] for k in range(n)]
        sco_d2[j] = spline(nh2, sco_rows[j])
    return dict(lambda_grid=lam, xlambda_grid=xl, xlambda_d2=spline(lam, xl), sco_rows=sco_rows, sco_d2=sco_d2)


# --------------------------------------------------------------------------
# emission
# --------------------------------------------------------------------------
def _c_array(name, arr, ctype, per_line=16, qual="__device__"):
    arr = np.asarray(arr).ravel()
    if ctype == "double":
        body = [repr(float(v)) if np.isfinite(v) else "0.0" for v in arr]
    else:
        body = [str(int(v)) for v in arr]
        if ctype.startswith("uint"):
            body = [b + "u" for b in body]
    lines = [", ".join(body[i:i + per_line]) for i in range(0, len(body), per_line)]
    n = max(1, len(arr))
    if len(arr) == 0:
        lines = ["0"]
    return f"{qual} __align__(16) const {ctype} {name}[{n}] = {{\n  " + ",\n  ".join(lines) + "\n};\n"


def emit(gen: Generated, outdir: Path, tag: str) -> Path:
    net, sym = gen.net, gen.sym
    outdir.mkdir(parents=True, exist_ok=True)
    rt = rate_tables(net)
    ph = photo_tables()
    o = []
    w = o.append
    w("// GENERATED by uclchem_b200/makerates_cuda.py -- do not edit.\n")
    w(f"// network tag: {tag}; {net.nspec} species, {net.nreac} reactions\n")
    w("#pragma once\n#include <stdint.h>\n\n")
    d = {
        "NET_NSPEC": net.nspec, "NET_NEQ": sym.neq, "NET_NAUG": sym.naug, "NET_NREAC": net.nreac,
        "NET_NICE": len(net.ice_list), "NET_NSURF": len(net.surface_list), "NET_FWIDTH": sym.fwidth,
        "NET_NVAL": sym.nval, "NET_N0": sym.n0, "NET_M": sym.m, "NET_OFF_DENSE": sym.off_dense,
        "NET_ZERO_SLOT": sym.zero_slot, "NET_IB": sym.iB, "NET_IS": sym.iS, "NET_ID": sym.iD,
        "NET_ISG": sym.iSg, "NET_ITAU": sym.iTau, "NET_NTHREADS": NTHREADS,
        "NET_TAU_POS_B": sym.tau_pos_B, "NET_TAU_POS_S": sym.tau_pos_S, "NET_DD_POS": sym.dd_pos,
        "NET_NSWAP": len(sym.swap_reacs), "NET_SWAP_LO": int(sym.swap_reacs[0]) if len(sym.swap_reacs) else 0,
        "NET_NGAR": net.gar_params.shape[0] if net.type_ranges["GAR"] else 0,
        "NET_THREE_PHASE": 1 if net.three_phase else 0,
    }
    for k, v in d.items():
        w(f"#define {k} {v}\n")
    w(f'#define NET_TAG "{tag}"\n')
    # algorithmic work per operation, counted from the generated tables
    st_ = sym.stats
    n_gather = int(len(sym.g_reac))
    f_rhs = net.nreac * sym.fwidth + n_gather + 6 * len(net.surface_list)
    f_jac = 4 * st_["j_terms"]
    # LU: the sparse factor terms + an LU of the dense trailing block (2/3 m^3).  The engine inverts that block
    # explicitly (Gauss-Jordan, 2 m^3) to make the solves wide: that is an implementation choice, so it is
    # reported separately (NET_FLOP_LU_EXEC) and never enters the roofline figure.
    f_lu = 2 * st_["factor_terms"] + (2 * sym.m ** 3) // 3
    f_lu_exec = 2 * st_["factor_terms"] + 2 * sym.m ** 3
    f_solve = 2 * (st_["fwd_terms"] + st_["bwd_terms"] + st_["tail_terms"]) + 2 * sym.m ** 2
    w(f"#define NET_FLOP_RHS {float(f_rhs)}\n#define NET_FLOP_JAC {float(f_jac)}\n#define NET_FLOP_LU {float(f_lu)}\n")
    w(f"#define NET_FLOP_SOLVE {float(f_solve)}\n#define NET_FLOP_RATES {float(40 * net.nreac)}\n")
    w(f"#define NET_FLOP_LU_EXEC {float(f_lu_exec)}\n")
    # algorithmic HBM bytes per cell-model with the state chip-resident: parameters in, result row out
    # (y_final, physics, flag, counters); trajectories, when requested, add (8 + nspec) * 8 per output row
    w(f"#define NET_BYTES_CELL {float(64 * 8 + sym.neq * 8 + 8 * 8 + 4 + 20 * 8)}\n")
    w(f"#define NET_BYTES_INTERVAL {float(2 * sym.neq * 8 + 64)}\n")
    for t in TYPE_NAMES:
        r = net.type_ranges[t]
        w(f"#define NET_{t}_LO {r[0] if r else -1}\n#define NET_{t}_HI {r[1] if r else -2}\n")
    for k, v in net.species_idx.items():
        w(f"#define NET_{k.upper()} {v}\n")
    for k, v in net.reaction_idx.items():
        w(f"#define NET_{k.upper()} {v}\n")
    w(f"#define NET_GDR {rt['gdr']!r}\n#define NET_NSITES {rt['nsites']!r}\n")
    w("\n// ---- species / reaction tables -------------------------------------------------\n")
    w("static const char *const net_species_names[] = {" + ", ".join('"%s"' % n for n in net.names) + "};\n")
    w(_c_array("net_mass", net.mass, "double"))
    w(_c_array("net_surface_list", net.surface_list, "int16_t"))
    w(_c_array("net_bulk_list", net.bulk_list, "int16_t"))
    w(_c_array("net_ice_list", net.ice_list, "int16_t"))
    w(_c_array("net_gas_ice_list", net.gas_ice_list, "int16_t"))
    w(_c_array("net_is_refractory_ice", [1 if int(s) in set(net.refractory_list.tolist()) else 0 for s in net.ice_list], "uint8_t"))
    w(_c_array("net_binding_energy", net.binding_energy, "double"))
    w(_c_array("net_vdiff", rt["vdiff"], "double"))
    w(_c_array("net_is_ion", [1 if "+" in n else 0 for n in net.names], "uint8_t"))
    w(_c_array("net_rtype", net.rtype, "uint8_t"))
    w(_c_array("net_re1", net.re[:, 0], "int16_t"))
    for k in ("alpha", "beta", "gama", "min_temps", "max_temps"):
        w(_c_array("net_" + k, getattr(net, k), "double"))
    w(_c_array("net_extrapolate", net.extrapolate.astype(int), "uint8_t"))
    w(_c_array("net_ia", rt["ia"], "int16_t"))
    w(_c_array("net_ib", rt["ib"], "int16_t"))
    w(_c_array("net_phase", rt["phase"], "uint8_t"))
    w(_c_array("net_desfrac", rt["desfrac"], "double"))
    w(_c_array("net_tunnel", rt["tunnel"], "double"))
    w(_c_array("net_partner", rt["partner"], "int16_t"))
    w(_c_array("net_mass1", rt["mass1"], "double"))
    w(_c_array("net_freeze_partners", net.freeze_partners, "int16_t"))
    w(_c_array("net_gar_params", net.gar_params, "double"))
    # photo-rate tables are walked by a single lane per RHS evaluation (a latency chain): constant memory
    w(_c_array("net_lambda_grid", ph["lambda_grid"], "double", qual="__constant__"))
    w(_c_array("net_xlambda_grid", ph["xlambda_grid"], "double", qual="__constant__"))
    w(_c_array("net_xlambda_d2", ph["xlambda_d2"], "double", qual="__constant__"))
    w(_c_array("net_sco_rows", ph["sco_rows"], "double", qual="__constant__"))
    w(_c_array("net_sco_d2", ph["sco_d2"], "double", qual="__constant__"))
    w("\n// ---- RHS: flux table + gather program ---------------------------------------------\n")
    w(f"#define NET_NPLAIN {gen.n_plain}\n#define NET_NFLUX {len(gen.flux_order)}\n")
    w(_c_array("net_flux_tab", gen.flux_tab, "uint32_t"))
    # deferred photo reactions: {species, sign} rows, 4 slots per reaction (H2 + hv, CO + hv), -1 padded
    dspec, dsign = [], []
    for r in gen.deferred:
        rows = gen.deferred_rows[r]
        assert len(rows) <= 4
        dspec += [i for i, _ in rows] + [-1] * (4 - len(rows))
        dsign += [sg for _, sg in rows] + [0] * (4 - len(rows))
    w(_c_array("net_deferred_spec", dspec, "int16_t"))
    w(_c_array("net_deferred_sign", dsign, "int8_t"))
    w(f"#define NET_DEFERRED_RE_H2 {int(sym.flux_f[gen.deferred[0]][0])}\n#define NET_DEFERRED_RE_CO {int(sym.flux_f[gen.deferred[1]][0])}\n")
    # the last two warps of the CTA evaluate the deferred photo rates during the gather
    _emit_program(w, "net_gather", [gen.gather], term16=True, nthreads=NTHREADS - 64)
    w("\n// ---- analytic Jacobian assembly ---------------------------------------------------\n")
    _emit_program(w, "net_jac", [gen.jac], term64=True)
    w(_c_array("net_diag_pos", sym.diag_pos, "uint16_t"))
    w(_c_array("net_pos_code", gen.position_codes(), "uint8_t"))
    w(_c_array("net_tr_pos", sym.tr_pos, "uint16_t"))
    w(_c_array("net_tau_pos_b", sym.tau_pos_b, "uint16_t"))
    w("\n// ---- symbolic LU: factor levels, dense block, solves ---------------------------------\n")
    _emit_program(w, "net_factor", [p for p, _ in gen.factor])
    w(_c_array("net_factor_diag", gen.factor_diag_table(), "uint16_t"))
    _emit_program(w, "net_fwd", gen.fwd + [gen.tail], warp_chains=True)  # forward levels, then b_T -= L21 x
    _emit_program(w, "net_bwd", gen.bwd, warp_chains=True)
    w("\n// ---- product-form solves (product_form.py; device side behind -DUCLGPU_PRODUCT_FORM) ----------\n")
    pf = gen.pf
    w(f"#define NET_NVAL_PF {pf.nval_pf}\n#define NET_PF_NX {pf.nx}\n#define NET_PF_NY {pf.ny}\n")
    f_solve_pf = 2 * (pf.stats["p1_terms"] + st_["tail_terms"] + pf.stats["p4_terms"] + pf.stats["p5_terms"]) + 2 * sym.m ** 2
    f_lu_pf = f_lu_exec + 2 * pf.stats["inv_terms"]
    w(f"#define NET_FLOP_SOLVE_PF {float(f_solve_pf)}\n#define NET_FLOP_LU_PF {float(f_lu_pf)}\n")
    w(f"#define NET_PF_DIAG0 {pf.stg_diag0}\n#define NET_PF_ONE {pf.stg_one}\n#define NET_PF_NSTG {pf.nstg}\n")
    # (the product-form build static_asserts NET_PF_NSTG <= NREAC: its staging buffer is the flux array)
    _emit_program(w, "net_pf_inv", gen.pf_inv)
    w(_c_array("net_pf_inv_scale", gen.pf_inv_scale, "uint16_t"))
    # the same program in passes of PF_SIDE_THREADS slots: run by the warps that idle during the dense
    # Gauss-Jordan inverse (-DUCLGPU_PF_OVERLAP)
    side = NTHREADS - (((-(-sym.m // 5)) ** 2 + 31) & ~31)
    w(f"#define NET_PF_SIDE_THREADS {side}\n")
    _emit_program(w, "net_pf_side", gen.pf_inv, nthreads=side)
    w(_c_array("net_pf_final_pos", pf.final_pos, "uint16_t"))
    _emit_program(w, "net_pf_p1", [gen.pf_p1])
    _emit_program(w, "net_pf_tail", [gen.pf_tail])
    _emit_program(w, "net_pf_p4", [gen.pf_p4])
    _emit_program(w, "net_pf_p5", [gen.pf_p5])
    w(_c_array("net_perm", sym.perm, "uint16_t"))
    w(_c_array("net_iperm", sym.iperm, "uint16_t"))
    path = outdir / "net_tables.cuh"
    text = "".join(o)
    path.write_text(text)
    (outdir / "net_tables.sha256").write_text(hashlib.sha256(text.encode()).hexdigest() + "\n")
    return path


def _emit_program(w, name, programs, term16=False, term64=False, warp_chains=False, nthreads=NTHREADS):
    """Concatenate level programs: desc (uint2 per slot), terms, level slot ranges."""
    desc, terms, lv = [], [], [0]
    for p in programs:
        d = p.desc.copy()
        d[:, 0] += len(terms)
        desc.append(d)
        terms.extend(p.terms.tolist())
        lv.append(lv[-1] + p.nslots)
    desc = np.concatenate(desc) if desc else np.zeros((0, 2), np.uint32)
    # the device executor loads term batches unconditionally (lane + u*team, u < 8) and masks the
    # sum: pad so that the last row's batch stays inside the table (index 0 is valid everywhere)
    terms.extend([0] * TERM_PAD)
    w(f"#define {name.upper()}_NLEVELS {len(programs)}\n")
    w(_c_array(name + "_desc", desc, "uint32_t"))
    if term16:
        assert max(terms) < 65536
        w(_c_array(name + "_terms", terms, "uint16_t"))
    elif term64:
        # uint2 {y = low word (other factors), x = high word (reaction|kind|sign)} as two uint32
        flat = []
        for t in terms:
            flat.append(int(t) >> 32)
            flat.append(int(t) & 0xFFFFFFFF)
        w(_c_array(name + "_terms", flat, "uint32_t"))
    else:
        w(_c_array(name + "_terms", terms, "uint32_t"))
    w(_c_array(name + "_levels", lv, "uint32_t"))
    # unit table: one entry per (level, pass of NTHREADS slots): {first slot, end slot | barrier-after flag << 31}
    # sync after a unit: bit 31 = block barrier; bit 30 = warp-level sync only -- this level and the next
    # both fit in the 32 slots of warp 0, so the chain of tiny levels at the tail of a triangular solve
    # runs inside one warp while the other warps wait at the next block barrier
    units = []
    for k in range(len(programs)):
        b0, e0 = lv[k], lv[k + 1]
        starts = list(range(b0, e0, nthreads)) or [b0]
        warp_local = (warp_chains and k + 1 < len(programs) and e0 - b0 <= 32 and lv[k + 2] - lv[k + 1] <= 32)
        for i, st in enumerate(starts):
            last = i == len(starts) - 1
            units.append((st, e0 | (((1 << 30) if warp_local else (1 << 31)) if last else 0)))
    w(f"#define {name.upper()}_NUNITS {len(units)}\n")
    # unit tables are read with a block-uniform index: constant memory (no L2 round trip per level)
    w(_c_array(name + "_units", np.asarray(units, dtype=np.uint64).ravel(), "uint32_t", qual="__constant__"))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("network_f90", type=Path)
    ap.add_argument("outdir", type=Path)
    ap.add_argument("--tag", default="default")
    ap.add_argument("--json", type=Path, default=None, help="also save the parsed network as JSON")
    a = ap.parse_args(argv)
    net = Network.from_network_f90(a.network_f90)
    if a.json:
        net.to_json(a.json)
    gen = Generated(net)
    p = emit(gen, a.outdir, a.tag)
    print(f"wrote {p} ({p.stat().st_size / 1e6:.2f} MB); symbolic stats: {gen.sym.stats}")


if __name__ == "__main__":
    main()

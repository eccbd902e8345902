"""Numpy emulator of the generated table programs (host-side validation only).

Executes the flux / gather / Jacobian-assembly / factor / solve programs of
:class:`uclchem_b200.symbolic.Symbolic` exactly as the CUDA kernels do (same
tables, same formulas, sequential instead of parallel).  The CPU tests compare
it with the oracle's RHS and with dense linear algebra, so generator bugs are
caught without a GPU.  It is NOT a fallback: nothing in the product path calls it.
"""
from __future__ import annotations

import numpy as np

from .network import EXT_BLR, EXT_INV_SM, EXT_ONE, EXT_SWAP_SM, N_EXT
from .symbolic import KIND_BLR, KIND_ISM, Symbolic, unpack_jterm

# surfacereactions.f90:34-50 constants (PI is a single-precision literal in constants.f90:9)
PI_F = float(np.float32(3.141592654))
AMU = 1.66053892e-24
GRAIN_RADIUS = 1.0e-5
GAS_DUST_DENSITY_RATIO = (4.0 * PI_F * (GRAIN_RADIUS * GRAIN_RADIUS * GRAIN_RADIUS) * 3.0 * 100.0) / (3.0 * AMU)
NUM_SITES_PER_GRAIN = GRAIN_RADIUS * GRAIN_RADIUS * 1.5e15 * 4.0 * PI_F
COV0 = 0.5 * GAS_DUST_DENSITY_RATIO / NUM_SITES_PER_GRAIN


class TableEngine:
    def __init__(self, sym: Symbolic):
        self.s = sym

    # -- extended state -----------------------------------------------------
    def ext_state(self, y, rate):
        s = self.s
        neq = s.neq
        ye = np.empty(neq + N_EXT)
        ye[:neq] = y
        sm = max(1e-30, y[s.iS])
        sb = max(1e-30, y[s.iB])
        blr = min(1.0, NUM_SITES_PER_GRAIN / (GAS_DUST_DENSITY_RATIO * sb))
        ism = 1.0 / sm
        total_swap = float(np.sum(rate[s.swap_reacs] * y[s.net.re[s.swap_reacs, 0]] * blr))
        ye[neq + EXT_ONE] = 1.0
        ye[neq + EXT_BLR] = blr
        ye[neq + EXT_INV_SM] = ism
        ye[neq + EXT_SWAP_SM] = total_swap * ism
        dblr = 0.0 if (blr >= 1.0 or y[s.iB] <= 1e-30) else -blr / sb
        dism = 0.0 if y[s.iS] <= 1e-30 else -ism * ism
        return ye, dict(sm=sm, sb=sb, blr=blr, ism=ism, tsw=total_swap * ism, dblr=dblr, dism=dism)

    # -- RHS ----------------------------------------------------------------
    def rhs(self, y, rate, densdot=0.0):
        s = self.s
        net = s.net
        ye, e = self.ext_state(y, rate)
        flux = rate * np.prod(ye[s.flux_f], axis=1)
        ydot = np.zeros(s.neq)
        contrib = flux[s.g_reac] * s.g_sign
        rows = np.repeat(np.arange(net.nspec), np.diff(s.g_ptr))
        ydot[: net.nspec] = np.bincount(rows, weights=contrib, minlength=net.nspec)
        surf, bulk = net.surface_list, net.bulk_list
        S = ydot[surf].sum()
        mb = ydot[bulk].sum()
        if S < 0:
            q = min(1.0, e["sb"] / e["sm"]) / e["sb"]
            c = S * q * y[bulk]
            ydot[surf] -= c
            ydot[bulk] += c
        else:
            c = S * COV0 * y[surf]
            ydot[surf] -= c
            ydot[bulk] += c
        ydot[s.iB] = mb + c.sum()
        ydot[s.iS] = S - c.sum()
        ydot[s.iD] = densdot
        return ydot, S

    # -- Newton matrix P = I - gamma*J in generated storage ---------------------
    def assemble(self, y, rate, gamma, ddensdot=0.0):
        s = self.s
        net = s.net
        ye, e = self.ext_state(y, rate)
        val = np.zeros(s.nval)
        val[s.diag_pos] = 1.0
        val[s.const_pos] = s.const_val
        F = s.flux_f
        r, k, kind, neg = unpack_jterm(s.j_term)
        fac = ye[F[r]]                                  # [nterm, fwidth]
        mask = np.ones_like(fac, dtype=bool)
        mask[np.arange(len(r)), k] = False
        part = rate[r] * np.prod(np.where(mask, fac, 1.0), axis=1)
        part = np.where(kind == KIND_BLR, part * e["dblr"], part)
        part = np.where(kind == KIND_ISM, part * e["dism"], part)
        part = np.where(neg == 1, -part, part)
        sums = np.add.reduceat(part, s.j_ptr[:-1])
        scale = np.where(s.j_gamma == 1, -gamma, -1.0)
        np.add.at(val, s.j_pos, scale * sums)
        # tau row: tau = blr*ism*sum(rate_b y_b)
        rb = rate[s.swap_reacs]
        val[s.tau_pos_b] += -(rb * e["blr"] * e["ism"])
        if e["blr"] > 0:
            val[s.tau_pos_B] += -(e["tsw"] / e["blr"] * e["dblr"])
        val[s.tau_pos_S] += -(e["tsw"] / e["ism"] * e["dism"])
        # transfer (c-part) terms
        _, S = self.rhs(y, rate)
        surf, bulk = net.surface_list, net.bulk_list
        tp = s.tr_pos
        if S < 0:
            cov = min(1.0, e["sb"] / e["sm"])
            q = cov / e["sb"]
            yb = y[bulk]
            np.add.at(val, tp[:, 0], -gamma * (-q * yb))      # (s,S)
            np.add.at(val, tp[:, 1], -gamma * (q * yb))       # (b,S)
            np.add.at(val, tp[:, 2], -gamma * (-S * q))       # (s,b)
            np.add.at(val, tp[:, 3], -gamma * (S * q))        # (b,b)
            if e["sb"] < e["sm"]:
                dq = 0.0 if y[s.iS] <= 1e-30 else -1.0 / (e["sm"] * e["sm"])
                np.add.at(val, tp[:, 8], -gamma * (-S * yb * dq))   # (s,SURF)
                np.add.at(val, tp[:, 9], -gamma * (S * yb * dq))    # (b,SURF)
            else:
                dq = 0.0 if y[s.iB] <= 1e-30 else -1.0 / (e["sb"] * e["sb"])
                np.add.at(val, tp[:, 6], -gamma * (-S * yb * dq))   # (s,BULK)
                np.add.at(val, tp[:, 7], -gamma * (S * yb * dq))    # (b,BULK)
        else:
            ys = y[surf]
            np.add.at(val, tp[:, 0], -gamma * (-COV0 * ys))
            np.add.at(val, tp[:, 1], -gamma * (COV0 * ys))
            np.add.at(val, tp[:, 4], -gamma * (-S * COV0))    # (s,s)
            np.add.at(val, tp[:, 5], -gamma * (S * COV0))     # (b,s)
        val[s.dd_pos] += -gamma * ddensdot
        return val

    def to_dense(self, val):
        """Dense matrix in NEW (permuted) indexing from storage (before factorisation)."""
        s = self.s
        A = np.zeros((s.naug, s.naug))
        A[s.ent_row, s.ent_col] = val[: s.off_dense]
        A[s.n0:, s.n0:] = val[s.off_dense: s.off_dense + s.m * s.m].reshape(s.m, s.m)
        return A

    # -- factor / solve --------------------------------------------------------------
    def factor(self, val):
        s = self.s
        val = val.copy()
        val[s.zero_slot] = 0.0
        for L in s.f_levels:
            new = val[L["target"]].copy()
            prod = val[L["tl"]] * val[L["tu"]]
            for e in range(len(L["target"])):
                a, b = L["ptr"][e], L["ptr"][e + 1]
                if b > a:
                    new[e] -= prod[a:b].sum()
                if L["diag"][e] >= 0:
                    new[e] *= val[L["diag"][e]]      # pivots are stored as reciprocals
                elif L["diag"][e] == -2:
                    new[e] = 1.0 / new[e]
            val[L["target"]] = new
        # dense block: explicit inverse (the kernel uses Gauss-Jordan without pivoting)
        T = val[s.off_dense: s.off_dense + s.m * s.m].reshape(s.m, s.m)
        val[s.off_dense: s.off_dense + s.m * s.m] = _gauss_jordan_inverse(T).ravel()
        return val

    def solve(self, val, b_old):
        """Solve P x = b; b_old in OLD augmented indexing, returns x in OLD indexing."""
        s = self.s
        x = np.asarray(b_old, float)[s.perm].copy()
        for L in s.fwd_levels:
            for e, n in enumerate(L["rows"]):
                a, b = L["ptr"][e], L["ptr"][e + 1]
                x[n] -= np.dot(val[L["pos"][a:b]], x[L["cols"][a:b]])
        for t in range(s.m):
            a, b = s.tail_l_ptr[t], s.tail_l_ptr[t + 1]
            x[s.n0 + t] -= np.dot(val[s.tail_l_pos[a:b]], x[s.tail_l_col[a:b]])
        Tinv = val[s.off_dense: s.off_dense + s.m * s.m].reshape(s.m, s.m)
        x[s.n0:] = Tinv @ x[s.n0:]
        for L in s.bwd_levels:
            for e, n in enumerate(L["rows"]):
                a, b = L["ptr"][e], L["ptr"][e + 1]
                x[n] = (x[n] - np.dot(val[L["pos"][a:b]], x[L["cols"][a:b]])) * val[s.diag_pos[n]]
        out = np.empty(s.naug)
        out[s.perm] = x
        return out


def _gauss_jordan_inverse(T):
    """In-place style Gauss-Jordan inversion without pivoting (what the kernel does)."""
    A = T.copy()
    m = A.shape[0]
    for k in range(m):
        p = 1.0 / A[k, k]
        A[k, :] *= p
        A[k, k] = p
        col = A[:, k].copy()
        col[k] = 0.0
        A -= np.outer(col, A[k, :])
        A[:, k] = -col * p
        A[k, k] = p
    return A

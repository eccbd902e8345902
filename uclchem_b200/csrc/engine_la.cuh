// engine_la.cuh -- RHS, analytic Jacobian, fixed-pattern LU and triangular solves.
//
//   F / GETYDOT  chemistry.f90:294-352, odes.f90      -> rhs_eval()
//   DVJAC (dvode.f90:8182): Jacobian evaluation + saved copy (JSV=1) -> jac_eval(), form_p()
//   DGEFA (dvode.f90:11982) replaced by the generated sparse LU + dense inverse -> factor_p()
//   DVSOL/DGESL (dvode.f90:8698,12091)                -> lin_solve()
#pragma once
#include "engine_core.cuh"
#include "engine_gj.cuh"

// ---- team-program executor ---------------------------------------------------------
// desc[slot] = {term_begin, target | nterms<<16 | log2(team)<<28}; teams are aligned
// power-of-two groups of consecutive slots; lane l of a team sums terms l, l+T, ...
// A program is a list of units (one pass of NT slots of one level; units[k] = {first slot,
// end slot | sync-after << 30}, in constant memory); levels are separated by block barriers, or by
// a warp-level sync where consecutive levels fit in the 32 slots of warp 0.  A warp whose 32 slots lie beyond the unit's end goes straight to the barrier:
// most solve levels only occupy a few warps, and issue slots -- not bandwidth -- are what
// these phases cost.
//
// The tables live in L2 (the shared-memory carve-out leaves only ~28 KB of L1), and a level is
// a dependent chain descriptor -> terms -> shared-memory operands.  None of the table reads
// depends on numeric data, so the executor is software-pipelined across the level barriers:
// while unit k computes, the first U terms per lane (and the optional per-target `pre` word) of
// unit k+1 and the descriptor of unit k+2 are already in flight.  Only rows longer than U terms
// per lane fetch a second batch inside their own level.  (prefetch.global.L1 instead of register
// staging was measured first: the L1 hit rate barely moved and the term loads stayed exposed.)
struct TeamSlot {
    uint2 d;       // descriptor
    uint32_t sl;   // slot index
    bool warp_on;  // this warp has slots in the unit (warp-uniform)
};

// acc = p ? fma(a, x, acc) : acc as one predicated DFMA (keeps the select off the dependent chain)
__device__ __forceinline__ double masked_fma(double a, double x, double acc, bool p)
{
    asm("{ .reg .pred q; setp.ne.u32 q, %3, 0; @q fma.rn.f64 %0, %1, %2, %0; }" : "+d"(acc) : "d"(a), "d"(x), "r"((uint32_t)p));
    return acc;
}

// GROUP = 0: the whole CTA runs the program (block barriers).  GROUP = g > 0: the g threads starting at
// tid_base run it on their own (named barrier 2), e.g. the warps that idle during the dense inverse.
template <int U, int GROUP = 0, class TermT, class TermF, class PreF, class FinF>
__device__ __forceinline__ void run_levels(const uint32_t *__restrict__ desc, const TermT *__restrict__ terms,
                                           const uint32_t *units, int nunits, TermF term, PreF pre, FinF fin,
                                           uint32_t tid_base = 0u)
{
    const uint2 *d2 = reinterpret_cast<const uint2 *>(desc);
    const uint2 *units2 = reinterpret_cast<const uint2 *>(units);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t tid_rel = GROUP ? threadIdx.x - tid_base : threadIdx.x;
    auto fetch_desc = [&](int k) {
        TeamSlot t;
        t.d = make_uint2(0u, 0xFFFFu);
        t.sl = 0u;
        t.warp_on = false;
        if (k < nunits) {
            const uint2 un = units2[k];
            t.sl = un.x + tid_rel;
            t.warp_on = t.sl - lane < (un.y & 0x3FFFFFFFu); // slot ranges are padded to multiples of 32
            if (t.warp_on) t.d = __ldg(d2 + t.sl);
        }
        return t;
    };
    // first term batch of a unit: unconditional loads (the generator pads every term table by 32*U
    // entries, and whatever follows a row are valid terms of other rows); masked when summed
    auto fetch_terms = [&](const TeamSlot &t, TermT (&tb)[U], uint32_t &aux) {
        const uint32_t target = t.d.y & 0xFFFFu;
        const uint32_t T = 1u << ((t.d.y >> 28) & 7u), lit = t.sl & (T - 1u);
        aux = 0u;
        if (target != 0xFFFFu) {
            const TermT *tp = terms + t.d.x + lit;
#pragma unroll
            for (int u = 0; u < U; u++) tb[u] = __ldg(tp + (uint32_t)u * T);
            if (lit == 0u) aux = pre(target);
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) tb[u] = TermT{};
        }
    };
    TeamSlot cur = fetch_desc(0), nxt = fetch_desc(1);
    TermT tb[U];
    uint32_t aux;
    fetch_terms(cur, tb, aux);
    for (int k = 0; k < nunits; k++) {
        // ---- look-ahead: descriptor of unit k+2, first term batch of unit k+1 ----
        const TeamSlot nn = fetch_desc(k + 2);
        TermT tbn[U];
        uint32_t auxn;
        fetch_terms(nxt, tbn, auxn);
        // ---- unit k ----
        if (cur.warp_on) {
            const uint32_t target = cur.d.y & 0xFFFFu, n = (cur.d.y >> 16) & 0xFFFu;
            const int tl = (int)((cur.d.y >> 28) & 7u);
            const uint32_t T = 1u << tl, lit = cur.sl & (T - 1u);
            double acc = 0.0;
            if (target != 0xFFFFu) {
                // branch-free batch: all shared-memory operand loads, then one chain of masked FMAs
                double2 ax[U];
#pragma unroll
                for (int u = 0; u < U; u++) ax[u] = term(tb[u]);
                asm volatile("" ::: "memory"); // issue every operand load before the dependent FMA chain
#pragma unroll
                for (int u = 0; u < U; u++) acc = masked_fma(ax[u].x, ax[u].y, acc, lit + (uint32_t)u * T < n);
                const TermT *tp = terms + cur.d.x + lit + T * U;
                for (uint32_t q0 = lit + T * U; q0 < n; q0 += T * U, tp += T * U) {
                    TermT tc[U];
#pragma unroll
                    for (int u = 0; u < U; u++) tc[u] = __ldg(tp + (uint32_t)u * T);
#pragma unroll
                    for (int u = 0; u < U; u++) ax[u] = term(tc[u]);
                    asm volatile("" ::: "memory");
#pragma unroll
                    for (int u = 0; u < U; u++) acc = masked_fma(ax[u].x, ax[u].y, acc, q0 + (uint32_t)u * T < n);
                }
            }
            const int tmax = __reduce_max_sync(0xffffffffu, tl);
            for (int o = 1; o < (1 << tmax); o <<= 1) {
                double v = __shfl_xor_sync(0xffffffffu, acc, o);
                if ((uint32_t)o < T) acc += v;
            }
            if (target != 0xFFFFu && lit == 0u) fin(target, acc, aux);
        }
        const uint32_t sync = units2[k].y >> 30; // 2: block barrier, 1: the next level also lives in warp 0
        if (sync & 2u) {
            if (GROUP) asm volatile("barrier.sync 2, %0;" ::"n"(GROUP ? GROUP : 32) : "memory");
            else BLOCK_SYNC();
        } else if (sync) __syncwarp();
        cur = nxt;
        nxt = nn;
        aux = auxn;
#pragma unroll
        for (int u = 0; u < U; u++) tb[u] = tbn[u];
    }
}

// the 14 worker warps of rhs_eval (the last two warps evaluate the deferred photo rates)
#define NWORK (NT - 64)
#define WORK_SYNC() asm volatile("barrier.sync 3, %0;" ::"n"(NWORK) : "memory")

// flux_r = rate_r * prod_k yext[f_rk]  (reaction.py:779-819); one packed 8-byte entry per reaction
__device__ __forceinline__ void flux_range(Smem &s, int begin, int end, int first, int stride)
{
    const uint2 *tab = reinterpret_cast<const uint2 *>(net_flux_tab);
    const double *y = s.y;
#pragma unroll 4
    for (int idx = begin + first; idx < end; idx += stride) {
        uint2 e = __ldg(tab + idx);
        uint32_t r = e.x & 0xFFFFu;
        double v = s.rate[r];
        v *= y[e.x >> 16];
        v *= y[e.y & 1023u];
        v *= y[(e.y >> 10) & 1023u];
        v *= y[e.y >> 20];
        s.flux[r] = v;
    }
}

// =====================================================================================
// F (chemistry.f90:294-352) at the state in s.y (must be complete and block-visible).
// Output: ydot (any shared vector but s.y).  Ends with a barrier.
// =====================================================================================
__device__ __noinline__ void rhs_eval(Smem &s, double *ydot)
{
    Scalars &st = s.st;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double *y = s.y;
    TIMER_START
    if (warp >= NWARPS - 2) {
        // ---- photo warps: the two rates that follow the state (chemistry.f90:311-319) are long
        // single-lane chains (log10/pow/exp, spline walks).  They run here, next to the flux and
        // gather phases of the 14 worker warps; their reactions are not in the flux table / gather
        // program and are added to ydot after the gather (net_deferred_*).
        if (lane == 0) {
            const double d = y[NET_ID];
            const double h2col = 0.0 + 0.5 * y[NET_NH2] * d * (st.cloudsize / (double)1.0f);
            // (postprocess tracers with supplied column densities keep the rates of calculateReactionRates:
            // chemistry.f90:310-319)
            if (warp == NWARPS - 2) {
                const double k = st.pp_coldens ? s.rate[NET_NR_H2_HV] : st.scat_h2_pre * h2_self_shielding_dev(h2col);
                s.rate[NET_NR_H2_HV] = k;
                const double f = k * y[NET_DEFERRED_RE_H2];
                s.flux[NET_NR_H2_HV] = f;
                st.dflux[0] = f;
            } else {
                const double cocol = 0.0 + 0.5 * y[NET_NCO] * d * (st.cloudsize / (double)1.0f);
                if (!st.pp_coldens) {
                    st.cocol = cocol;
                    st.h2col = h2col;
                }
                const double k = st.pp_coldens ? s.rate[NET_NR_CO_HV] : co_photo_rate_dev(h2col, cocol, st.radfield, st.av);
                s.rate[NET_NR_CO_HV] = k;
                const double f = k * y[NET_DEFERRED_RE_CO];
                s.flux[NET_NR_CO_HV] = f;
                st.dflux[1] = f;
            }
        }
        BLOCK_SYNC(); // pairs with the barrier that ends the gather
    } else {
        // ---- phase A: ext scalars (warp 0) | plain fluxes (warps 1..13) ----
        if (warp == 0) {
            double sm = fmax(1e-30, y[NET_IS]), sb = fmax(1e-30, y[NET_IB]);
            double blr = fmin(1.0, NET_NSITES / (NET_GDR * sb));
            double ism = 1.0 / sm;
            double ts = 0.0;
            for (int k = lane; k < NET_NSWAP; k += 32) {
                int r = NET_SWAP_LO + k;
                ts += s.rate[r] * y[net_re1[r]] * blr;
            }
            ts = warp_sum(ts);
            if (lane == 0) {
                // y[NEQ+0] (the constant-one padding factor) is set once per CTA at kernel start
                s.y[NEQ + 1] = blr;
                s.y[NEQ + 2] = ism;
                s.y[NEQ + 3] = ts * ism;
                st.e_sm = sm; st.e_sb = sb; st.e_blr = blr; st.e_ism = ism; st.e_tsw = ts * ism;
                st.e_dblr = (blr >= 1.0 || y[NET_IB] <= 1e-30) ? 0.0 : -blr / sb;
                st.e_dism = (y[NET_IS] <= 1e-30) ? 0.0 : -ism * ism;
            }
        } else {
            flux_range(s, 0, NET_NPLAIN, tid - 32, NWORK - 32);
        }
        WORK_SYNC();
        // ---- phase B: fluxes that need the ext factors ----------------------------------------
        flux_range(s, NET_NPLAIN, NET_NFLUX, tid, NWORK);
        WORK_SYNC();
        // ---- gather: ydot_i = sum of signed fluxes (io_functions.py:562-581); units of NWORK slots;
        // ends with the block barrier the photo warps arrive at when their rates are done ----------
        const double *flux = s.flux;
        run_levels<8>(
            net_gather_desc, net_gather_terms, net_gather_units, NET_GATHER_NUNITS,
            [&](uint16_t t) { return make_double2(flux[t & 0x7FFFu], (t & 0x8000u) ? -1.0 : 1.0); },
            [](uint32_t) { return 0u; },
            [&](uint32_t target, double acc, uint32_t) { ydot[target] = acc; });
    }
    if (tid == NSURF + 1) {
        // deferred photo reactions (gas-phase rows only: disjoint from the surface/bulk sums below)
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = net_deferred_spec[k];
            if (i >= 0) ydot[i] += (double)net_deferred_sign[k] * st.dflux[k >> 2];
        }
    }
    // ---- three-phase transfer odes.f90:4815-5180 ------------------------------------------------
    {
        double S = 0.0, MB = 0.0;
        for (int k = lane; k < NSURF; k += 32) {
            S += ydot[net_surface_list[k]];
            MB += ydot[net_bulk_list[k]];
        }
        S = warp_sum(S);
        MB = warp_sum(MB);
        double Y = 0.0;
        const bool shrink = S < 0.0;
        for (int k = lane; k < NSURF; k += 32) Y += shrink ? y[net_bulk_list[k]] : y[net_surface_list[k]];
        Y = warp_sum(Y);
        double q = shrink ? fmin(1.0, st.e_sb / st.e_sm) / st.e_sb : C_COV0;
        if (lane == 0) { st.dbg[warp] = S; st.dbg[16 + warp] = Y; st.dbg[32 + warp] = q; st.dbg[48 + warp] = MB; }
        BLOCK_SYNC(); // every warp has read the uncorrected ydot
        if (tid < NSURF) {
            int si = net_surface_list[tid], bi = net_bulk_list[tid];
            double c = S * q * (shrink ? y[bi] : y[si]);
            ydot[si] -= c;
            ydot[bi] += c;
        } else if (tid == NSURF) {
            double C = S * q * Y;
            ydot[NET_IB] = MB + C;
            ydot[NET_IS] = S - C;
            ydot[NET_ID] = densdot_dev(st, y[NET_ID]);
            st.e_S = S;
        }
    }
    if (st.p[UCL_P_ENFORCECHARGECONSERVATION] != 0.0 && warp == NWARPS - 1) {
        // chemistry.f90:333-334 (ions are gas species: untouched by the transfer above)
        double q2 = 0.0;
        for (int i = lane; i < NSPEC; i += 32)
            if (net_is_ion[i]) q2 += ydot[i];
        q2 = warp_sum(q2);
        if (lane == 0) ydot[NET_NELEC] = q2;
    }
    BLOCK_SYNC();
    TIMER_ADD(cyc_rhs)
}

// =====================================================================================
// Jacobian of F in the generated bordered-sparse layout ("raw": plain derivative sums),
// evaluated at s.y with the st.e_* of an rhs_eval at the same state.  Leaves the raw
// values in s.val.  Ends with a barrier.
// =====================================================================================
__device__ __noinline__ void jac_eval(Smem &s)
{
    Scalars &st = s.st;
    const int tid = threadIdx.x;
    double *val = s.val;
    const double *y = s.y;
    TIMER_START
    for (int i = tid; i < NET_NVAL; i += NT) val[i] = 0.0;
    BLOCK_SYNC();
    {
        const double dblr = st.e_dblr, dism = st.e_dism;
        const double *rate = s.rate;
        run_levels<4>(
            net_jac_desc, reinterpret_cast<const uint2 *>(net_jac_terms), net_jac_units, NET_JAC_NUNITS,
            [&](uint2 t) {
                // t.x = reaction | kind<<28 | neg<<31 ; t.y = the other (non differentiated) factors
                double v = rate[t.x & 0xFFFFFFu];
                v *= y[t.y & 1023u];
                v *= y[(t.y >> 10) & 1023u];
                v *= y[t.y >> 20];
                uint32_t kind = (t.x >> 28) & 7u;
                if (kind == 1u) v *= dblr;
                else if (kind == 2u) v *= dism;
                return make_double2(v, (t.x >> 31) ? -1.0 : 1.0);
            },
            [](uint32_t) { return 0u; },
            [&](uint32_t target, double acc, uint32_t) { val[target] = acc; });
    }
    {
        // tau row and three-phase transfer terms (each position has exactly one writer)
        const double S = st.e_S;
        if (tid < NET_NSWAP) val[net_tau_pos_b[tid]] += s.rate[NET_SWAP_LO + tid] * st.e_blr * st.e_ism;
        if (tid == NET_NSWAP) {
            if (st.e_blr > 0.0) val[NET_TAU_POS_B] += st.e_tsw / st.e_blr * st.e_dblr;
            val[NET_TAU_POS_S] += st.e_tsw / st.e_ism * st.e_dism;
            val[NET_DD_POS] += ddensdot_dev(st, y[NET_ID]);
        }
        if (tid >= 128 && tid < 128 + NSURF) {
            int k = tid - 128;
            const uint16_t *tp = net_tr_pos + 10 * k;
            int si = net_surface_list[k], bi = net_bulk_list[k];
            if (S < 0.0) {
                double q = fmin(1.0, st.e_sb / st.e_sm) / st.e_sb;
                double yb = y[bi];
                val[tp[0]] += -q * yb;
                val[tp[1]] += q * yb;
                val[tp[2]] += -S * q;
                val[tp[3]] += S * q;
                if (st.e_sb < st.e_sm) {
                    double dq = (y[NET_IS] <= 1e-30) ? 0.0 : -1.0 / (st.e_sm * st.e_sm);
                    val[tp[8]] += -S * yb * dq;
                    val[tp[9]] += S * yb * dq;
                } else {
                    double dq = (y[NET_IB] <= 1e-30) ? 0.0 : -1.0 / (st.e_sb * st.e_sb);
                    val[tp[6]] += -S * yb * dq;
                    val[tp[7]] += S * yb * dq;
                }
            } else {
                double ys = y[si];
                val[tp[0]] += -C_COV0 * ys;
                val[tp[1]] += C_COV0 * ys;
                val[tp[4]] += -S * C_COV0;
                val[tp[5]] += S * C_COV0;
            }
        }
    }
    BLOCK_SYNC();
    TIMER_ADD(cyc_jac)
}

// P = I - gamma*J (aux rows: -J) from the raw Jacobian.  fresh: raw J is in s.val and is also
// saved to the CTA's global scratch jsv (DVODE's JSV=1 copy, dvode.f90:8352); otherwise it
// is read back from there.  Ends with a barrier.
__device__ __noinline__ void form_p(Smem &s, double gamma, double *jsv, bool fresh)
{
    TIMER_START
    for (int i = threadIdx.x; i < NET_NVAL; i += NT) {
        double j;
        if (fresh) {
            j = s.val[i];
            jsv[i] = j;
        } else {
            j = jsv[i];
        }
        uint32_t code = net_pos_code[i];
        double base = (code & 3u) == 1u ? 1.0 : ((code & 3u) == 2u ? -1.0 : 0.0);
        s.val[i] = base - ((code & 4u) ? 1.0 : gamma) * j;
    }
    BLOCK_SYNC();
    TIMER_ADD(cyc_jac)
}

// Right-hand side of the two constraint rows of the bordered Newton system.  The rows of BULK and
// SURFACE in P = I - gamma*J are the sums of their members' rows (odes.f90:5155-5180 defines their
// ydot as those sums), so the generated system stores them as  dBULK - sum_b d_b = r_BULK - sum_b r_b.
// On entry s.xs holds the plain residual r of every equation (new ordering); this replaces the two
// constraint entries by r - sum(members).  Dropping the right-hand side (i.e. d = sum of members)
// would leave the Nordsieck history of BULK/SURFACE uncorrected: rounding-level content in its
// higher columns is then amplified by every step-size increase (eta^j) and the variable drifts away
// from the sum of its members.  Ends with a barrier.
__device__ __forceinline__ void constraint_rhs(Smem &s)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 2) {
        const int16_t *list = warp == 0 ? net_surface_list : net_bulk_list;
        double acc = 0.0;
        for (int k = lane; k < NSURF; k += 32) acc += s.xs[net_iperm[list[k]]];
        acc = warp_sum(acc);
        if (lane == 0) s.xs[net_iperm[warp == 0 ? NET_IS : NET_IB]] -= acc;
    }
    BLOCK_SYNC();
}

// Newton right-hand side in elimination order, constraint rows included, in one pass:
// xs[new] = rl1*h*savf[o] - (rl1*yh1[o] + acor[o]) (DVNLSD, dvode.f90:7995-8000) for the plain rows;
// the BULK / SURFACE rows get r - sum(member r) (see constraint_rhs) from warps 0/1, which recompute
// the member residuals instead of reading them back (one barrier less per corrector iteration).
// Ends with a barrier.
__device__ __forceinline__ void newton_rhs(Smem &s)
{
    const Scalars &st = s.st;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double c1 = st.rl1 * st.h, rl1 = st.rl1;
    auto resid = [&](int o) { return c1 * s.savf[o] - (rl1 * s.yh[1][o] + s.acor[o]); };
    if (warp < 2) {
        const int16_t *list = warp == 0 ? net_surface_list : net_bulk_list;
        double acc = 0.0;
        for (int k = lane; k < NSURF; k += 32) acc += resid(list[k]);
        acc = warp_sum(acc);
        const int o = warp == 0 ? NET_IS : NET_IB;
        if (lane == 0) s.xs[net_iperm[o]] = resid(o) - acc;
    }
    if (tid < NAUG) {
        int o = net_perm[tid];
        if (o != NET_IS && o != NET_IB) s.xs[tid] = (o < NEQ) ? resid(o) : 0.0;
    }
    BLOCK_SYNC();
}

#ifdef UCLGPU_PRODUCT_FORM
// X = inv(L11), Y = inv(U11) on their closure patterns (uclchem_b200/product_form.py): level programs
// from the factor storage into the staging buffer -- the flux array, dead outside rhs_eval.  GROUP = 0:
// whole CTA, staging header written by the caller before a block barrier.  GROUP > 0: the GROUP threads
// from tid_base on, on their own named barrier.
template <int GROUP>
__device__ __forceinline__ void pf_inverse_levels(Smem &s, uint32_t tid_base)
{
    static_assert(NET_PF_NSTG <= NREAC, "staging buffer is the flux array");
    const double *val = s.val;
    double *stg = s.flux;
    if (GROUP) {
        for (int i = (int)(threadIdx.x - tid_base); i < NET_N0; i += GROUP) stg[NET_PF_DIAG0 + i] = val[net_diag_pos[i]];
        if (threadIdx.x == tid_base) stg[NET_PF_ONE] = 1.0;
        asm volatile("barrier.sync 2, %0;" ::"n"(GROUP ? GROUP : 32) : "memory");
        run_levels<4, GROUP>(
            net_pf_side_desc, net_pf_side_terms, net_pf_side_units, NET_PF_SIDE_NUNITS,
            [&](uint32_t t) { return make_double2(val[t >> 16], stg[t & 0xFFFFu]); },
            [](uint32_t target) { return (uint32_t)__ldg(net_pf_inv_scale + target); },
            [&](uint32_t target, double acc, uint32_t sc) { stg[target] = -stg[sc] * acc; }, tid_base);
    } else {
        run_levels<4>(
            net_pf_inv_desc, net_pf_inv_terms, net_pf_inv_units, NET_PF_INV_NUNITS,
            [&](uint32_t t) { return make_double2(val[t >> 16], stg[t & 0xFFFFu]); },
            [](uint32_t target) { return (uint32_t)__ldg(net_pf_inv_scale + target); },
            [&](uint32_t target, double acc, uint32_t sc) { stg[target] = -stg[sc] * acc; });
    }
}
#endif

// In-place inverse of the dense trailing block: blocked Gauss-Jordan elimination without pivoting
// (engine_gj.cuh), one GJ_B x GJ_B register tile per thread, two named barriers per block step.
// Returns false if a pivot is zero / not finite.  Ends with a block barrier.
__device__ __noinline__ bool dense_inverse(Smem &s)
{
    static_assert(GJ_NT * GJ_NT <= NT, "dense tile grid must fit the block");
    const int tid = threadIdx.x;
    double *T = s.val + NET_OFF_DENSE;
    double *pan = s.gj_pan;
    constexpr int GJ_NTHR = (GJ_NT * GJ_NT + 31) & ~31; // whole warps take part in the named barrier
    if (tid < GJ_NTHR) {
        GjTile t;
        gj_load(t, T, tid);
        gj_init_panel(pan, tid);
        gj_publish(t, pan, tid, 0);
        for (int kb = 0; kb < GJ_NT; kb++) {
            asm volatile("barrier.sync 1, %0;" ::"n"(GJ_NTHR) : "memory");
            gj_scale_row_panel(pan, tid, GJ_NTHR, kb);
            asm volatile("barrier.sync 1, %0;" ::"n"(GJ_NTHR) : "memory");
            gj_update(t, pan, tid, kb);
            gj_publish(t, pan, tid, kb + 1);
        }
        gj_store(t, T, tid);
    }
#if defined(UCLGPU_PRODUCT_FORM) && defined(UCLGPU_PF_OVERLAP)
    else {
        // The warps that have no tile (NT - GJ_NTHR threads) form the explicit inverses of the sparse
        // factors meanwhile: the inverse program reads only L11 / U11 (final since the factor levels)
        // and writes only the staging buffer, so it shares nothing with the Gauss-Jordan loop.
        static_assert(NT - GJ_NTHR == NET_PF_SIDE_THREADS, "side group size is baked into the unit table");
        pf_inverse_levels<NT - GJ_NTHR>(s, (uint32_t)GJ_NTHR);
    }
#endif
    BLOCK_SYNC();
    const bool ok = pan[GJ_OK] != 0.0;
#if !defined(UCLGPU_PRODUCT_FORM) && !defined(UCLGPU_COMPACT_SMEM)
    BLOCK_SYNC(); // the panels alias invd, which factor_p writes next: everyone must have read the flag
#endif
    return ok;
}

// Numeric factorisation of P (in s.val) on the generated pattern.  Sparse pivots end up stored
// as reciprocals.  Returns false if a pivot is zero / not finite.  Ends with a barrier.
__device__ __noinline__ bool factor_p(Smem &s, Blk &b)
{
    const int tid = threadIdx.x;
    double *val = s.val;
    TIMER_START
#ifndef UCLGPU_FACTOR_U
#define UCLGPU_FACTOR_U 8 /* terms per lane and batch of the factor program (build.py TAG_OPTIONS) */
#endif
    run_levels<UCLGPU_FACTOR_U>(
        net_factor_desc, net_factor_terms, net_factor_units, NET_FACTOR_NUNITS,
        [&](uint32_t t) { return make_double2(val[t >> 16], val[t & 0xFFFFu]); },
        [](uint32_t target) { return (uint32_t)__ldg(net_factor_diag + target); },
        [&](uint32_t target, double acc, uint32_t d) {
            double x = val[target] - acc;
            if (d == 0xFFFEu) x = 1.0 / x;       // sparse pivot: keep the reciprocal
            else if (d != 0xFFFFu) x *= val[d];  // L entry: scale by 1/pivot
            val[target] = x;
        });
    TIMER_ADD(cyc_factor)
    bool ok = dense_inverse(s);
#ifdef UCLGPU_PRODUCT_FORM
    {
        // X = inv(L11), Y = inv(U11) on their closure patterns: level programs into the staging buffer
        // (the flux array: dead outside rhs_eval), then one copy into the final positions -- entries that
        // exist in L11 / U11 are overwritten (not needed any more), fill entries go to the extra slots.
        double *stg = s.flux;
#ifndef UCLGPU_PF_OVERLAP
        if (tid < NET_N0) stg[NET_PF_DIAG0 + tid] = val[net_diag_pos[tid]];
        if (tid == 0) stg[NET_PF_ONE] = 1.0;
        BLOCK_SYNC();
        pf_inverse_levels<0>(s, 0u);
#endif  // with UCLGPU_PF_OVERLAP the idle warps of dense_inverse have done this already
        for (int i = tid; i < NET_PF_NX + NET_PF_NY; i += NT) val[net_pf_final_pos[i]] = stg[i];
        BLOCK_SYNC();
    }
#endif
    double bad = 0.0;
    if (tid < NET_N0) {
        double d = val[net_diag_pos[tid]];
        s.invd[tid] = d;
        if (!isfinite(d) || d == 0.0) bad = 1.0;
    }
    bad = block_sum(s, b, bad);
    TIMER_ADD(cyc_dense)
    return ok && bad == 0.0;
}

#ifdef UCLGPU_PRODUCT_FORM
// Solve P x = b, product form.  s.xs holds b in elimination order (block-visible); on return
// s.tmpv = x (SOLVE_RESULT).  Five levels: y1 = b1 + X b1 | b2' = b2 - L21 y1 | x2 = Tinv b2' |
// w = y1 - U12 x2 | x1 = Y w; xs and tmpv alternate as source and destination so no level
// reads what it writes.
__device__ __noinline__ void lin_solve(Smem &s)
{
    const int tid = threadIdx.x;
    const double *val = s.val;
    double *xs = s.xs, *tv = s.tmpv;
    TIMER_START
    run_levels<4>(
        net_pf_p1_desc, net_pf_p1_terms, net_pf_p1_units, NET_PF_P1_NUNITS,
        [&](uint32_t t) { return make_double2(val[t >> 16], xs[t & 0xFFFFu]); },
        [](uint32_t) { return 0u; },
        [&](uint32_t target, double acc, uint32_t) { tv[target] = xs[target] + acc; });
    run_levels<4>(
        net_pf_tail_desc, net_pf_tail_terms, net_pf_tail_units, NET_PF_TAIL_NUNITS,
        [&](uint32_t t) { return make_double2(val[t >> 16], tv[t & 0xFFFFu]); },
        [](uint32_t) { return 0u; },
        [&](uint32_t target, double acc, uint32_t) { xs[target] -= acc; });
    {
        // x_T = Tinv * b_T : 4 lanes per row, straight into tmpv[n0..]
        const double *T = val + NET_OFF_DENSE;
        int row = tid >> 2, sub = tid & 3;
        double acc = 0.0;
        if (row < MDENSE) {
#pragma unroll 4
            for (int j = sub; j < MDENSE; j += 4) acc += T[row * MDENSE + j] * xs[NET_N0 + j];
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (row < MDENSE && sub == 0) tv[NET_N0 + row] = acc;
    }
    BLOCK_SYNC();
    run_levels<4>(
        net_pf_p4_desc, net_pf_p4_terms, net_pf_p4_units, NET_PF_P4_NUNITS,
        [&](uint32_t t) { return make_double2(val[t >> 16], tv[t & 0xFFFFu]); },
        [](uint32_t) { return 0u; },
        [&](uint32_t target, double acc, uint32_t) { xs[target] = tv[target] - acc; });
    run_levels<4>(
        net_pf_p5_desc, net_pf_p5_terms, net_pf_p5_units, NET_PF_P5_NUNITS,
        [&](uint32_t t) { return make_double2(val[t >> 16], xs[t & 0xFFFFu]); },
        [](uint32_t) { return 0u; },
        [&](uint32_t target, double acc, uint32_t) { tv[target] = acc; });
    TIMER_ADD(cyc_solve)
}
#else
// Solve P x = b.  s.xs holds b in elimination order (block-visible); on return s.xs = x.
__device__ __noinline__ void lin_solve(Smem &s)
{
    const int tid = threadIdx.x;
    const double *val = s.val;
    double *xs = s.xs;
    TIMER_START
    // forward substitution through the sparse rows, then b_T -= L21 x (one program: fwd levels + tail)
    run_levels<4>(
        net_fwd_desc, net_fwd_terms, net_fwd_units, NET_FWD_NUNITS,
        [&](uint32_t t) { return make_double2(val[t >> 16], xs[t & 0xFFFFu]); },
        [](uint32_t) { return 0u; },
        [&](uint32_t target, double acc, uint32_t) { xs[target] -= acc; });
    {
        // x_T = Tinv * b_T : 4 lanes per row
        const double *T = val + NET_OFF_DENSE;
        int row = tid >> 2, sub = tid & 3;
        double acc = 0.0;
        if (row < MDENSE) {
#pragma unroll 4
            for (int j = sub; j < MDENSE; j += 4) acc += T[row * MDENSE + j] * xs[NET_N0 + j];
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (row < MDENSE && sub == 0) s.tmpv[row] = acc;
    }
    BLOCK_SYNC();
    if (tid < MDENSE) xs[NET_N0 + tid] = s.tmpv[tid];
    BLOCK_SYNC();
    run_levels<4>(
        net_bwd_desc, net_bwd_terms, net_bwd_units, NET_BWD_NUNITS,
        [&](uint32_t t) { return make_double2(val[t >> 16], xs[t & 0xFFFFu]); },
        [](uint32_t) { return 0u; },
        [&](uint32_t target, double acc, uint32_t) { xs[target] = (xs[target] - acc) * s.invd[target]; });
    TIMER_ADD(cyc_solve)
}
#endif // UCLGPU_PRODUCT_FORM

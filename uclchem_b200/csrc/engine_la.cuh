// engine_la.cuh -- RHS, analytic Jacobian, fixed-pattern LU and triangular solves.
//
//   F / GETYDOT  chemistry.f90:294-352, odes.f90      -> rhs_eval()
//   DVJAC (dvode.f90:8182): Jacobian evaluation + saved copy (JSV=1) -> jac_eval(), form_p()
//   DGEFA (dvode.f90:11982) replaced by the generated sparse LU + dense inverse -> factor_p()
//   DVSOL/DGESL (dvode.f90:8698,12091)                -> lin_solve()
#pragma once
#include "engine_core.cuh"

// ---- team-program executor ---------------------------------------------------------
// desc[slot] = {term_begin, target | nterms<<16 | log2(team)<<28}; teams are aligned
// power-of-two groups of consecutive slots; lane l of a team sums terms l, l+T, ...
// A program is a list of units (one pass of NT slots of one level; units[k] = {first slot,
// end slot | barrier-after << 31}, in constant memory); levels are separated by block
// barriers.  A warp whose 32 slots lie beyond the unit's end goes straight to the barrier:
// most solve levels only occupy a few warps, and issue slots -- not bandwidth -- are what
// these phases cost.
//
// The tables live in L2 (the shared-memory carve-out leaves only ~28 KB of L1), and a level is
// a dependent chain descriptor -> terms -> shared-memory operands.  None of the table reads
// depends on numeric data, so the executor runs ahead of the level barriers: the descriptor (and
// the optional per-target `pre` word) of unit k+1 is loaded into registers while unit k computes,
// and the term lines of unit k+1 are pulled into L1 with prefetch instructions before the barrier
// of unit k.  Device functions only get ~60 registers here (the call chain shares the 128 of a
// 512-thread CTA), so the terms themselves are not held in registers across the barrier.
struct TeamSlot {
    uint2 d;       // descriptor
    uint32_t sl;   // slot index
    uint32_t aux;  // pre(target)
    bool warp_on;  // this warp has slots in the unit (warp-uniform)
};

// acc = p ? fma(a, x, acc) : acc as one predicated DFMA (keeps the select off the dependent chain)
__device__ __forceinline__ double masked_fma(double a, double x, double acc, bool p)
{
    asm("{ .reg .pred q; setp.ne.u32 q, %3, 0; @q fma.rn.f64 %0, %1, %2, %0; }" : "+d"(acc) : "d"(a), "d"(x), "r"((uint32_t)p));
    return acc;
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int U, class TermT, class TermF, class PreF, class FinF>
__device__ __forceinline__ void run_levels(const uint32_t *__restrict__ desc, const TermT *__restrict__ terms,
                                           const uint32_t *units, int nunits, TermF term, PreF pre, FinF fin)
{
    const uint2 *d2 = reinterpret_cast<const uint2 *>(desc);
    const uint2 *units2 = reinterpret_cast<const uint2 *>(units);
    const uint32_t lane = threadIdx.x & 31u;
    constexpr uint32_t PER_LINE = 128u / sizeof(TermT);
    auto fetch_desc = [&](int k) {
        TeamSlot t;
        t.d = make_uint2(0u, 0xFFFFu);
        t.sl = 0u;
        t.aux = 0u;
        t.warp_on = false;
        if (k < nunits) {
            const uint2 un = units2[k];
            t.sl = un.x + threadIdx.x;
            t.warp_on = t.sl - lane < (un.y & 0x7FFFFFFFu); // slot ranges are padded to multiples of 32
            if (t.warp_on) t.d = __ldg(d2 + t.sl);
        }
        return t;
    };
    // second half of the look-ahead (needs the descriptor to have arrived): term lines -> L1, pre word
    auto look_ahead = [&](TeamSlot &t) {
        const uint32_t target = t.d.y & 0xFFFFu, n = (t.d.y >> 16) & 0xFFFu;
        const uint32_t T = 1u << ((t.d.y >> 28) & 7u), lit = t.sl & (T - 1u);
        if (target != 0xFFFFu) {
            const TermT *tp = terms + t.d.x;
            if (lit * PER_LINE < n) prefetch_l1(tp + lit * PER_LINE);
            if (lit == 0u) {
                prefetch_l1(tp + n - 1u);
                t.aux = pre(target);
            }
        }
    };
    TeamSlot cur = fetch_desc(0);
    look_ahead(cur);
    for (int k = 0; k < nunits; k++) {
        TeamSlot nxt = fetch_desc(k + 1); // in flight while unit k computes
        if (cur.warp_on) {
            const uint32_t target = cur.d.y & 0xFFFFu, n = (cur.d.y >> 16) & 0xFFFu;
            const int tl = (int)((cur.d.y >> 28) & 7u);
            const uint32_t T = 1u << tl, lit = cur.sl & (T - 1u);
            double acc = 0.0;
            if (target != 0xFFFFu) {
                // branch-free batch: U unconditional table loads (the generator pads every term table
                // by 32*U entries, and whatever follows a row are valid terms of other rows), then all
                // shared-memory operand loads, then one chain of masked FMAs in term order
                const TermT *tp = terms + cur.d.x + lit;
                for (uint32_t q0 = lit; q0 < n; q0 += T * U, tp += T * U) {
                    TermT tb[U];
#pragma unroll
                    for (int u = 0; u < U; u++) tb[u] = __ldg(tp + (uint32_t)u * T);
                    double2 ax[U];
#pragma unroll
                    for (int u = 0; u < U; u++) ax[u] = term(tb[u]);
                    asm volatile("" ::: "memory"); // issue every operand load before the dependent FMA chain
#pragma unroll
                    for (int u = 0; u < U; u++) acc = masked_fma(ax[u].x, ax[u].y, acc, q0 + (uint32_t)u * T < n);
                }
            }
            const int tmax = __reduce_max_sync(0xffffffffu, tl);
            for (int o = 1; o < (1 << tmax); o <<= 1) {
                double v = __shfl_xor_sync(0xffffffffu, acc, o);
                if ((uint32_t)o < T) acc += v;
            }
            if (target != 0xFFFFu && lit == 0u) fin(target, acc, cur.aux);
        }
        look_ahead(nxt);
        if (units2[k].y >> 31) BLOCK_SYNC();
        cur = nxt;
    }
}

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
    // ---- phase A: ext scalars + H2 photo rate (warp 0) | CO photo rate (warp 1) | plain fluxes ----
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
            // chemistry.f90:311-319: the H2 photo rate follows the current H2 abundance
            double h2col = 0.0 + 0.5 * y[NET_NH2] * y[NET_ID] * (st.cloudsize / (double)1.0f);
            s.rate[NET_NR_H2_HV] = st.scat_h2_pre * h2_self_shielding_dev(h2col);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            double d = y[NET_ID];
            double cocol = 0.0 + 0.5 * y[NET_NCO] * d * (st.cloudsize / (double)1.0f);
            double h2col = 0.0 + 0.5 * y[NET_NH2] * d * (st.cloudsize / (double)1.0f);
            st.cocol = cocol;
            st.h2col = h2col;
            s.rate[NET_NR_CO_HV] = co_photo_rate_dev(h2col, cocol, st.radfield, st.av);
        }
    } else {
        flux_range(s, 0, NET_NPLAIN, tid - 64, NT - 64);
    }
    BLOCK_SYNC();
    // ---- phase B: fluxes that need ext factors or the fresh photo rates --------------------
    flux_range(s, NET_NPLAIN, NREAC, tid, NT);
    BLOCK_SYNC();
    // ---- gather: ydot_i = sum of signed fluxes (io_functions.py:562-581) ---------------------
    {
        const double *flux = s.flux;
        run_levels<8>(
            net_gather_desc, net_gather_terms, net_gather_units, NET_GATHER_NUNITS,
            [&](uint16_t t) { return make_double2(flux[t & 0x7FFFu], (t & 0x8000u) ? -1.0 : 1.0); },
            [](uint32_t) { return 0u; },
            [&](uint32_t target, double acc, uint32_t) { ydot[target] = acc; });
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
        if (st.p[UCL_P_ENFORCECHARGECONSERVATION] != 0.0 && warp == NWARPS - 1) {
            // chemistry.f90:333-334 (ions are gas species: untouched by the transfer above)
            double q2 = 0.0;
            for (int i = lane; i < NSPEC; i += 32)
                if (net_is_ion[i]) q2 += ydot[i];
            q2 = warp_sum(q2);
            if (lane == 0) ydot[NET_NELEC] = q2;
        }
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

// In-place inverse of the dense trailing block by Gauss-Jordan elimination without
// pivoting.  Each thread keeps a GJ_R x GJ_C tile of the block in registers for all M
// steps; per step only the pivot row, pivot column and 1/pivot go through shared memory.
// Every tile does the uniform update a_ij -= col_i * (row_j * p); only the few tiles that
// contain the pivot row or column patch their entries afterwards (a_kj p, -a_ik p, p).
__device__ __noinline__ bool dense_inverse(Smem &s)
{
    static_assert(GJ_TR * GJ_TC <= NT, "dense tile grid must fit the block");
    const int tid = threadIdx.x;
    double *T = s.val + NET_OFF_DENSE;
    const int tr = tid / GJ_TC, tc = tid - tr * GJ_TC;
    const bool active = tid < GJ_TR * GJ_TC;
    const int i0 = tr * GJ_R, j0 = tc * GJ_C;
    double a[GJ_R][GJ_C];
#pragma unroll
    for (int r = 0; r < GJ_R; r++)
#pragma unroll
        for (int c = 0; c < GJ_C; c++) {
            int i = i0 + r, j = j0 + c;
            a[r][c] = (active && i < MDENSE && j < MDENSE) ? T[i * MDENSE + j] : 0.0;
        }
    bool ok = true;
    constexpr int GJ_NTHR = (GJ_TR * GJ_TC + 31) & ~31; // whole warps take part in the named barrier
    if (tid < GJ_NTHR) {
        for (int k = 0; k < MDENSE; k++) {
            const int buf = k & 1;
            const int rk = k - i0, ck = k - j0;
            const bool own_r = active && (unsigned)rk < (unsigned)GJ_R, own_c = active && (unsigned)ck < (unsigned)GJ_C;
            if (own_r) {
#pragma unroll
                for (int r = 0; r < GJ_R; r++)
                    if (r == rk) {
#pragma unroll
                        for (int c = 0; c < GJ_C; c++) s.gj_row[buf][j0 + c] = a[r][c];
                    }
            }
            if (own_c) {
#pragma unroll
                for (int c = 0; c < GJ_C; c++)
                    if (c == ck) {
#pragma unroll
                        for (int r = 0; r < GJ_R; r++) {
                            s.gj_col[buf][i0 + r] = a[r][c];
                            if (r == rk) s.gj_piv[buf] = 1.0 / a[r][c];
                        }
                    }
            }
            asm volatile("barrier.sync 1, %0;" ::"r"(GJ_NTHR) : "memory");
            const double p = s.gj_piv[buf];
            if (!isfinite(p) || p == 0.0) ok = false;
            double rp[GJ_C], cl[GJ_R];
#pragma unroll
            for (int c = 0; c < GJ_C; c++) rp[c] = s.gj_row[buf][j0 + c] * p;
#pragma unroll
            for (int r = 0; r < GJ_R; r++) cl[r] = s.gj_col[buf][i0 + r];
#pragma unroll
            for (int r = 0; r < GJ_R; r++)
#pragma unroll
                for (int c = 0; c < GJ_C; c++) a[r][c] -= cl[r] * rp[c];
            if (own_r) {
#pragma unroll
                for (int r = 0; r < GJ_R; r++)
                    if (r == rk) {
#pragma unroll
                        for (int c = 0; c < GJ_C; c++) a[r][c] = rp[c];
                    }
            }
            if (own_c) {
#pragma unroll
                for (int c = 0; c < GJ_C; c++)
                    if (c == ck) {
#pragma unroll
                        for (int r = 0; r < GJ_R; r++) a[r][c] = (r == rk) ? p : -cl[r] * p;
                    }
            }
        }
#pragma unroll
        for (int r = 0; r < GJ_R; r++)
#pragma unroll
            for (int c = 0; c < GJ_C; c++) {
                int i = i0 + r, j = j0 + c;
                if (active && i < MDENSE && j < MDENSE) T[i * MDENSE + j] = a[r][c];
            }
        if (tid == 0) s.gj_piv[0] = ok ? 1.0 : 0.0; // every participant saw the same pivots
    }
    BLOCK_SYNC();
    return s.gj_piv[0] != 0.0;
}

// Numeric factorisation of P (in s.val) on the generated pattern.  Sparse pivots end up stored
// as reciprocals.  Returns false if a pivot is zero / not finite.  Ends with a barrier.
__device__ __noinline__ bool factor_p(Smem &s, Blk &b)
{
    const int tid = threadIdx.x;
    double *val = s.val;
    TIMER_START
    run_levels<8>(
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

// Solve P x = b.  s.xs holds b in elimination order (block-visible); on return s.xs = x.
__device__ __noinline__ void lin_solve(Smem &s)
{
    const int tid = threadIdx.x;
    const double *val = s.val;
    double *xs = s.xs;
    TIMER_START
    // forward substitution through the sparse rows, then b_T -= L21 x (one program: fwd levels + tail)
    run_levels<8>(
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
    run_levels<8>(
        net_bwd_desc, net_bwd_terms, net_bwd_units, NET_BWD_NUNITS,
        [&](uint32_t t) { return make_double2(val[t >> 16], xs[t & 0xFFFFu]); },
        [](uint32_t) { return 0u; },
        [&](uint32_t target, double acc, uint32_t) { xs[target] = (xs[target] - acc) * s.invd[target]; });
    TIMER_ADD(cyc_solve)
}

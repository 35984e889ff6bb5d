// Fused 3D energy-condition kernel: phi -> grad phi -> grad grad phi -> K_ij -> rho -> sums.
//
// Replaces the array pipeline of compute_energy_3d
// (/root/reference/scripts/reproduce_rodal_exact.py:171-224) and phi_exact_rodal /
// g_rodal_exact (/root/reference/src/irrotational_warp/potential.py:37-99); no
// intermediate ever reaches HBM.
//
// Decomposition.  The volume is C-order [i, j, k] (k contiguous).  A work item is
// (tile of the (j, k) plane) x (range of i planes); the CTA marches along i.
// The CTA's region is the tile plus a 2-point halo: RJ rows (j) x 64 columns (k),
// lanes along k.  Warp w owns region rows jr = w, w+NW, ...; lane tx owns columns tx, tx+32.
//
// Shared memory, 11 planes of RJ x 64 doubles:
//     phi ring [3]   (plane q in slot q mod 3)
//     phi_x ring [4] (plane q in slot q mod 4)
//     phi_y, phi_z   double-buffered (plane q in buffer q mod 2)
// The rings are sized so that ONE barrier per output plane suffices, and that barrier is split
// (mbarrier arrive ... wait).  Per output plane i every warp runs
//     O(i)  : second derivatives, rho, predicated sums of plane i             -- smem heavy, then FP64
//     C(i+1): phi_y(i+1), phi_z(i+1) from the phi(i+1) plane                  -- smem heavy
//     -- arrive --
//     P(i+3): evaluate phi(i+3) at the own columns and finish phi_x(i+2)      -- FP64 heavy, ~no smem
//     -- wait --
// P touches only the own columns' ring entries until two steps later, so it is legal between the
// arrival and the wait; warps drift apart by up to one task, and the shared-memory-bound tasks of
// some warps overlap the FP64-bound task of others (measured: 59 -> 77 Gpoints/s over the version
// with two __syncthreads per plane, whose smem and FP64 phases alternated in lock step).
// Row loops are deliberately NOT unrolled (two columns of ILP per thread): the IEEE division / sqrt
// sequences are long and an unrolled body overflows the instruction cache.
// Sums are kept in registers and flushed (warp shuffle + fixed-order block tree) once per
// IWB_FLUSH_PLANES planes into partial[flush block][tile][row]; a second kernel adds the tiles in
// fixed order.  Flush blocks are aligned to the GLOBAL plane index, so the result is bit-identical
// for any slab decomposition along i.
#pragma once
#include <type_traits>

#include "iwb_device.cuh"
#include "../../include/iwb200.h"

#ifndef IWB_BARRIER_BACKOFF_NS
#define IWB_BARRIER_BACKOFF_NS 100
#endif
#ifndef IWB_DBG
#define IWB_DBG 0      // development only: 1 = skip task O, 2 = skip the phi evaluation, 3 = both
#endif

namespace iwb {

constexpr int RK = 64;              // region columns (k): two per lane
constexpr int TK_MAX = RK - 4;      // tile columns
constexpr int ROW = IWB_ROW_WIDTH;  // doubles per partial row
constexpr int FLUSH = IWB_FLUSH_PLANES;
constexpr int SMEM_PAD = 8;         // guard doubles before and after the planes (lane 0 / lane 31 reads of idx -+ 1)
constexpr int NPLANES = 11;

struct K3Params {
    int n0, n1, n2;
    int i_begin, i_end;            // planes to evaluate (global indices)
    int seg_len;                   // planes per work item (multiple of FLUSH)
    int nseg, ntj, ntk, nitems;
    const double* x;               // [n0]  (float32 values widened for the f32ref path)
    const double* xx;              // [n0]  x*x   (rounded in the path's coordinate dtype)
    const double* yy;              // [n1]  y*y
    const double* zz;              // [n2]  z*z
    PhiConst pc;
    AxisCoef cx, cy, cz;
    double dV;                     // (dx*dy)*dz
    double c16pi, rc16pi;          // 16.0*pi and RN(1/(16.0*pi))
    double shell_thr[IWB_MAX_SHELLS];   // r <= R  <=>  s <= shell_thr (s = r*r before sqrt), see host
    double* partial;               // [nfb_total][ntiles][ROW]
    double* rho_out;               // NULL or [i_end-i_begin][n1][n2]
    unsigned int* counter;         // dynamic work-item counter (0 at launch; reset by k_reduce_tiles)
    int i_zero, j_zero;            // index of the coordinate that is exactly 0 on axis 0 / 1 (-1: none): r == 0 is possible there
};

__host__ __device__ inline int tile_begin(int t, int nt, int n) { return (int)(((long long)t * n) / nt); }

template <int RJ, int NW>
struct Geo {
    static constexpr int PLANE = RJ * RK;
    static constexpr int TJ_MAX = RJ - 4;
    static constexpr int NT = NW * 32;
    static constexpr int SCRATCH = NW * ROW;     // block-reduction scratch (doubles)
    static constexpr size_t SMEM_BYTES = (size_t(NPLANES) * PLANE + 2 * SMEM_PAD + RJ + SCRATCH) * sizeof(double);
};

template <int DT, int NSH, int RJ, int NW, bool EMIT, int MINB = 1, int MODE = 0>
__global__ void __launch_bounds__(NW * 32, MINB)
k_energy3d(const __grid_constant__ K3Params P)
{
    using G = Geo<RJ, NW>;
    constexpr int PLANE = G::PLANE;
    constexpr int NSHA = NSH > 0 ? NSH : 1;
    extern __shared__ double smem_raw[];
    double* const phi_ring = smem_raw + SMEM_PAD;          // 3 planes
    double* const px_ring = phi_ring + 3 * PLANE;          // 4 planes
    double* const py_buf = phi_ring + 7 * PLANE;           // 2 planes
    double* const pz_buf = phi_ring + 9 * PLANE;           // 2 planes
    double* const yy_s = phi_ring + NPLANES * PLANE + SMEM_PAD;   // [RJ] y*y of the region rows
    double* const scratch = yy_s + RJ;                     // [NW][ROW]

    const int tx = threadIdx.x & 31;
    const int ty = threadIdx.x >> 5;
    const int ntiles = P.ntj * P.ntk;
    const AxisCoef cx = P.cx, cy = P.cy, cz = P.cz;
    const PhiConst pc = P.pc;

    __shared__ int s_item;
    // Split (arrive / wait) barrier of the march: one mbarrier phase per output plane.
    __shared__ unsigned long long s_mbar;
    const unsigned mbar_addr = (unsigned)__cvta_generic_to_shared(&s_mbar);
    unsigned mbar_parity = 0;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_addr), "r"(NW * 32));
    __syncthreads();
    auto step_arrive = [&]() { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar_addr) : "memory"); };
    auto step_wait = [&]() {
        unsigned done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mbar_addr), "r"(mbar_parity) : "memory");
        while (!done) {
            // back off: a polling warp would otherwise take issue slots from the warps still working
            __nanosleep(IWB_BARRIER_BACKOFF_NS);
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(mbar_addr), "r"(mbar_parity) : "memory");
        }
        mbar_parity ^= 1u;
    };
    for (;;) {
        // dynamic schedule: items are handed out in index order; their partial sums go to fixed slots,
        // so the result does not depend on which CTA ran which item
        __syncthreads();                 // previous item's shared memory (and s_item) is dead
        if (threadIdx.x == 0) s_item = (int)atomicAdd(P.counter, 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= P.nitems) break;
        const int seg = item / ntiles;
        // Tiles that touch a global face cost ~1.3x an interior tile (one-sided formulas, divergent lanes), so each
        // segment hands them out first: longest-processing-time-first keeps the tail of the dynamic schedule short.
        // rank -> tile: row tj = 0, row tj = ntj-1, the two edge columns of the rows between, then the interior.
        int tile;
        {
            const int rnk = item - seg * ntiles;
            if (P.ntj <= 2 || P.ntk <= 2) tile = rnk;
            else if (rnk < P.ntk) tile = rnk;
            else if (rnk < 2 * P.ntk) tile = (P.ntj - 1) * P.ntk + (rnk - P.ntk);
            else {
                const int r1 = rnk - 2 * P.ntk, nside = 2 * (P.ntj - 2);
                if (r1 < nside) tile = (1 + (r1 >> 1)) * P.ntk + ((r1 & 1) ? P.ntk - 1 : 0);
                else {
                    const int r2 = r1 - nside, wi = P.ntk - 2;
                    tile = (1 + r2 / wi) * P.ntk + 1 + r2 % wi;
                }
            }
        }
        const int tj = tile / P.ntk, tk = tile - tj * P.ntk;
        const int j0 = tile_begin(tj, P.ntj, P.n1), Tj = tile_begin(tj + 1, P.ntj, P.n1) - j0;
        const int k0 = tile_begin(tk, P.ntk, P.n2), Tk = tile_begin(tk + 1, P.ntk, P.n2) - k0;
        const int i0 = P.i_begin + seg * P.seg_len;
        const int i1 = min(P.i_end, i0 + P.seg_len);

        // region rows [jr_lo, jr_hi] exist in the grid; j == 0 is region row 2 of a low-edge tile,
        // j == n1-1 is region row Tj+1 of a high-edge tile
        const bool t_jlo = (j0 == 0), t_jhi = (j0 + Tj == P.n1);
        const bool t_kedge = (k0 == 0) || (k0 + Tk == P.n2);
        const bool tile_edge = t_jlo || t_jhi || t_kedge;
        const int jr_lo = t_jlo ? 2 : 0, jr_hi = t_jhi ? Tj + 1 : Tj + 3;
        const int py_lo = max(jr_lo, 1), py_hi = min(jr_hi, Tj + 2);

        if (threadIdx.x < RJ) yy_s[threadIdx.x] = P.yy[min(max(j0 - 2 + (int)threadIdx.x, 0), P.n1 - 1)];

        // this lane's two columns
        double zz0, zz1;
        unsigned cf0, cf1;       // bit1 k==0, bit2 k==n2-1, bit4 output column
        {
            const int ka = k0 - 2 + tx, kb = ka + 32;
            zz0 = P.zz[min(max(ka, 0), P.n2 - 1)];
            zz1 = P.zz[min(max(kb, 0), P.n2 - 1)];
            cf0 = ((ka == 0) ? 2u : 0u) | ((ka == P.n2 - 1) ? 4u : 0u) | ((tx >= 2 && tx <= Tk + 1) ? 16u : 0u);
            cf1 = ((kb == 0) ? 2u : 0u) | ((kb == P.n2 - 1) ? 4u : 0u) | ((tx + 32 <= Tk + 1) ? 16u : 0u);
        }

        // ---- accumulators (registers) -------------------------------------------------
        double acc_pos = 0.0, acc_neg = 0.0;
        double sh_pos[NSHA], sh_neg[NSHA];
        int c_neg = 0;
        int c_zn = 0;            // points that are neither positive nor negative: zeros in the low half, NaNs in the high half
        int flush_from = 0;      // first plane of the running flush interval (set below)
        int sh_cnt[NSHA];
#pragma unroll
        for (int s = 0; s < NSHA; ++s) { sh_pos[s] = 0.0; sh_neg[s] = 0.0; sh_cnt[s] = 0; }

        const int w0 = min(max(i0 - 1, 0), P.n0 - 3);
        const int pstart = max(0, w0 - 1);

        // d/dz (along lanes) and d/dy (along rows) of a staged plane.  The EDGE = false flavours are
        // the pure interior formula (no branches: the two columns of a thread interleave freely);
        // EDGE = true adds the one-sided formulas at the global faces.
        auto dZ = [&](const double* B, int idx, unsigned cf, auto edge_tag) -> double {
            double d = cdivm<MODE>(B[idx + 1] - B[idx - 1], cz.h2, cz.rh2);
            if (decltype(edge_tag)::value) {
                if (cf & 2u) d = edge3(cz.a0, B[idx], cz.b0, B[idx + 1], cz.c0, B[idx + 2]);
                else if (cf & 4u) d = edge3(cz.a1, B[idx - 2], cz.b1, B[idx - 1], cz.c1, B[idx]);
            }
            return d;
        };
        // jedge: 0 interior, 1 row j == 0, 2 row j == n1-1 (warp-uniform)
        auto dY = [&](const double* B, int idx, int jedge, auto edge_tag) -> double {
            if (decltype(edge_tag)::value) {
                if (jedge == 1) return edge3(cy.a0, B[idx], cy.b0, B[idx + RK], cy.c0, B[idx + 2 * RK]);
                if (jedge == 2) return edge3(cy.a1, B[idx - 2 * RK], cy.b1, B[idx - RK], cy.c1, B[idx]);
            }
            return cdivm<MODE>(B[idx + RK] - B[idx - RK], cy.h2, cy.rh2);
        };

        // ---- task P(p): phi(p) at the own columns, and the x-derivative it completes -----------------
        // phi_x(p-1) = (phi(p) - phi(p-2)) / 2dx needs phi(p-2) of the own column only.
        auto taskP = [&](int p, int s3 /* p mod 3 */, double xp, double xxp, auto edge_tag) {
            constexpr bool EDGE = decltype(edge_tag)::value;
            const int s3m1 = (s3 == 0) ? 2 : s3 - 1, s3m2 = (s3 == 2) ? 0 : s3 + 1;
            double* const ph_p = phi_ring + s3 * PLANE;
            const double* const ph_m1 = phi_ring + s3m1 * PLANE;               // plane p-1
            const double* const ph_m2 = phi_ring + s3m2 * PLANE;               // plane p-2
            double* const px_w = px_ring + ((p + 3) & 3) * PLANE;              // phi_x(p-1)
            const bool do_px = (p - 2 >= pstart);                              // interior formula
            const bool do_px0 = EDGE && (p == 2) && (pstart == 0);             // phi_x(0), one-sided
#pragma unroll 1
            for (int jr = ty; jr <= jr_hi; jr += NW) {        // rows jr == ty (mod NW), as in task O
                if (jr < jr_lo) continue;
                const int base = jr * RK + tx;
                double f[2];
                if (IWB_DBG & 2) { f[0] = xxp + zz0; f[1] = xxp + zz1; }
                else if (DT == IWB_F64) {
                    const double sxy = xxp + yy_s[jr];
                    const double s[2] = {sxy + zz0, sxy + zz1};
                    // the row through the origin (r == 0 at one lane) takes the reference flavour
                    const bool origin_row = (p == P.i_zero) && (j0 - 2 + jr == P.j_zero);
                    // (exotic grids / parameters fail the host's operand-range proof: reference flavour throughout)
                    if (MODE == 0) phi_exact_f64<2, true>(s, xp, pc, f, origin_row || !pc.prechecked);
                    else phi_fast_f64<2>(s, xp, pc, f, origin_row || !pc.prechecked);
                } else {
                    const float sxy = __fadd_rn((float)xxp, (float)yy_s[jr]);
                    const float s[2] = {__fadd_rn(sxy, (float)zz0), __fadd_rn(sxy, (float)zz1)};
                    const bool origin_row = (p == P.i_zero) && (j0 - 2 + jr == P.j_zero);
                    phi_exact_f32ref<2, true>(s, (float)xp, pc, f, origin_row || !pc.prechecked);
                }
                ph_p[base] = f[0];
                ph_p[base + 32] = f[1];
                if (do_px) {
                    px_w[base] = cdivm<MODE>(f[0] - ph_m2[base], cx.h2, cx.rh2);
                    px_w[base + 32] = cdivm<MODE>(f[1] - ph_m2[base + 32], cx.h2, cx.rh2);
                }
                if (EDGE && do_px0) {
                    px_ring[base] = edge3(cx.a0, ph_m2[base], cx.b0, ph_m1[base], cx.c0, f[0]);
                    px_ring[base + 32] = edge3(cx.a0, ph_m2[base + 32], cx.b0, ph_m1[base + 32], cx.c0, f[1]);
                }
            }
        };

        // ---- phi_x(n0-1), one-sided, from the last three phi planes (own columns) -----------------------
        auto taskP_last = [&]() {
            const int p = P.n0 - 1;
            const double* const ph_p = phi_ring + (p % 3) * PLANE;
            const double* const ph_m1 = phi_ring + ((p + 2) % 3) * PLANE;
            const double* const ph_m2 = phi_ring + ((p + 1) % 3) * PLANE;
            double* const px_w = px_ring + (p % 4) * PLANE;
#pragma unroll 1
            for (int jr = ty; jr <= jr_hi; jr += NW) {
                if (jr < jr_lo) continue;
                const int base = jr * RK + tx;
                px_w[base] = edge3(cx.a1, ph_m2[base], cx.b1, ph_m1[base], cx.c1, ph_p[base]);
                px_w[base + 32] = edge3(cx.a1, ph_m2[base + 32], cx.b1, ph_m1[base + 32], cx.c1, ph_p[base + 32]);
            }
        };

        // ---- task C(q): phi_y(q), phi_z(q) from the phi(q) plane (reads neighbours' columns) ------------
        auto taskC = [&](int q, int s3 /* q mod 3 */, auto edge_tag) {
            constexpr bool EDGE = decltype(edge_tag)::value;
            const double* const ph = phi_ring + s3 * PLANE;
            double* const py_w = py_buf + (q & 1) * PLANE;
            double* const pz_w = pz_buf + (q & 1) * PLANE;
#pragma unroll 1
            for (int jr = py_lo + ty; jr <= py_hi; jr += NW) {
                const int base = jr * RK + tx;
                const int jedge = EDGE ? ((t_jlo && jr == 2) ? 1 : ((t_jhi && jr == Tj + 1) ? 2 : 0)) : 0;
                py_w[base] = dY(ph, base, jedge, edge_tag);
                py_w[base + 32] = dY(ph, base + 32, jedge, edge_tag);
                if (jr >= 2 && jr <= Tj + 1) {
                    pz_w[base] = dZ(ph, base, cf0, edge_tag);
                    pz_w[base + 32] = dZ(ph, base + 32, cf1, edge_tag);
                }
            }
        };

        // ---- predicated sums of one row's two points (reproduce_rodal_exact.py:57-61, 220-224) ------
        // NaN points are not counted here: every point is exactly one of pos / neg / zero / NaN, so the
        // host derives cnt_nan from the number of points.  The shell sums are skipped for a warp none of
        // whose lanes lies inside any shell (adding +0.0 is the identity, so this does not change a
        // single bit of the result).
        auto tally2 = [&](const double (&e)[2], double yyj, double xxi) {
            const bool out0 = (cf0 & 16u) != 0, out1 = (cf1 & 16u) != 0;
            const bool p0 = out0 && (e[0] > 0.0), n0 = out0 && (e[0] < 0.0);
            const bool p1 = out1 && (e[1] > 0.0), n1 = out1 && (e[1] < 0.0);
            const double ep0 = p0 ? e[0] : 0.0, en0 = n0 ? e[0] : 0.0;
            const double ep1 = p1 ? e[1] : 0.0, en1 = n1 ? e[1] : 0.0;
            acc_pos += ep0; acc_neg += en0;
            acc_pos += ep1; acc_neg += en1;
            // en is e where e < 0 and +0.0 elsewhere, so its sign bit IS the "negative" flag
            c_neg += (int)((unsigned)__double2hiint(en0) >> 31) + (int)((unsigned)__double2hiint(en1) >> 31);
            // Every output point is exactly one of pos / neg / zero / NaN.  Zeros and NaNs are rare (exact cancellation,
            // overflow), so they are counted on a warp-uniform side path and the positive count is derived at the flush
            // from the number of output points of the interval.
            const bool o0 = out0 && !p0 && !n0, o1 = out1 && !p1 && !n1;
            if (__any_sync(0xffffffffu, o0 || o1)) {
                if (o0) c_zn += (e[0] == e[0]) ? 1 : 0x10000;
                if (o1) c_zn += (e[1] == e[1]) ? 1 : 0x10000;
            }
            if (NSH > 0) {
                bool in0[NSHA], in1[NSHA];
                bool any_in = false;
                if (DT == IWB_F64) {
                    const double sxy = xxi + yyj;
                    const double s0 = sxy + zz0, s1 = sxy + zz1;
#pragma unroll
                    for (int sh = 0; sh < NSH; ++sh) {
                        in0[sh] = out0 && (s0 <= P.shell_thr[sh]);
                        in1[sh] = out1 && (s1 <= P.shell_thr[sh]);
                        any_in |= in0[sh] | in1[sh];
                    }
                } else {
                    const float sxy = __fadd_rn((float)xxi, (float)yyj);
                    const float s0 = __fadd_rn(sxy, (float)zz0), s1 = __fadd_rn(sxy, (float)zz1);
#pragma unroll
                    for (int sh = 0; sh < NSH; ++sh) {
                        in0[sh] = out0 && (s0 <= (float)P.shell_thr[sh]);
                        in1[sh] = out1 && (s1 <= (float)P.shell_thr[sh]);
                        any_in |= in0[sh] | in1[sh];
                    }
                }
                if (__any_sync(0xffffffffu, any_in)) {
#pragma unroll
                    for (int sh = 0; sh < NSH; ++sh) {
                        // fma(e, m, acc) with m in {0, 1} rounds exactly like acc + (m ? e : 0): the mask multiply the
                        // reference writes (e * (e > 0) * mask), one instruction per sum instead of a 64-bit select + add
                        const double m0 = in0[sh] ? 1.0 : 0.0, m1 = in1[sh] ? 1.0 : 0.0;
                        sh_pos[sh] = fma(ep0, m0, sh_pos[sh]);
                        sh_neg[sh] = fma(en0, m0, sh_neg[sh]);
                        sh_pos[sh] = fma(ep1, m1, sh_pos[sh]);
                        sh_neg[sh] = fma(en1, m1, sh_neg[sh]);
                        if (in0[sh]) ++sh_cnt[sh];
                        if (in1[sh]) ++sh_cnt[sh];
                    }
                }
            }
        };

        // ---- task O(i): second derivatives, rho, sums at the output columns of plane i --------------------
        auto taskO = [&](int i, double xxi, auto edge_tag) {
            constexpr bool EDGE = decltype(edge_tag)::value;
            const double* const px_i = px_ring + (i & 3) * PLANE;
            const double* const px_p = px_ring + ((i + 1) & 3) * PLANE;   // phi_x(i+1)
            const double* const px_m = px_ring + ((i + 3) & 3) * PLANE;   // phi_x(i-1)
            const double* const px_m2 = px_ring + ((i + 2) & 3) * PLANE;  // phi_x(i-2) (last plane only)
            const double* const py_r = py_buf + (i & 1) * PLANE;
            const double* const pz_r = pz_buf + (i & 1) * PLANE;
            const int iedge = EDGE ? ((i == 0) ? 1 : ((i == P.n0 - 1) ? 2 : 0)) : 0;
#pragma unroll 1
            for (int jr = ty; jr <= Tj + 1; jr += NW) {       // the thread that produced phi_x of a column consumes it
                if (jr < 2) continue;
                const int base = jr * RK + tx;
                const int jedge = EDGE ? ((t_jlo && jr == 2) ? 1 : ((t_jhi && jr == Tj + 1) ? 2 : 0)) : 0;
                const double yyj = yy_s[jr];
                double rho2[2];
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const int idx = base + 32 * a;
                    const unsigned cf = a ? cf1 : cf0;
                    double pxx;
                    if (EDGE && iedge == 1)        // phi_x(0), (1), (2)
                        pxx = edge3(cx.a0, px_i[idx], cx.b0, px_p[idx], cx.c0, px_m2[idx]);
                    else if (EDGE && iedge == 2)   // phi_x(n0-3), (n0-2), (n0-1)
                        pxx = edge3(cx.a1, px_m2[idx], cx.b1, px_m[idx], cx.c1, px_i[idx]);
                    else
                        pxx = cdivm<MODE>(px_p[idx] - px_m[idx], cx.h2, cx.rh2);
                    const double pxy = dY(px_i, idx, jedge, edge_tag);
                    const double pxz = dZ(px_i, idx, cf, edge_tag);
                    const double pyy = dY(py_r, idx, jedge, edge_tag);
                    const double pyz = dZ(py_r, idx, cf, edge_tag);
                    const double pzz = dZ(pz_r, idx, cf, edge_tag);
                    // K_ij = -phi_ij: the sign cancels in K^2 and in K_ij K^ij (reproduce_rodal_exact.py:198-212)
                    const double tr = (pxx + pyy) + pzz;
                    const double diag = (pxx * pxx + pyy * pyy) + pzz * pzz;
                    const double off = (pxy * pxy + pxz * pxz) + pyz * pyz;
                    const double kk = fma(2.0, off, diag);     // 2.0*off is exact -> same rounding as diag + 2.0*off
                    rho2[a] = cdivm<MODE>(tr * tr - kk, P.c16pi, P.rc16pi);
                }
                if (EMIT) {
                    const long long j = j0 - 2 + jr, k = k0 - 2 + tx;
                    double* dst = P.rho_out + ((long long)(i - P.i_begin) * P.n1 + j) * P.n2 + k;
                    if (cf0 & 16u) dst[0] = rho2[0];
                    if (cf1 & 16u) dst[32] = rho2[1];
                }
                const double e2[2] = {rho2[0] * P.dV, rho2[1] * P.dV};
                tally2(e2, yyj, xxi);
            }
        };

        // ---- block-level flush of the register sums into partial[fb][tile] ----------------
        auto flush = [&](int fb, int i_last) {
            constexpr int NV = 6 + 3 * NSH;    // values actually reduced
            double vals[NV];
            vals[0] = acc_pos; vals[1] = acc_neg;
            vals[2] = __longlong_as_double(0LL);      // cnt_pos: derived below from the other three
            vals[3] = __longlong_as_double((long long)c_neg);
            vals[4] = __longlong_as_double((long long)(c_zn & 0xffff));
            vals[5] = __longlong_as_double((long long)(c_zn >> 16));
#pragma unroll
            for (int s = 0; s < NSH; ++s) {
                vals[6 + s] = sh_pos[s];
                vals[6 + NSH + s] = sh_neg[s];
                vals[6 + 2 * NSH + s] = __longlong_as_double((long long)sh_cnt[s]);
            }
            // warp butterfly (fixed order), integers added as integers
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const bool is_int = (q >= 2 && q < 6) || (q >= 6 + 2 * NSH);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    double o = __shfl_xor_sync(0xffffffffu, vals[q], off);
                    if (is_int) vals[q] = __longlong_as_double(__double_as_longlong(vals[q]) + __double_as_longlong(o));
                    else vals[q] += o;
                }
            }
            if (tx == 0) {
#pragma unroll
                for (int q = 0; q < NV; ++q) scratch[ty * NV + q] = vals[q];
            }
            __syncthreads();
            if (threadIdx.x < ROW) {
                // destination slot -> index in vals (or -1: slot of an unused shell, written as zero)
                const int slot = threadIdx.x;
                int q;
                if (slot < 6) q = slot;
                else {
                    const int grp = (slot - 6) / IWB_MAX_SHELLS, sidx = (slot - 6) % IWB_MAX_SHELLS;
                    q = (sidx < NSH) ? 6 + grp * NSH + sidx : -1;
                }
                double acc = 0.0;
                if (q == 2) {
                    // positive points = output points of the interval - (negative + zero + NaN)
                    long long rest = 0;
                    for (int wv = 0; wv < NW; ++wv)
                        rest += __double_as_longlong(scratch[wv * NV + 3]) + __double_as_longlong(scratch[wv * NV + 4]) +
                                __double_as_longlong(scratch[wv * NV + 5]);
                    acc = __longlong_as_double((long long)Tj * Tk * (i_last - flush_from + 1) - rest);
                } else if (q >= 0) {
                    const bool is_int = (q >= 2 && q < 6) || (q >= 6 + 2 * NSH);
                    acc = scratch[q];
                    for (int wv = 1; wv < NW; ++wv) {
                        double o = scratch[wv * NV + q];
                        if (is_int) acc = __longlong_as_double(__double_as_longlong(acc) + __double_as_longlong(o));
                        else acc += o;
                    }
                }
                P.partial[((size_t)fb * ntiles + tile) * ROW + slot] = acc;
            }
            // the scratch is rewritten at the next flush only, at least one step barrier from here
            acc_pos = 0.0; acc_neg = 0.0; c_neg = c_zn = 0;
            flush_from = i_last + 1;
#pragma unroll
            for (int s = 0; s < NSHA; ++s) { sh_pos[s] = 0.0; sh_neg[s] = 0.0; sh_cnt[s] = 0; }
        };

        auto runP = [&](int p, int s3, double xp, double xxp) {
            if (tile_edge || (p == 2 && pstart == 0)) taskP(p, s3, xp, xxp, std::true_type{});
            else taskP(p, s3, xp, xxp, std::false_type{});
        };
        auto runC = [&](int q, int s3) {
            if (tile_edge) taskC(q, s3, std::true_type{});
            else taskC(q, s3, std::false_type{});
        };
        auto runO = [&](int i, double xxi) {
            if (IWB_DBG & 1) return;
            if (tile_edge || i == 0 || i == P.n0 - 1) taskO(i, xxi, std::true_type{});
            else taskO(i, xxi, std::false_type{});
        };
        // phi_x(n0-1) may be written once every phi plane exists and its ring slot (that of phi_x(n0-5),
        // last read by O(n0-4)) is dead, i.e. from the step of output plane n0-3 on; it is first needed
        // by O(n0-2) (n0 == 3: by every plane, and 0 >= n0-3 holds).
        auto px_last_ok = [&](int i) { return i >= P.n0 - 3; };

        // ---- prologue: everything O(i0) needs ----------------------------------------------------------
        __syncthreads();                 // yy_s visible
        int produced = pstart - 1;
        bool px_last_done = false;
        {
            const int upto = min(i0 + 2, P.n0 - 1);
            while (produced < upto) { ++produced; runP(produced, produced % 3, P.x[produced], P.xx[produced]); }
            __syncthreads();
            runC(i0, i0 % 3);
            if (i0 == 0 && P.n0 > 3) {   // O(0) needs phi_x(2), i.e. phi(3), which reuses the slot of phi(0)
                __syncthreads();
                ++produced;
                runP(produced, produced % 3, P.x[produced], P.xx[produced]);
            }
            if (produced == P.n0 - 1 && px_last_ok(i0)) { taskP_last(); px_last_done = true; }
            __syncthreads();
        }

        // ---- march: one (split) barrier per output plane -----------------------------------------------
        const int plim = min(P.n0 - 1, i1 + 1);          // phi(i1+1) is the last plane this item needs
        flush_from = i0;
        int m3 = i0 % 3;                                 // i mod 3, kept incrementally (no division in the loop)
        for (int i = i0; i < i1; ++i) {
            const bool doP = produced < min(i + 3, plim);
            const bool doL = !doP && !px_last_done && produced == P.n0 - 1 && px_last_ok(i);
            const bool doC = (i + 1 < i1);
            // coordinates of this step's planes, requested before the tasks that hide their latency
            const int pn = min(produced + 1, P.n0 - 1);
            const double xp_n = P.x[pn], xxp_n = P.xx[pn], xxi = P.xx[i];
            const int m3p1 = (m3 == 2) ? 0 : m3 + 1;
            // What OTHER warps' next step needs from this warp is O(i) (they rewrite the phi_y/phi_z buffer it reads)
            // and C(i+1) (phi_y/phi_z(i+1) for their O(i+1); phi(i+1) dead for their P(i+4)).  Task P touches only ring
            // entries of the own columns until two steps later (tasks P and O assign rows to warps identically), so it
            // runs between the arrival and the wait: the barrier latency and the imbalance of O/C hide behind it.
            runO(i, xxi);
            if (doC) runC(i + 1, m3p1);
            step_arrive();
            if (doP) {
                // steady state: produced + 1 == i + 3, whose slot is i mod 3; otherwise (start / end of the axis) compute it
                const int pp = produced + 1;
                runP(pp, (pp == i + 3) ? m3 : pp % 3, xp_n, xxp_n);
            }
            if (doL) taskP_last();
            step_wait();
            if (doP) ++produced;
            if (doL) px_last_done = true;
            if (((i + 1) & (FLUSH - 1)) == 0 || i == i1 - 1) flush(i / FLUSH, i);
            m3 = m3p1;
        }
    }
}

// ---- fixed-order sum over tiles: partial[fb][tile][ROW] -> table[fb][ROW] ----------------------
// (static: every translation unit that includes this header gets its own copy)
static __global__ void __launch_bounds__(128) k_reduce_tiles(const double* __restrict__ partial,
                                                             double* __restrict__ table, int fb_begin, int ntiles,
                                                             unsigned int* counter = nullptr)
{
    if (counter && blockIdx.x == 0 && threadIdx.x == 0) *counter = 0u;    // re-arm the work queue of k_energy3d
    const int fb = fb_begin + blockIdx.x;
    const int q = threadIdx.x;            // one thread per row slot
    if (q >= ROW) return;
    const bool is_int = (q >= 2 && q < 6) || (q >= 6 + 2 * IWB_MAX_SHELLS);
    const double* src = partial + (size_t)fb * ntiles * ROW + q;
    // 4 interleaved chains (tile t goes to chain t mod 4), then a fixed combine: deterministic
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    long long i0 = 0, i1 = 0, i2 = 0, i3 = 0;
    int t = 0;
    for (; t + 4 <= ntiles; t += 4) {
        const double v0 = src[(size_t)(t + 0) * ROW], v1 = src[(size_t)(t + 1) * ROW];
        const double v2 = src[(size_t)(t + 2) * ROW], v3 = src[(size_t)(t + 3) * ROW];
        if (is_int) {
            i0 += __double_as_longlong(v0); i1 += __double_as_longlong(v1);
            i2 += __double_as_longlong(v2); i3 += __double_as_longlong(v3);
        } else { a0 += v0; a1 += v1; a2 += v2; a3 += v3; }
    }
    if (t < ntiles) { const double v = src[(size_t)t * ROW]; if (is_int) i0 += __double_as_longlong(v); else a0 += v; ++t; }
    if (t < ntiles) { const double v = src[(size_t)t * ROW]; if (is_int) i1 += __double_as_longlong(v); else a1 += v; ++t; }
    if (t < ntiles) { const double v = src[(size_t)t * ROW]; if (is_int) i2 += __double_as_longlong(v); else a2 += v; ++t; }
    table[(size_t)fb * ROW + q] = is_int ? __longlong_as_double((i0 + i1) + (i2 + i3)) : ((a0 + a1) + (a2 + a3));
}

// ---- fixed-order sum over flush blocks: table[nrows][ROW] -> out[ROW] ------------------------------
static __global__ void __launch_bounds__(128) k_reduce_rows(const double* __restrict__ table, double* __restrict__ out, int nrows)
{
    const int q = threadIdx.x;
    if (q >= ROW) return;
    const bool is_int = (q >= 2 && q < 6) || (q >= 6 + 2 * IWB_MAX_SHELLS);
    double a = 0.0;
    long long ia = 0;
    for (int r = 0; r < nrows; ++r) {
        double v = table[(size_t)r * ROW + q];
        if (is_int) ia += __double_as_longlong(v);
        else a += v;
    }
    out[q] = is_int ? __longlong_as_double(ia) : a;
}

}  // namespace iwb

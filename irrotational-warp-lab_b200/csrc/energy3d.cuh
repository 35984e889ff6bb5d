// Fused 3D energy-condition kernel: phi -> grad phi -> grad grad phi -> K_ij -> rho -> sums.
//
// Replaces the array pipeline of compute_energy_3d
// (/root/reference/scripts/reproduce_rodal_exact.py:171-224) and phi_exact_rodal /
// g_rodal_exact (/root/reference/src/irrotational_warp/potential.py:37-99); no
// intermediate ever reaches HBM.
//
// Decomposition.  The volume is C-order [i, j, k] (k contiguous).  A work unit is
// (tile of the (j, k) plane) x (flush block of IWB_FLUSH_PLANES planes of i); a CTA takes spans of
// consecutive units and marches along i (see "Work distribution" in the kernel).
// The CTA's region is the tile plus a 2-point halo: RJ rows (j) x 64 columns (k), lanes along k
// (tiles at the two k faces need no halo beyond the face and hold 62 instead of 60 columns).
// Warp w owns region rows jr = w, w+NW, ...; lane tx owns columns tx, tx+32.
//
// Shared memory, 11 planes of RJ x 64 doubles:
//     phi ring [3]   (plane q in slot q mod 3)
//     phi_x ring [4] (plane q in slot q mod 4)
//     phi_y, phi_z   double-buffered (plane q in buffer q mod 2)
// The rings are sized so that ONE barrier per output plane suffices, and that barrier is split
// (mbarrier arrive ... wait).  Per output plane i every warp runs
//     O(i) + C(i+1): second derivatives, rho, sums of plane i; phi_y(i+1), phi_z(i+1)   -- smem heavy, one row loop
//     -- arrive --
//     P(i+3): evaluate phi(i+3) at the own columns and finish phi_x(i+2)                -- FP64 heavy, ~no smem
//     -- wait --
// P touches only the own columns' ring entries until two steps later, so it is legal between the
// arrival and the wait; warps drift apart by up to one task, and the shared-memory-bound tasks of
// some warps overlap the FP64-bound task of others (measured: 59 -> 77 Gpoints/s over the version
// with two __syncthreads per plane, whose smem and FP64 phases alternated in lock step).
// The tasks address shared memory explicitly (one laundered base address per thread, CTA-uniform plane
// offsets, immediates for the neighbours) and evaluate the central formulas everywhere; rows / columns /
// planes on a global face then override the affected derivatives with the one-sided formulas.
// Row loops are deliberately NOT unrolled: the IEEE division / sqrt sequences are long and an unrolled
// body overflows the instruction cache (task P alone handles both rows of a warp at once, for ILP).
// Sums are kept in registers and flushed (warp shuffle + fixed-order block tree) once per
// IWB_FLUSH_PLANES planes into partial[flush block][tile slot][row]; a second kernel adds the tiles in
// fixed order (16 chains + binary tree).  Flush blocks are aligned to the GLOBAL plane index and a tile
// shard owns whole chains, so the result is bit-identical for any slab decomposition along i and any
// interleaved tile sharding (DESIGN.md sections 4.2, 5).
#pragma once
#include <type_traits>

#include "iwb_device.cuh"
#include "../../include/iwb200.h"

#ifndef IWB_BARRIER_BACKOFF_NS
#define IWB_BARRIER_BACKOFF_NS 100
#endif

#ifndef IWB_OC_UNROLL
#define IWB_OC_UNROLL 1     // both rows of the O + C loop unrolled: no loop overhead, row b's loads overlap row a's tally (+0.9 %)
#endif
#ifndef IWB_P4
#define IWB_P4 1       // task P evaluates both rows of a warp at once: four independent phi evaluations per thread
#endif

namespace iwb {

constexpr int RK = 64;              // region columns (k): two per lane
constexpr int TK_MAX = RK - 4;      // tile columns
constexpr int ROW = IWB_ROW_WIDTH;  // doubles per partial row
constexpr int FLUSH = IWB_FLUSH_PLANES;
constexpr int SMEM_PAD = 8;         // guard doubles before and after the planes (lane 0 / lane 31 reads of idx -+ 1)
constexpr int NPLANES = 11;
constexpr int NEDGE = 18;           // one-sided coefficients a0 b0 c0 a1 b1 c1 of the three axes (shared-memory copy)
// Fused radial profile (k_energy3d_prof): a work unit (tile x 16 planes) spans a few length units of r, i.e. a handful
// of bins, so every warp keeps a WINDOW of PROF_WIN consecutive bins (sums + counts) in its slice of the flush scratch.
constexpr int PROF_WIN = 14;                 // bins per window: 14 sums + 14 counts (7 doubles) <= ROW doubles per warp
constexpr int PROF_FUSED_MAX_BINS = 144;     // thresholds of all bins live in shared memory
constexpr int PROF_THR_S = PROF_FUSED_MAX_BINS + 1 + PROF_WIN + 1;   // thresholds + infinite padding behind the last edge
static_assert(PROF_WIN + (PROF_WIN + 1) / 2 <= IWB_ROW_WIDTH, "a warp's profile window must fit its slice of the flush scratch");

struct K3Params {
    int n0, n1, n2;
    int i_begin, i_end;            // planes to evaluate (global indices)
    int ntj, ntk;                  // tiles of the (j, k) plane
    int nfb_slab;                  // flush blocks (FLUSH planes each) in [i_begin, i_end)
    int tile_first, tile_stride;   // this launch covers the tiles of schedule index tile_first + m * tile_stride
    int nspans;                    // work spans, handed out dynamically in index order
    const int* span_begin;         // [nspans + 1] unit boundaries; unit u = m * nfb_slab + flush block
    const double* x;               // [n0]  (float32 values widened for the f32ref path)
    const double* xx;              // [n0]  x*x   (rounded in the path's coordinate dtype)
    const double* yy;              // [n1]  y*y
    const double* zz;              // [n2]  z*z
    PhiConst pc;
    AxisCoef cx, cy, cz;
    double dV;                     // (dx*dy)*dz
    double c16pi, rc16pi;          // 16.0*pi and RN(1/(16.0*pi))
    double shell_thr[IWB_MAX_SHELLS];   // r <= R  <=>  s <= shell_thr (s = r*r before sqrt), see host
    double shell_thr_max;               // largest of them (row test of the shell sums)
    double* partial;               // [nfb_total][ntiles][ROW]
    double* rho_out;               // NULL or [i_end-i_begin][n1][n2]
    unsigned int* counter;         // dynamic work-item counter (0 at launch; reset by k_reduce_tiles)
    int i_zero, j_zero;            // index of the coordinate that is exactly 0 on axis 0 / 1 (-1: none): r == 0 is possible there
    // fused radial profile (k_energy3d_prof only; see radial_profile.cuh for the binning rule)
    const double* prof_thr;        // [prof_nb + 1] thresholds on s = r*r before the square root: r >= edges[q] <=> s >= prof_thr[q]
    int prof_nb;                   // bins (<= PROF_FUSED_MAX_BINS)
    double* prof_sum;              // [nfb_total][ntiles][PROF_WIN] per-unit window sums of rho
    int* prof_cnt;                 // [nfb_total][ntiles][PROF_WIN] per-unit window point counts
    int* prof_base;                // [nfb_total][ntiles] first bin of the unit's window
    unsigned int* prof_overflow;   // set when a unit touches a bin beyond its window (the host then takes the two-pass path)
};

__host__ __device__ inline int tile_begin(int t, int nt, int n) { return (int)(((long long)t * n) / nt); }

// Tiles along one axis of the (j, k) plane for a region of R points: an interior tile holds at most R - 4 output points
// (2-point halo on either side), but the tile at a global face needs no halo beyond the face and holds R - 2 -- a single
// tile R.  n = 1024 columns are 17 tiles (2 x 62 + 15 x 60) instead of 18.  tiles_1d: number of tiles; tile_begin_1d:
// first point of tile t (t = nt: n), the slack spread evenly over the tiles.
__host__ __device__ inline int tiles_1d(int n, int R)
{
    if (n <= R) return 1;
    if (n <= 2 * (R - 2)) return 2;
    return 2 + (n - 2 * (R - 2) + (R - 4) - 1) / (R - 4);
}
__host__ __device__ inline int tile_begin_1d(int t, int nt, int n, int R)
{
    if (t <= 0) return 0;
    if (t >= nt) return n;
    const int slack = 2 * (R - 2) + (nt - 2) * (R - 4) - n;
    return (R - 2) + (t - 1) * (R - 4) - (int)(((long long)t * slack) / nt);
}
__host__ __device__ inline int tiles_k(int n) { return tiles_1d(n, RK); }
__host__ __device__ inline int tile_begin_k(int t, int nt, int n) { return tile_begin_1d(t, nt, n, RK); }
// Rows (j) of a region of RJ rows: an even split into tiles of at most RJ - 4 output rows.  (Face tiles of RJ - 2 rows,
// as for the columns, were measured twice: one tile row less at n = 256 / 400 (+2 %), but -4 % at n = 1024, where the
// tile count does not change and the row offset costs the hot loops registers -- profiles/r2_k_energy3d_history.txt.)
__host__ __device__ inline int tiles_j(int n, int RJ) { return (n + (RJ - 4) - 1) / (RJ - 4); }
__host__ __device__ inline int tile_begin_j(int t, int nt, int n) { return tile_begin(t, nt, n); }

template <int RJ, int NW>
struct Geo {
    static constexpr int PLANE = RJ * RK;
    static constexpr int TJ_MAX = RJ - 4;
    static constexpr int NT = NW * 32;
    static constexpr int SCRATCH = NW * ROW;     // block-reduction scratch (doubles)
    static constexpr int NROWTAB = 2;            // per-row tables: y*y, and the largest x*x for which the row can hold a shell point
    static constexpr size_t SMEM_BYTES = (size_t(NPLANES) * PLANE + 2 * SMEM_PAD + NROWTAB * RJ + SCRATCH + NEDGE) * sizeof(double);
};

// The kernel body; PROF adds the fused radial profile (k_energy3d_prof below), everything else is k_energy3d.
template <int DT, int NSH, int RJ, int NW, bool EMIT, int MODE, bool PROF>
__device__ __forceinline__ void energy3d_body(const K3Params& P)
{
    using G = Geo<RJ, NW>;
    constexpr int PLANE = G::PLANE;
    constexpr int NSHA = NSH > 0 ? NSH : 1;
    extern __shared__ double smem_raw[];
    double* const phi_ring = smem_raw + SMEM_PAD;          // 3 planes
    double* const px_ring = phi_ring + 3 * PLANE;          // 4 planes
    double* const py_buf = phi_ring + 7 * PLANE;           // 2 planes
    double* const pz_buf = phi_ring + 9 * PLANE;           // 2 planes
    double* const yy_s = phi_ring + NPLANES * PLANE + SMEM_PAD;   // [RJ] y*y of the region rows, then [RJ] shell-test limits on x*x
    double* const scratch = yy_s + G::NROWTAB * RJ;        // [NW][ROW]
    // One-sided (global face) coefficients live in shared memory: only the edge flavours of the tasks read them,
    // and as kernel parameters they would occupy 36 uniform registers that the interior loops need for their constants.
    double* const ec_s = scratch + G::SCRATCH;             // [3][6]: x, y, z -> a0 b0 c0 a1 b1 c1
    double* const thr_s = ec_s + NEDGE;                    // PROF: [PROF_THR_S] bin thresholds on s, +inf behind the last edge
    __shared__ int s_b0;                                   // PROF: first bin of the current window

    const int tx = threadIdx.x & 31;
    const int ty = threadIdx.x >> 5;
    const int ntiles = P.ntj * P.ntk;
    const AxisCoef cx = P.cx, cy = P.cy, cz = P.cz;
    const PhiConst pc = P.pc;

    // Interior (hot) loops address shared memory explicitly: byte address of this thread's first row in plane 0 of
    // the phi ring, plus CTA-uniform plane offsets, plus immediates (iwb_device.cuh: lds_f64 / sts_f64).
    constexpr unsigned PLB = PLANE * 8u;                   // bytes per plane
    constexpr unsigned PXB = 3u * PLB, PYB = 7u * PLB;     // phi_x ring, phi_y buffers (phi_z buffers: PYB + 2 PLB)
    constexpr int ROWB = NW * RK * 8;                      // bytes from a thread's row to its next row
    const unsigned a_row0 = keep_u32((unsigned)__cvta_generic_to_shared(phi_ring) + (unsigned)(ty * RK + tx) * 8u);
    const unsigned yy_row0 = keep_u32((unsigned)__cvta_generic_to_shared(yy_s) + (unsigned)ty * 8u);

    __shared__ int s_item;
    // Split (arrive / wait) barrier of the march: one mbarrier phase per output plane.
    __shared__ unsigned long long s_mbar;
    const unsigned mbar_addr = (unsigned)__cvta_generic_to_shared(&s_mbar);
    unsigned mbar_parity = 0;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_addr), "r"(NW * 32));
    if (threadIdx.x < NEDGE) {
        // AxisCoef = {h2, rh2, a0, b0, c0, a1, b1, c1}; cx, cy, cz are consecutive members of K3Params
        static_assert(sizeof(AxisCoef) == 8 * sizeof(double), "AxisCoef layout");
        const int axis = threadIdx.x / 6, q = threadIdx.x - 6 * axis;
        ec_s[threadIdx.x] = (&P.cx.h2)[8 * axis + 2 + q];
    }
    if (PROF) {
        for (int q = threadIdx.x; q < PROF_THR_S; q += NW * 32) thr_s[q] = (q <= P.prof_nb) ? P.prof_thr[q] : INFINITY;
    }
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
    // Work distribution.  A unit is (tile, flush block); units are ordered tile-major, and a SPAN is a contiguous range
    // of units.  A CTA marches through the units of a span without interruption -- the rings stay primed, so only the
    // first unit of a tile within a span pays the 4-plane phi prologue -- and spans are handed out in index order by an
    // atomic counter: one large span per CTA first (70 % of the work, no scheduling tail inside it), then ever
    // smaller ones that even out the differences (host: make_spans).  Partial sums go to fixed slots, so the result
    // does not depend on which CTA ran which span.
    for (;;) {
    __syncthreads();                     // s_item of the previous span is dead
    if (threadIdx.x == 0) s_item = (int)atomicAdd(P.counter, 1u);
    __syncthreads();
    const int span = s_item;
    if (span >= P.nspans) break;
    int unit = P.span_begin[span];
    const int unit_end = P.span_begin[span + 1];
    while (unit < unit_end) {
        __syncthreads();                 // previous item's shared memory is dead
        const int rnk_local = unit / P.nfb_slab;
        const int fblk0 = unit - rnk_local * P.nfb_slab;
        const int rnk = P.tile_first + rnk_local * P.tile_stride;      // schedule index of the tile
        const int fblk1 = min(P.nfb_slab, fblk0 + (unit_end - unit));
        unit += fblk1 - fblk0;
        // Tiles that touch a global face cost more than interior tiles (one-sided formulas, divergent lanes); they come
        // first in the unit order, i.e. in the large spans.
        // rank -> tile: row tj = 0, row tj = ntj-1, the two edge columns of the rows between, then the interior.
        int tile;
        {
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
        const int j0 = tile_begin_j(tj, P.ntj, P.n1), Tj = tile_begin_j(tj + 1, P.ntj, P.n1) - j0;
        const int k0 = tile_begin_k(tk, P.ntk, P.n2), Tk = tile_begin_k(tk + 1, P.ntk, P.n2) - k0;
        const int hl = (k0 == 0) ? 0 : 2;           // halo columns on the low side: region column c is k = k0 - hl + c
        const int i0 = P.i_begin + fblk0 * FLUSH;
        const int i1 = min(P.i_end, P.i_begin + fblk1 * FLUSH);

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
            const int ka = k0 - hl + tx, kb = ka + 32;
            zz0 = P.zz[min(max(ka, 0), P.n2 - 1)];
            zz1 = P.zz[min(max(kb, 0), P.n2 - 1)];
            cf0 = ((ka == 0) ? 2u : 0u) | ((ka == P.n2 - 1) ? 4u : 0u) | ((tx >= hl && tx < hl + Tk) ? 16u : 0u);
            cf1 = ((kb == 0) ? 2u : 0u) | ((kb == P.n2 - 1) ? 4u : 0u) | ((tx + 32 < hl + Tk) ? 16u : 0u);
        }

        // all-ones / zero words of the two columns: AND-mask of the energy density at non-output columns
        const int om0 = (int)keep_u32((cf0 & 16u) ? 0xffffffffu : 0u), om1 = (int)keep_u32((cf1 & 16u) ? 0xffffffffu : 0u);
        // smallest z*z among the tile's output columns (row test of the shell sums); every warp computes the same value
        double zzmin_s;
        {
            double m = fmin((cf0 & 16u) ? zz0 : INFINITY, (cf1 & 16u) ? zz1 : INFINITY);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, off));
            zzmin_s = m;
        }
        if (NSH > 0 && threadIdx.x < RJ) {
            // Row table of the shell test.  s = RN(RN(x*x + y*y) + z*z) is monotone in each term, so for a row (fixed y) of
            // this tile  x*x > xout[row]  means that even the column with the smallest z*z lies outside the largest shell:
            // the row skips the shell arithmetic after ONE comparison (round 1 formed (x*x + y*y) + zz_min per row).  The
            // limit is conservative -- a relative margin far above the rounding errors of s (float64 2^-40, float32 2^-18)
            // -- so a row it leaves undecided simply takes the exact per-point comparisons.
            const double yyv = P.yy[min(max(j0 - 2 + (int)threadIdx.x, 0), P.n1 - 1)];
            const double mrg = (DT == IWB_F64) ? 0x1p-40 : 0x1p-18;
            const double tm = P.shell_thr_max;
            yy_s[RJ + threadIdx.x] = (tm < 0.0) ? -1.0 : ((tm - yyv) - zzmin_s) + ((tm + yyv) + zzmin_s) * mrg;
        }

        // ---- accumulators (registers) -------------------------------------------------
        double acc_sum = 0.0, acc_abs = 0.0;        // sum(e), sum(|e|) over the thread's output points
        double sh_sum[NSHA], sh_abs[NSHA];
        int cnt3 = 0;            // 8-bit fields: negative | positive << 8 | zero (incl. masked columns) << 16
        int flush_from = 0;      // first plane of the running flush interval (set below)
        int sh_cnt[NSHA];
#pragma unroll
        for (int s = 0; s < NSHA; ++s) { sh_sum[s] = 0.0; sh_abs[s] = 0.0; sh_cnt[s] = 0; }
        static_assert(2 * ((RJ + NW - 1) / NW) * FLUSH <= 255, "count fields of cnt3 are 8 bits wide");

        const int w0 = min(max(i0 - 1, 0), P.n0 - 3);
        const int pstart = max(0, w0 - 1);

        // ---- fused radial profile (PROF): per-warp window of PROF_WIN bins in the warp's slice of the flush scratch ----
        int pb = 0;                    // window-relative bins of the thread's four points at the previous plane, 8 bits each
        bool beyond_seen = false;      // an output point of this thread lay beyond the window
        int b0_reg = 0;                // first bin of the current window (copy of s_b0)
        double* const win_sum = scratch + ty * ROW;                            // [PROF_WIN] sums of rho ...
        int* const win_cnt = reinterpret_cast<int*>(win_sum + PROF_WIN);       // ... and [PROF_WIN] point counts
        // Window of the flush block with planes [ia, ib): its first bin is the bin of the smallest s the block can hold,
        // s_min = (min x*x + min y*y) + min z*z over the block's planes and the tile's output rows / columns (rounding is
        // monotone, so every point of the block has s >= s_min).  Called by warp 0; publishes s_b0.
        auto prof_rebase = [&](int ia, int ib) {
            double m = INFINITY, ym = INFINITY;
            for (int q = ia + tx; q < ib; q += 32) m = fmin(m, P.xx[q]);
            for (int jr = 2 + tx; jr <= Tj + 1; jr += 32) ym = fmin(ym, yy_s[jr]);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                m = fmin(m, __shfl_xor_sync(0xffffffffu, m, off));
                ym = fmin(ym, __shfl_xor_sync(0xffffffffu, ym, off));
            }
            const double smin = (DT == IWB_F64) ? (m + ym) + zzmin_s
                                                : (double)__fadd_rn(__fadd_rn((float)m, (float)ym), (float)zzmin_s);
            int below = 0;             // thresholds <= s_min (they are sorted)
            for (int q0 = 0; q0 <= P.prof_nb; q0 += 32) {
                const int q = q0 + tx;
                below += __popc(__ballot_sync(0xffffffffu, q <= P.prof_nb && thr_s[q] <= smin));
            }
            if (tx == 0) s_b0 = max(below - 1, 0);
        };
        // Keys (window-relative bin, 255: in no bin) and rho of the warp's rows of one plane, filled by the tally; then ONE
        // reduction for all of them: per row the three bins from the row's lowest one are summed over the lanes by
        // independent shuffle trees (six dependent chains in flight instead of one after the other), lane 0 adds the
        // results to the warp's window in fixed order; lanes further up -- rare: a row of 64 columns seldom spans more than
        // three bins -- are handled bin by bin afterwards.  The order of every addition is fixed.
        constexpr int NROWS = (RJ + NW - 1) / NW, PROF_KC = 3;
        int pkey[NROWS][2];
        double prho[NROWS][2];
        auto prof_clear = [&]() {
#pragma unroll
            for (int it = 0; it < NROWS; ++it) { pkey[it][0] = 255; pkey[it][1] = 255; prho[it][0] = 0.0; prho[it][1] = 0.0; }
        };
        auto prof_reduce = [&]() {
            int lo[NROWS];
            double t[NROWS][PROF_KC];
            int c[NROWS][PROF_KC];
#pragma unroll
            for (int it = 0; it < NROWS; ++it) lo[it] = __reduce_min_sync(0xffffffffu, min(pkey[it][0], pkey[it][1]));
#pragma unroll
            for (int it = 0; it < NROWS; ++it) {
#pragma unroll
                for (int k = 0; k < PROF_KC; ++k) {
                    const bool m0 = pkey[it][0] == lo[it] + k && lo[it] != 255, m1 = pkey[it][1] == lo[it] + k && lo[it] != 255;
                    t[it][k] = (m0 ? prho[it][0] : 0.0) + (m1 ? prho[it][1] : 0.0);
                    c[it][k] = __popc(__ballot_sync(0xffffffffu, m0)) + __popc(__ballot_sync(0xffffffffu, m1));
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int it = 0; it < NROWS; ++it) {
#pragma unroll
                    for (int k = 0; k < PROF_KC; ++k) t[it][k] += __shfl_xor_sync(0xffffffffu, t[it][k], off);
                }
            }
            if (tx == 0) {
#pragma unroll
                for (int it = 0; it < NROWS; ++it) {
#pragma unroll
                    for (int k = 0; k < PROF_KC; ++k)
                        if (c[it][k]) { win_sum[lo[it] + k] += t[it][k]; win_cnt[lo[it] + k] += c[it][k]; }
                }
            }
#pragma unroll
            for (int it = 0; it < NROWS; ++it) {
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const bool left = pkey[it][a] != 255 && pkey[it][a] >= lo[it] + PROF_KC;
                    unsigned todo = __ballot_sync(0xffffffffu, left);
                    while (todo) {
                        const int b = __shfl_sync(0xffffffffu, pkey[it][a], __ffs(todo) - 1);
                        const bool mine = left && pkey[it][a] == b;
                        const unsigned same = __ballot_sync(0xffffffffu, mine);
                        double v = mine ? prho[it][a] : 0.0;
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                        if (tx == 0) { win_sum[b] += v; win_cnt[b] += __popc(same); }
                        todo &= ~same;
                    }
                }
            }
        };

        // ---- sums of one row's two points (reproduce_rodal_exact.py:57-61, 220-224) ---------------------
        // e is masked to +0.0 at the non-output columns (an AND with the column's all-ones / zero word), and the
        // accumulators are sum(e) and sum(|e|): E+ = (sum + sum|.|)/2, E- = (sum - sum|.|)/2, formed per thread at the
        // flush.  The shell sums use the reference's own mask multiply, fma(e, m, acc) with m in {0, 1}.  A NaN or an
        // infinite point therefore poisons exactly the sums it poisons in the reference (NaN * 0 = inf * 0 = NaN).
        // Counts: every point is exactly one of neg / pos / zero / NaN; the first three are counted here in 8-bit
        // fields of one register (<= 64 points per thread and flush interval), NaN is derived by the host.  The zero
        // field also counts the masked non-output columns; the flush subtracts their (known) number.
        auto tally2 = [&](const double (&e)[2], const double (&rho)[2], unsigned ya, double xxi, int it) {
            if (PROF) {
                // Radial profile: bin q holds the points with edges[q] <= r < edges[q+1], i.e. thr[q] <= s < thr[q+1]
                // (compared as integers: s >= +0).  A point's bin rarely moves by more than one from plane to plane, so
                // the search starts at the bin it had one plane ago: one branch-free step, loops only when some lane of
                // the warp needs more.  The row's keys and values are kept for prof_reduce (after both rows of the warp).
                const double yyp = lds_f64<0>(ya);
                double sp[2];
                if (DT == IWB_F64) {
                    const double sxy = xxi + yyp;
                    sp[0] = sxy + zz0; sp[1] = sxy + zz1;
                } else {
                    const float sxy = __fadd_rn((float)xxi, (float)yyp);
                    sp[0] = (double)__fadd_rn(sxy, (float)zz0); sp[1] = (double)__fadd_rn(sxy, (float)zz1);
                }
                const long long* const tw = reinterpret_cast<const long long*>(thr_s) + b0_reg;
                const int om[2] = {om0, om1};
                int rel[2] = {(pb >> (16 * it)) & 0xff, (pb >> (16 * it + 8)) & 0xff};
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const long long sb = __double_as_longlong(sp[a]);
                    int r = rel[a];
                    long long tlo = tw[r], thi = tw[r + 1];
                    r += ((sb >= thi && r < PROF_WIN - 1) ? 1 : 0) - ((sb < tlo && r > 0) ? 1 : 0);
                    tlo = tw[r]; thi = tw[r + 1];
                    const bool unsettled = (sb < tlo && r > 0) || (sb >= thi && r < PROF_WIN - 1);
                    if (__any_sync(0xffffffffu, unsettled)) {
                        while (r > 0 && sb < tw[r]) --r;
                        while (r < PROF_WIN - 1 && sb >= tw[r + 1]) ++r;
                        tlo = tw[r]; thi = tw[r + 1];
                    }
                    rel[a] = r;
                    const bool beyond = sb >= thi;                 // only with r == PROF_WIN - 1
                    const bool inb = om[a] != 0 && sb >= tlo && !beyond;
                    beyond_seen = beyond_seen || (beyond && om[a] != 0);
                    pkey[it][a] = inb ? r : 255;
                    prho[it][a] = rho[a];
                }
                pb = (pb & ~(0xffff << (16 * it))) | ((rel[0] | (rel[1] << 8)) << (16 * it));
            }
            const int hm0 = __double2hiint(e[0]) & om0, hm1 = __double2hiint(e[1]) & om1;
            const double em0 = __hiloint2double(hm0, __double2loint(e[0]) & om0);
            const double em1 = __hiloint2double(hm1, __double2loint(e[1]) & om1);
            acc_sum += em0; acc_abs += fabs(em0);
            acc_sum += em1; acc_abs += fabs(em1);
            if (em0 < 0.0) cnt3 += 1;
            if (em1 < 0.0) cnt3 += 1;
            if (em0 > 0.0) cnt3 += 0x100;
            if (em1 > 0.0) cnt3 += 0x100;
            if (((hm0 & 0x7fffffff) | __double2loint(em0)) == 0) cnt3 += 0x10000;
            if (((hm1 & 0x7fffffff) | __double2loint(em1)) == 0) cnt3 += 0x10000;
            if (NSH > 0) {
                // Row test first (warp-uniform): x*x beyond the row's limit means no column of this row lies inside any
                // shell (table filled above).  Rows that pass compare every point exactly: s is the evaluator's own
                // (x*x + y*y) + z*z, in float32 arithmetic on the float32 grid.
                if (xxi > lds_f64<RJ * 8>(ya)) return;
                const double yyj = lds_f64<0>(ya);
                double s0, s1;
                if (DT == IWB_F64) {
                    const double sxy = xxi + yyj;
                    s0 = sxy + zz0; s1 = sxy + zz1;
                } else {
                    const float sxy = __fadd_rn((float)xxi, (float)yyj);
                    s0 = (double)__fadd_rn(sxy, (float)zz0); s1 = (double)__fadd_rn(sxy, (float)zz1);
                }
#pragma unroll
                for (int sh = 0; sh < NSH; ++sh) {
                    // the reference's own mask multiply (float32 thresholds are float32 values: the compare is exact in double)
                    const bool in0 = (s0 <= P.shell_thr[sh]), in1 = (s1 <= P.shell_thr[sh]);
                    const double m0 = in0 ? 1.0 : 0.0, m1 = in1 ? 1.0 : 0.0;
                    sh_sum[sh] = fma(em0, m0, sh_sum[sh]);
                    sh_abs[sh] = fma(fabs(em0), m0, sh_abs[sh]);
                    sh_sum[sh] = fma(em1, m1, sh_sum[sh]);
                    sh_abs[sh] = fma(fabs(em1), m1, sh_abs[sh]);
                    if (in0 && om0) ++sh_cnt[sh];
                    if (in1 && om1) ++sh_cnt[sh];
                }
            }
        };

        // ---- one-sided value at a global face: ((a f0) + (b f1)) + (c f2) over three consecutive entries --------
        // `addr` is the entry AT the face, STEP the byte distance to its neighbour along the axis; `hi` selects the
        // far face (entries idx-2, idx-1, idx with a1 b1 c1) instead of the near one (idx, idx+1, idx+2 with a0 b0 c0).
        auto edge_at = [&](unsigned addr, auto step_tag, bool hi, const double* ec /* a0 b0 c0 a1 b1 c1 */) -> double {
            constexpr int STEP = decltype(step_tag)::value;
            const unsigned b = hi ? addr - 2u * (unsigned)STEP : addr;
            const double* c = hi ? ec + 3 : ec;
            return edge3(c[0], lds_f64<0>(b), c[1], lds_f64<STEP>(b), c[2], lds_f64<2 * STEP>(b));
        };
        using step_k = std::integral_constant<int, 8>;           // next column
        using step_j = std::integral_constant<int, RK * 8>;      // next row

        // ================= the three tasks of a step ===========================================================
        // Explicit shared-memory addresses: per row one base register per plane and immediates for the
        // neighbours; all loads of a row are issued before its arithmetic, stores last.  Every lane evaluates
        // the interior (central) formula; rows / columns / planes ON a global face then override the affected
        // derivatives with the one-sided formulas (numpy.gradient, edge_order=2) -- blocks that interior tiles
        // skip with one uniform branch.  O(i) and C(q) share the row loop (they read and write disjoint
        // planes), so the short dependent chains of C fill the gaps of O.
        auto stepOC = [&](bool doO, int i, double xxi, bool doC, int q, int s3q) {
            // CTA-uniform plane offsets of this step (kept: computed once per step, not once per row)
            const unsigned o_pxi = keep_u32(PXB + (unsigned)(i & 3) * PLB);
            const unsigned o_pxp = keep_u32(PXB + (unsigned)((i + 1) & 3) * PLB);      // phi_x(i+1)
            const unsigned o_pxm = keep_u32(PXB + (unsigned)((i + 3) & 3) * PLB);      // phi_x(i-1)
            const unsigned o_py = keep_u32(PYB + (unsigned)(i & 1) * PLB);             // phi_y(i); phi_z(i) at + 2 PLB
            const unsigned o_ph = keep_u32((unsigned)s3q * PLB);                       // phi(q)
            const unsigned o_w = keep_u32(PYB + (unsigned)(q & 1) * PLB);              // phi_y(q); phi_z(q) at + 2 PLB
            constexpr int R8 = RK * 8, ZB = 2 * (int)PLB;
            const int iedge = (i == 0) ? 1 : ((i == P.n0 - 1) ? 2 : 0);
            const bool o_face = tile_edge || (iedge != 0);
            unsigned A = a_row0, ya = yy_row0;
            if (PROF && doO) prof_clear();
#if IWB_OC_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
            for (int jr = ty, it = 0; jr < RJ; jr += NW, A += ROWB, ya += NW * 8, ++it) {
                const int jedge = (t_jlo && jr == 2) ? 1 : ((t_jhi && jr == Tj + 1) ? 2 : 0);
                if (doO && jr >= 2 && jr <= Tj + 1) {
                    const unsigned Ai = A + o_pxi, Ap = A + o_pxp, Am = A + o_pxm, Ay = A + o_py;
                    // column a = this lane, column b = lane + 32 (256 bytes further)
                    const double xpa = lds_f64<0>(Ap), xma = lds_f64<0>(Am);
                    const double xpb = lds_f64<256>(Ap), xmb = lds_f64<256>(Am);
                    const double xjpa = lds_f64<R8>(Ai), xjma = lds_f64<-R8>(Ai);
                    const double xjpb = lds_f64<256 + R8>(Ai), xjmb = lds_f64<256 - R8>(Ai);
                    const double xkpa = lds_f64<8>(Ai), xkma = lds_f64<-8>(Ai);
                    const double xkpb = lds_f64<256 + 8>(Ai), xkmb = lds_f64<256 - 8>(Ai);
                    const double yjpa = lds_f64<R8>(Ay), yjma = lds_f64<-R8>(Ay);
                    const double yjpb = lds_f64<256 + R8>(Ay), yjmb = lds_f64<256 - R8>(Ay);
                    const double ykpa = lds_f64<8>(Ay), ykma = lds_f64<-8>(Ay);
                    const double ykpb = lds_f64<256 + 8>(Ay), ykmb = lds_f64<256 - 8>(Ay);
                    const double zkpa = lds_f64<ZB + 8>(Ay), zkma = lds_f64<ZB - 8>(Ay);
                    const double zkpb = lds_f64<ZB + 256 + 8>(Ay), zkmb = lds_f64<ZB + 256 - 8>(Ay);
                    double pxx[2] = {cdivm<MODE>(xpa - xma, cx.h2, cx.rh2), cdivm<MODE>(xpb - xmb, cx.h2, cx.rh2)};
                    double pxy[2] = {cdivm<MODE>(xjpa - xjma, cy.h2, cy.rh2), cdivm<MODE>(xjpb - xjmb, cy.h2, cy.rh2)};
                    double pxz[2] = {cdivm<MODE>(xkpa - xkma, cz.h2, cz.rh2), cdivm<MODE>(xkpb - xkmb, cz.h2, cz.rh2)};
                    double pyy[2] = {cdivm<MODE>(yjpa - yjma, cy.h2, cy.rh2), cdivm<MODE>(yjpb - yjmb, cy.h2, cy.rh2)};
                    double pyz[2] = {cdivm<MODE>(ykpa - ykma, cz.h2, cz.rh2), cdivm<MODE>(ykpb - ykmb, cz.h2, cz.rh2)};
                    double pzz[2] = {cdivm<MODE>(zkpa - zkma, cz.h2, cz.rh2), cdivm<MODE>(zkpb - zkmb, cz.h2, cz.rh2)};
                    if (o_face) {
                        if (iedge == 1) {            // phi_x(0), (1), (2)
                            const unsigned A2 = A + PXB + (unsigned)((i + 2) & 3) * PLB;
                            pxx[0] = edge3(ec_s[0], lds_f64<0>(Ai), ec_s[1], lds_f64<0>(Ap), ec_s[2], lds_f64<0>(A2));
                            pxx[1] = edge3(ec_s[0], lds_f64<256>(Ai), ec_s[1], lds_f64<256>(Ap), ec_s[2], lds_f64<256>(A2));
                        } else if (iedge == 2) {     // phi_x(n0-3), (n0-2), (n0-1)
                            const unsigned A2 = A + PXB + (unsigned)((i + 2) & 3) * PLB;
                            pxx[0] = edge3(ec_s[3], lds_f64<0>(A2), ec_s[4], lds_f64<0>(Am), ec_s[5], lds_f64<0>(Ai));
                            pxx[1] = edge3(ec_s[3], lds_f64<256>(A2), ec_s[4], lds_f64<256>(Am), ec_s[5], lds_f64<256>(Ai));
                        }
                        if (jedge) {                 // row on the face j == 0 / j == n1-1 (warp-uniform)
                            pxy[0] = edge_at(Ai, step_j{}, jedge == 2, ec_s + 6);
                            pxy[1] = edge_at(Ai + 256u, step_j{}, jedge == 2, ec_s + 6);
                            pyy[0] = edge_at(Ay, step_j{}, jedge == 2, ec_s + 6);
                            pyy[1] = edge_at(Ay + 256u, step_j{}, jedge == 2, ec_s + 6);
                        }
                        if (cf0 & 6u) {              // column on the face k == 0 / k == n2-1 (single lanes)
                            pxz[0] = edge_at(Ai, step_k{}, (cf0 & 4u) != 0, ec_s + 12);
                            pyz[0] = edge_at(Ay, step_k{}, (cf0 & 4u) != 0, ec_s + 12);
                            pzz[0] = edge_at(Ay + (unsigned)ZB, step_k{}, (cf0 & 4u) != 0, ec_s + 12);
                        }
                        if (cf1 & 6u) {
                            pxz[1] = edge_at(Ai + 256u, step_k{}, (cf1 & 4u) != 0, ec_s + 12);
                            pyz[1] = edge_at(Ay + 256u, step_k{}, (cf1 & 4u) != 0, ec_s + 12);
                            pzz[1] = edge_at(Ay + (unsigned)ZB + 256u, step_k{}, (cf1 & 4u) != 0, ec_s + 12);
                        }
                    }
                    double rho2[2];
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        // K_ij = -phi_ij: the sign cancels in K^2 and in K_ij K^ij (reproduce_rodal_exact.py:198-212)
                        const double tr = (pxx[a] + pyy[a]) + pzz[a];
                        const double diag = (pxx[a] * pxx[a] + pyy[a] * pyy[a]) + pzz[a] * pzz[a];
                        const double off = (pxy[a] * pxy[a] + pxz[a] * pxz[a]) + pyz[a] * pyz[a];
                        const double kk = fma(2.0, off, diag);     // 2.0*off is exact -> same rounding as diag + 2.0*off
                        rho2[a] = cdivm<MODE>(tr * tr - kk, P.c16pi, P.rc16pi);
                    }
                    if (EMIT) {
                        const long long j = j0 - 2 + jr, k = k0 - hl + tx;
                        double* dst = P.rho_out + ((long long)(i - P.i_begin) * P.n1 + j) * P.n2 + k;
                        if (cf0 & 16u) dst[0] = rho2[0];
                        if (cf1 & 16u) dst[32] = rho2[1];
                    }
                    const double e2[2] = {rho2[0] * P.dV, rho2[1] * P.dV};
                    tally2(e2, rho2, ya, xxi, it);
                }
                if (doC && jr >= py_lo && jr <= py_hi) {
                    const unsigned Ah = A + o_ph, Aw = A + o_w;
                    const bool z_row = (jr >= 2 && jr <= Tj + 1);
                    const double hjpa = lds_f64<R8>(Ah), hjma = lds_f64<-R8>(Ah);
                    const double hjpb = lds_f64<256 + R8>(Ah), hjmb = lds_f64<256 - R8>(Ah);
                    const double hkpa = lds_f64<8>(Ah), hkma = lds_f64<-8>(Ah);
                    const double hkpb = lds_f64<256 + 8>(Ah), hkmb = lds_f64<256 - 8>(Ah);
                    double pya = cdivm<MODE>(hjpa - hjma, cy.h2, cy.rh2);
                    double pyb = cdivm<MODE>(hjpb - hjmb, cy.h2, cy.rh2);
                    double pza = cdivm<MODE>(hkpa - hkma, cz.h2, cz.rh2);
                    double pzb = cdivm<MODE>(hkpb - hkmb, cz.h2, cz.rh2);
                    if (tile_edge) {
                        if (jedge) {
                            pya = edge_at(Ah, step_j{}, jedge == 2, ec_s + 6);
                            pyb = edge_at(Ah + 256u, step_j{}, jedge == 2, ec_s + 6);
                        }
                        if (cf0 & 6u) pza = edge_at(Ah, step_k{}, (cf0 & 4u) != 0, ec_s + 12);
                        if (cf1 & 6u) pzb = edge_at(Ah + 256u, step_k{}, (cf1 & 4u) != 0, ec_s + 12);
                    }
                    sts_f64<0>(Aw, pya);
                    sts_f64<256>(Aw, pyb);
                    if (z_row) {
                        sts_f64<ZB>(Aw, pza);
                        sts_f64<ZB + 256>(Aw, pzb);
                    }
                }
            }
            if (PROF && doO) prof_reduce();
        };

        // P(p): phi(p) at the own columns and phi_x(p-1) = (phi(p) - phi(p-2)) / 2dx, which needs phi(p-2) of the
        // own column only; with p == 2 at the start of the axis also the one-sided phi_x(0).
        auto taskP = [&](int p, int s3 /* p mod 3 */, double xp, double xxp) {
            const int s3m2 = (s3 == 2) ? 0 : s3 + 1;
            const unsigned o_p = keep_u32((unsigned)s3 * PLB);                         // phi(p)
            const unsigned o_m2 = keep_u32((unsigned)s3m2 * PLB);                      // phi(p-2)
            const unsigned o_xw = keep_u32(PXB + (unsigned)((p + 3) & 3) * PLB);       // phi_x(p-1)
            const bool do_px = (p - 2 >= pstart);
            const bool do_px0 = (p == 2) && (pstart == 0);                             // phi_x(0), one-sided
            unsigned A = a_row0, ya = yy_row0;
#if IWB_P4
            // Both rows of the warp at once when both exist: the sqrt / division chains of phi are ~20 dependent
            // FP64 operations long, and four independent evaluations per thread (instead of two) keep the FP64 pipe
            // fed with five warps per scheduler (measured: 92 -> 97 Gpoints/s).
            if (!do_px0 && ty >= jr_lo && ty + NW <= jr_hi) {
                const unsigned Am2 = A + o_m2;
                const double yyA = lds_f64<0>(ya), yyB = lds_f64<NW * 8>(ya);
                const double m2[4] = {lds_f64<0>(Am2), lds_f64<256>(Am2), lds_f64<ROWB>(Am2), lds_f64<ROWB + 256>(Am2)};
                const bool origin = (p == P.i_zero) && (j0 - 2 + ty == P.j_zero || j0 - 2 + ty + NW == P.j_zero);
                double f[4];
                if (DT == IWB_F64) {
                    const double sA = xxp + yyA, sB = xxp + yyB;
                    const double s[4] = {sA + zz0, sA + zz1, sB + zz0, sB + zz1};
                    if (MODE == 0) phi_exact_f64<4, true>(s, xp, pc, f, origin || !pc.prechecked);
                    else phi_fast_f64<4>(s, xp, pc, f, origin || !pc.prechecked);
                } else {
                    const float sA = __fadd_rn((float)xxp, (float)yyA), sB = __fadd_rn((float)xxp, (float)yyB);
                    const float s[4] = {__fadd_rn(sA, (float)zz0), __fadd_rn(sA, (float)zz1),
                                        __fadd_rn(sB, (float)zz0), __fadd_rn(sB, (float)zz1)};
                    phi_exact_f32ref<4, true>(s, (float)xp, pc, f, origin || !pc.prechecked);
                }
                const unsigned Ap = A + o_p;
                sts_f64<0>(Ap, f[0]); sts_f64<256>(Ap, f[1]); sts_f64<ROWB>(Ap, f[2]); sts_f64<ROWB + 256>(Ap, f[3]);
                if (do_px) {
                    const unsigned Ax = A + o_xw;
                    sts_f64<0>(Ax, cdivm<MODE>(f[0] - m2[0], cx.h2, cx.rh2));
                    sts_f64<256>(Ax, cdivm<MODE>(f[1] - m2[1], cx.h2, cx.rh2));
                    sts_f64<ROWB>(Ax, cdivm<MODE>(f[2] - m2[2], cx.h2, cx.rh2));
                    sts_f64<ROWB + 256>(Ax, cdivm<MODE>(f[3] - m2[3], cx.h2, cx.rh2));
                }
                return;
            }
#endif
#pragma unroll 1
            for (int jr = ty; jr <= jr_hi; jr += NW, A += ROWB, ya += NW * 8) {
                if (jr < jr_lo) continue;
                const unsigned Am2 = A + o_m2;
                const double yyj = lds_f64<0>(ya);
                const double m2a = lds_f64<0>(Am2), m2b = lds_f64<256>(Am2);   // harmless when !do_px (never used)
                double f[2];
                if (DT == IWB_F64) {
                    const double sxy = xxp + yyj;
                    const double s[2] = {sxy + zz0, sxy + zz1};
                    // the row through the origin (r == 0 at one lane) takes the reference flavour
                    const bool origin_row = (p == P.i_zero) && (j0 - 2 + jr == P.j_zero);
                    // (exotic grids / parameters fail the host's operand-range proof: reference flavour throughout)
                    if (MODE == 0) phi_exact_f64<2, true>(s, xp, pc, f, origin_row || !pc.prechecked);
                    else phi_fast_f64<2>(s, xp, pc, f, origin_row || !pc.prechecked);
                } else {
                    const float sxy = __fadd_rn((float)xxp, (float)yyj);
                    const float s[2] = {__fadd_rn(sxy, (float)zz0), __fadd_rn(sxy, (float)zz1)};
                    const bool origin_row = (p == P.i_zero) && (j0 - 2 + jr == P.j_zero);
                    phi_exact_f32ref<2, true>(s, (float)xp, pc, f, origin_row || !pc.prechecked);
                }
                const unsigned Ap = A + o_p;
                sts_f64<0>(Ap, f[0]);
                sts_f64<256>(Ap, f[1]);
                if (do_px) {
                    const unsigned Ax = A + o_xw;
                    sts_f64<0>(Ax, cdivm<MODE>(f[0] - m2a, cx.h2, cx.rh2));
                    sts_f64<256>(Ax, cdivm<MODE>(f[1] - m2b, cx.h2, cx.rh2));
                }
                if (do_px0) {      // p == 2: phi(0) is plane p-2 (slot 0), phi(1) in slot 1; phi_x(0) goes to slot 0 of its ring
                    const unsigned A1 = A + PLB, Ax0 = A + PXB;
                    sts_f64<0>(Ax0, edge3(ec_s[0], m2a, ec_s[1], lds_f64<0>(A1), ec_s[2], f[0]));
                    sts_f64<256>(Ax0, edge3(ec_s[0], m2b, ec_s[1], lds_f64<256>(A1), ec_s[2], f[1]));
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
                px_w[base] = edge3(ec_s[3], ph_m2[base], ec_s[4], ph_m1[base], ec_s[5], ph_p[base]);
                px_w[base + 32] = edge3(ec_s[3], ph_m2[base + 32], ec_s[4], ph_m1[base + 32], ec_s[5], ph_p[base + 32]);
            }
        };

        // ---- block-level flush of the register sums into partial[fb][tile] ----------------
        auto flush = [&](int fb, int i_last, int i_end_march) {
            constexpr int NV = 6 + 3 * NSH;    // values actually reduced
            if (PROF) {
                // every warp has finished the tallies of this flush block (they arrived at the step barrier): warp 0 adds
                // the windows in warp order, stores the unit's window and sets up the window of the next block
                if (beyond_seen && b0_reg + PROF_WIN < P.prof_nb) atomicOr(P.prof_overflow, 1u);
                beyond_seen = false;
                if (ty == 0) {
                    if (tx < PROF_WIN) {
                        double a = 0.0;
                        int c = 0;
                        for (int w = 0; w < NW; ++w) {
                            a += scratch[w * ROW + tx];
                            c += reinterpret_cast<const int*>(scratch + w * ROW + PROF_WIN)[tx];
                        }
                        const size_t o = ((size_t)fb * ntiles + rnk) * PROF_WIN + tx;
                        P.prof_sum[o] = a;
                        P.prof_cnt[o] = c;
                    }
                    if (tx == 0) P.prof_base[(size_t)fb * ntiles + rnk] = b0_reg;
                    __syncwarp();
                    if (i_last + 1 < i_end_march) prof_rebase(i_last + 1, min(i_last + 1 + FLUSH, i_end_march));
                }
                __syncthreads();       // the windows have been read: the scratch may take the flush values
            }
            double vals[NV];
            // E+ = (sum + sum|.|)/2 and E- = (sum - sum|.|)/2, per thread (halving is exact)
            vals[0] = (acc_sum + acc_abs) * 0.5; vals[1] = (acc_sum - acc_abs) * 0.5;
            vals[2] = __longlong_as_double((long long)((cnt3 >> 8) & 0xff));
            vals[3] = __longlong_as_double((long long)(cnt3 & 0xff));
            vals[4] = __longlong_as_double((long long)((cnt3 >> 16) & 0xff));   // zeros + masked columns (corrected below)
            vals[5] = __longlong_as_double(0LL);      // cnt_nan: derived by the host from the other three
#pragma unroll
            for (int s = 0; s < NSH; ++s) {
                vals[6 + s] = (sh_sum[s] + sh_abs[s]) * 0.5;
                vals[6 + NSH + s] = (sh_sum[s] - sh_abs[s]) * 0.5;
                vals[6 + 2 * NSH + s] = __longlong_as_double((long long)sh_cnt[s]);
            }
            // warp butterfly (fixed order).  The counts are at most 64 per thread and 2048 per warp: two of them share a
            // 32-bit word (16-bit fields), so 3 + NSH counts cost (3 + NSH + 1) / 2 one-register shuffles per step
            // instead of 3 + NSH two-register ones.
            constexpr int NCNT = 3 + NSH, NCW = (NCNT + 1) / 2;
            unsigned cw[NCW];
            {
                int c[2 * NCW];
                c[0] = cnt3 & 0xff; c[1] = (cnt3 >> 8) & 0xff; c[2] = (cnt3 >> 16) & 0xff;     // neg, pos, zero
#pragma unroll
                for (int s2 = 0; s2 < NSH; ++s2) c[3 + s2] = sh_cnt[s2];
                if (NCNT & 1) c[NCNT] = 0;
#pragma unroll
                for (int k = 0; k < NCW; ++k) cw[k] = (unsigned)c[2 * k] | ((unsigned)c[2 * k + 1] << 16);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int k = 0; k < NCW; ++k) cw[k] += __shfl_xor_sync(0xffffffffu, cw[k], off);
            }
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const bool is_int = (q >= 2 && q < 6) || (q >= 6 + 2 * NSH);
                if (is_int) continue;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) vals[q] += __shfl_xor_sync(0xffffffffu, vals[q], off);
            }
            {
                auto field = [&](int idx) { return (long long)((cw[idx >> 1] >> ((idx & 1) * 16)) & 0xffffu); };
                vals[3] = __longlong_as_double(field(0));
                vals[2] = __longlong_as_double(field(1));
                vals[4] = __longlong_as_double(field(2));
#pragma unroll
                for (int s2 = 0; s2 < NSH; ++s2) vals[6 + 2 * NSH + s2] = __longlong_as_double(field(3 + s2));
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
                if (q >= 0) {
                    const bool is_int = (q >= 2 && q < 6) || (q >= 6 + 2 * NSH);
                    acc = scratch[q];
                    for (int wv = 1; wv < NW; ++wv) {
                        double o = scratch[wv * NV + q];
                        if (is_int) acc = __longlong_as_double(__double_as_longlong(acc) + __double_as_longlong(o));
                        else acc += o;
                    }
                    // the zero field counted the masked (non-output) columns of every row as well
                    if (q == 4)
                        acc = __longlong_as_double(__double_as_longlong(acc) -
                                                   (long long)Tj * (RK - Tk) * (i_last - flush_from + 1));
                }
                P.partial[((size_t)fb * ntiles + rnk) * ROW + slot] = acc;     // slot of the tile's schedule index
            }
            // the scratch is rewritten at the next flush only, at least one step barrier from here
            acc_sum = 0.0; acc_abs = 0.0; cnt3 = 0;
            flush_from = i_last + 1;
#pragma unroll
            for (int s = 0; s < NSHA; ++s) { sh_sum[s] = 0.0; sh_abs[s] = 0.0; sh_cnt[s] = 0; }
            if (PROF) {
                __syncthreads();       // the final reduction has read the scratch: it holds the windows again
                if (tx < PROF_WIN + (PROF_WIN + 1) / 2) win_sum[tx] = 0.0;      // sums and (as zero bits) counts
                __syncwarp();
                b0_reg = s_b0;
                pb = 0;
            }
        };

        auto runP = [&](int p, int s3, double xp, double xxp) { taskP(p, s3, xp, xxp); };
        auto runC = [&](int q, int s3) { stepOC(false, 0, 0.0, true, q, s3); };
        // O(i) followed by C(i+1) (the latter when doC)
        auto runOC = [&](int i, double xxi, bool doC, int m3p1) { stepOC(true, i, xxi, doC, i + 1, m3p1); };
        // phi_x(n0-1) may be written once every phi plane exists and its ring slot (that of phi_x(n0-5),
        // last read by O(n0-4)) is dead, i.e. from the step of output plane n0-3 on; it is first needed
        // by O(n0-2) (n0 == 3: by every plane, and 0 >= n0-3 holds).
        auto px_last_ok = [&](int i) { return i >= P.n0 - 3; };

        // ---- prologue: everything O(i0) needs ----------------------------------------------------------
        __syncthreads();                 // yy_s visible
        if (PROF) {                      // window of the first flush block of this march (published by the barriers below)
            if (ty == 0) prof_rebase(i0, min(i1, (i0 / FLUSH + 1) * FLUSH));
            if (tx < PROF_WIN + (PROF_WIN + 1) / 2) win_sum[tx] = 0.0;
        }
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

        if (PROF) { b0_reg = s_b0; pb = 0; beyond_seen = false; }

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
            // steady state: produced + 1 == i + 3, whose slot is i mod 3; otherwise (start / end of the axis) compute it
            const int pp = produced + 1;
            runOC(i, xxi, doC, m3p1);
            step_arrive();
            if (doP) runP(pp, (pp == i + 3) ? m3 : pp % 3, xp_n, xxp_n);
            if (doL) taskP_last();
            step_wait();
            if (doP) ++produced;
            if (doL) px_last_done = true;
            if (((i + 1) & (FLUSH - 1)) == 0 || i == i1 - 1) flush(i / FLUSH, i, i1);
            m3 = m3p1;
        }
    }   // units of the span
    }   // spans
}

template <int DT, int NSH, int RJ, int NW, bool EMIT, int MINB = 1, int MODE = 0>
__global__ void __launch_bounds__(NW * 32, MINB)
k_energy3d(const __grid_constant__ K3Params P)
{
    energy3d_body<DT, NSH, RJ, NW, EMIT, MODE, false>(P);
}

// The same march with the radial profile of rho fused in (binning rule of radial_profile.cuh):
// launched with Geo::SMEM_BYTES + PROF_THR_S doubles of dynamic shared memory.
template <int DT, int NSH, int RJ, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
k_energy3d_prof(const __grid_constant__ K3Params P)
{
    energy3d_body<DT, NSH, RJ, NW, false, 0, true>(P);
}

// ---- fixed-order sum over tiles: partial[fb][tile][ROW] -> table[fb][ROW] ----------------------
// (static: every translation unit that includes this header gets its own copy)
// The tiles of a flush block (slots in schedule order) are added as NCH chains -- chain c adds the slots c, c + NCH,
// c + 2 NCH, ... in increasing order -- which a fixed binary tree then combines.  Deterministic and independent of
// the decomposition: a slab sees all tiles of its flush blocks, an interleaved tile shard (stride S | NCH) owns whole
// chains.  k_reduce_tiles does both steps in one block per flush block (blockDim.x = 32 * NCH, one slot per lane);
// k_chain_sums / k_combine_chains are the two halves around the all-gather of the tile-shard path and perform the
// same additions in the same order.  With 16 chains the 522 tiles of n = 1024 are 33 dependent additions per thread.
static_assert(ROW <= 32, "one slot per lane");
__device__ __forceinline__ bool row_slot_is_int(int q) { return (q >= 2 && q < 6) || (q >= 6 + 2 * IWB_MAX_SHELLS); }

// sum of the slots c, c + nch, ... of `src` (= partial + fb * ntiles * ROW + q), as a double or as int64 bits
__device__ __forceinline__ double chain_sum(const double* __restrict__ src, int ntiles, int c, int nch, bool is_int)
{
    double a = 0.0;
    long long ia = 0;
    int t = c;
    for (; t + 7 * nch < ntiles; t += 8 * nch) {      // eight loads in flight, added in tile order
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = src[(size_t)(t + u * nch) * ROW];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (is_int) ia += __double_as_longlong(v[u]); else a += v[u];
        }
    }
    for (; t < ntiles; t += nch) {
        const double v = src[(size_t)t * ROW];
        if (is_int) ia += __double_as_longlong(v); else a += v;
    }
    return is_int ? __longlong_as_double(ia) : a;
}

// binary tree over the chains held in s_part[c][q]; result in s_part[0][q]
__device__ __forceinline__ void chain_tree(double (*s_part)[33], int q, int c, int nch, bool is_int)
{
    __syncthreads();
    for (int h = nch >> 1; h > 0; h >>= 1) {
        if (c < h) {
            const double lo = s_part[c][q], hi = s_part[c + h][q];
            s_part[c][q] = is_int ? __longlong_as_double(__double_as_longlong(lo) + __double_as_longlong(hi)) : lo + hi;
        }
        __syncthreads();
    }
}

static __global__ void __launch_bounds__(1024) k_reduce_tiles(const double* __restrict__ partial,
                                                              double* __restrict__ table, int fb_begin, int ntiles,
                                                              unsigned int* counter = nullptr)
{
    __shared__ double s_part[32][33];
    if (counter && blockIdx.x == 0 && threadIdx.x == 0) *counter = 0u;    // re-arm the work queue of k_energy3d
    const int fb = fb_begin + blockIdx.x;
    const int q = threadIdx.x & 31, c = threadIdx.x >> 5, nch = blockDim.x >> 5;
    const bool is_int = row_slot_is_int(q);
    s_part[c][q] = (q < ROW) ? chain_sum(partial + (size_t)fb * ntiles * ROW + q, ntiles, c, nch, is_int) : 0.0;
    chain_tree(s_part, q, c, nch, is_int);
    if (c == 0 && q < ROW) table[(size_t)fb * ROW + q] = s_part[0][q];
}

// tile-shard path, first half: the chains c = first + lc * stride (lc = 0 .. NCH/stride - 1) of every flush block
// -> chains[fb][lc][ROW].  One block per flush block, blockDim.x = 32 * NCH / stride.
static __global__ void __launch_bounds__(1024) k_chain_sums(const double* __restrict__ partial, double* __restrict__ chains,
                                                            int fb_begin, int ntiles, int first, int stride, int nch,
                                                            unsigned int* counter = nullptr)
{
    if (counter && blockIdx.x == 0 && threadIdx.x == 0) *counter = 0u;
    const int fb = fb_begin + blockIdx.x;
    const int q = threadIdx.x & 31, lc = threadIdx.x >> 5, nlc = blockDim.x >> 5;
    if (q >= ROW) return;
    const double v = chain_sum(partial + (size_t)fb * ntiles * ROW + q, ntiles, first + lc * stride, nch, row_slot_is_int(q));
    chains[((size_t)fb * nlc + lc) * ROW + q] = v;
}

// tile-shard path, second half: gathered[shard][fb][lc][ROW] -> table[fb][ROW]; chain c is entry lc = c / stride of
// shard c % stride.  One block per flush block, blockDim.x = 32 * NCH.
static __global__ void __launch_bounds__(1024) k_combine_chains(const double* __restrict__ gathered, double* __restrict__ table,
                                                                int nfb, int stride)
{
    __shared__ double s_part[32][33];
    const int fb = blockIdx.x;
    const int q = threadIdx.x & 31, c = threadIdx.x >> 5, nch = blockDim.x >> 5;
    const int nlc = nch / stride;
    const bool is_int = row_slot_is_int(q);
    s_part[c][q] = (q < ROW) ? gathered[(((size_t)(c % stride) * nfb + fb) * nlc + c / stride) * ROW + q] : 0.0;
    chain_tree(s_part, q, c, nch, is_int);
    if (c == 0 && q < ROW) table[(size_t)fb * ROW + q] = s_part[0][q];
}

// ---- tile-shard path with the exchange fused in: NVLink peer stores instead of an NCCL all-gather ------------------------
// k_chain_sums_push is k_chain_sums writing every chain sum straight into the gathered buffer of EVERY rank (its own
// included) through peer-mapped pointers; k_exchange_combine then signals the peers (one release-store of the epoch per
// peer, ordered after the push kernel by the stream), waits for the epoch flag of every peer, and runs the chain tree and
// the row sum of the single-GPU path (the same additions in the same order: bit-identical energies on every rank).
// Buffers and flags are double-buffered by epoch parity, so a fast rank may already push evaluation e + 1 while a slow one
// still combines evaluation e; it cannot reach e + 2 before the slow rank has signalled e + 1, i.e. finished reading e.
constexpr int PEER_MAX = IWB_TILE_CHAINS;
struct PeerPtrs {
    double* gath[PEER_MAX];                  // this epoch's gathered buffer [S][nfb][NCH / S][ROW] of every rank (peer-mapped)
    unsigned long long* flag[PEER_MAX];      // this epoch's flag row [S] of every rank
};

static __global__ void __launch_bounds__(1024) k_chain_sums_push(const double* __restrict__ partial, const PeerPtrs peers,
                                                                 int fb_begin, int nfb, int ntiles, int first, int stride,
                                                                 int nch, unsigned int* counter)
{
    if (counter && blockIdx.x == 0 && threadIdx.x == 0) *counter = 0u;
    const int fb = fb_begin + blockIdx.x;
    const int q = threadIdx.x & 31, lc = threadIdx.x >> 5, nlc = blockDim.x >> 5;
    if (q >= ROW) return;
    const double v = chain_sum(partial + (size_t)fb * ntiles * ROW + q, ntiles, first + lc * stride, nch, row_slot_is_int(q));
    const size_t o = (((size_t)first * nfb + fb) * nlc + lc) * ROW + q;
    for (int r = 0; r < stride; ++r) peers.gath[r][o] = v;
}

struct ExchParams {
    PeerPtrs peers;
    const double* gath_local;                // this rank's gathered buffer of the epoch
    const unsigned long long* flag_local;    // this rank's flag row of the epoch
    unsigned long long epoch;
    int rank, stride, nfb;
    double* table;                           // [nfb][ROW] scratch
    double* row;                             // [ROW] result
    unsigned int* ticket;                    // 0 at launch; the last block resets it
    unsigned long long timeout_ns;           // longest wait for the peers' flags before the kernel traps
};

static __global__ void __launch_bounds__(1024) k_exchange_combine(const ExchParams X)
{
    __shared__ double s_part[32][33];
    __shared__ unsigned s_last;
    const int fb = blockIdx.x;
    const int q = threadIdx.x & 31, c = threadIdx.x >> 5, nch = blockDim.x >> 5;
    // every chain sum of this rank is in place on every peer (the push kernel precedes this one on the stream): tell them
    if (blockIdx.x == 0 && threadIdx.x < X.stride)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(X.peers.flag[threadIdx.x] + X.rank), "l"(X.epoch) : "memory");
    if (threadIdx.x < X.stride) {
        // A peer that never signals (its process died, or it never made the call) must not leave this GPU spinning for
        // ever: after the time-out (60 s; IWB_EXCHANGE_TIMEOUT_S in the environment of the process) the kernel traps, the
        // stream reports an error and the host call fails loudly.
        unsigned long long seen, t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(X.flag_local + threadIdx.x) : "memory");
            if (seen >= X.epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > X.timeout_ns) __trap();
        }
    }
    __syncthreads();
    const int nlc = nch / X.stride;
    const bool is_int = row_slot_is_int(q);
    s_part[c][q] = (q < ROW) ? __ldcg(X.gath_local + (((size_t)(c % X.stride) * X.nfb + fb) * nlc + c / X.stride) * ROW + q) : 0.0;
    chain_tree(s_part, q, c, nch, is_int);
    if (c == 0 && q < ROW) X.table[(size_t)fb * ROW + q] = s_part[0][q];
    // the last block to finish adds the rows in index order (k_reduce_rows)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(X.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < ROW) {
        const int qq = threadIdx.x;
        const bool ii = row_slot_is_int(qq);
        double a = 0.0;
        long long ia = 0;
        for (int r = 0; r < X.nfb; ++r) {
            const double v = __ldcg(X.table + (size_t)r * ROW + qq);
            if (ii) ia += __double_as_longlong(v); else a += v;
        }
        X.row[qq] = ii ? __longlong_as_double(ia) : a;
    }
    if (threadIdx.x == 0) *X.ticket = 0u;
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

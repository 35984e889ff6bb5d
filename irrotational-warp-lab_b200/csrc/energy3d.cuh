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
// lanes along k.  Thread (tx, ty) owns region columns (jr = ty + NW*b, kr = tx + 32*a).
//
// Shared memory (doubles, RJ x 64 each):   phi ring [3]   planes p-2, p-1, p
//                                           phi_x ring [3] planes of d(phi)/dx
//                                           phi_y, phi_z   of the current output plane
// Per output plane i:   produce phi(i+2) and phi_x(i+1) (own columns only),
//                       phi_y(i), phi_z(i) from the phi(i) plane   | __syncthreads
//                       second derivatives, rho, predicated sums    | __syncthreads
// Sums are kept in registers and flushed (warp shuffle + fixed-order block tree) once
// per IWB_FLUSH_PLANES planes into partial[flush block][tile][row]; a second kernel adds
// the tiles in fixed order.  Flush blocks are aligned to the GLOBAL plane index, so the
// result is bit-identical for any slab decomposition along i.
#pragma once
#include "iwb_device.cuh"
#include "../../include/iwb200.h"

namespace iwb {

constexpr int RK = 64;              // region columns (k): two per lane
constexpr int TK_MAX = RK - 4;      // tile columns
constexpr int ROW = IWB_ROW_WIDTH;  // doubles per partial row
constexpr int FLUSH = IWB_FLUSH_PLANES;

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
};

// region geometry helpers
__host__ __device__ inline int tile_begin(int t, int nt, int n) { return (int)(((long long)t * n) / nt); }

template <int RJ, int NW>
struct Geo {
    static constexpr int NB = RJ / NW;          // rows per thread
    static constexpr int NC = NB * 2;           // columns per thread
    static constexpr int PLANE = RJ * RK;
    static constexpr int TJ_MAX = RJ - 4;
    static constexpr int NT = NW * 32;
    static constexpr size_t SMEM_BYTES = size_t(8) * PLANE * sizeof(double);
    static_assert(RJ % NW == 0, "rows must divide among warps");
    static_assert(NB % 2 == 0, "phi is evaluated in batches of 2 rows x 2 columns");
};

template <int DT, int NSH, int RJ, int NW, bool EMIT>
__global__ void __launch_bounds__(NW * 32, (Geo<RJ, NW>::SMEM_BYTES <= 112 * 1024) ? 2 : 1)
k_energy3d(const __grid_constant__ K3Params P)
{
    using G = Geo<RJ, NW>;
    constexpr int NB = G::NB;
    constexpr int PLANE = G::PLANE;
    extern __shared__ double smem[];
    double* const phi_ring = smem;
    double* const px_ring = smem + 3 * PLANE;
    double* const py_buf = smem + 6 * PLANE;
    double* const pz_buf = smem + 7 * PLANE;

    const int tx = threadIdx.x & 31;
    const int ty = threadIdx.x >> 5;
    const int ntiles = P.ntj * P.ntk;
    const AxisCoef cx = P.cx, cy = P.cy, cz = P.cz;

    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const int seg = item / ntiles;
        const int tile = item - seg * ntiles;
        const int tj = tile / P.ntk, tk = tile - tj * P.ntk;
        const int j0 = tile_begin(tj, P.ntj, P.n1), Tj = tile_begin(tj + 1, P.ntj, P.n1) - j0;
        const int k0 = tile_begin(tk, P.ntk, P.n2), Tk = tile_begin(tk + 1, P.ntk, P.n2) - k0;
        const int i0 = P.i_begin + seg * P.seg_len;
        const int i1 = min(P.i_end, i0 + P.seg_len);

        // ---- loop-invariant description of this thread's columns -------------------
        double yyv[NB], zzv[2];
        int jr_[NB];
        unsigned rowf[NB];   // bit0 valid, bit1 j==0, bit2 j==n1-1, bit3 wants phi_y/phi_z, bit4 output row
        unsigned colf[2];    // same for k
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            int jr = ty + NW * b, j = j0 - 2 + jr;
            bool valid = (j >= 0) && (j < P.n1) && (jr < Tj + 4);
            jr_[b] = jr;
            yyv[b] = P.yy[min(max(j, 0), P.n1 - 1)];
            rowf[b] = (valid ? 1u : 0u) | ((j == 0) ? 2u : 0u) | ((j == P.n1 - 1) ? 4u : 0u) |
                      ((valid && jr >= 1 && jr <= Tj + 2) ? 8u : 0u) |
                      ((jr >= 2 && jr <= Tj + 1) ? 16u : 0u);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            int kr = tx + 32 * a, k = k0 - 2 + kr;
            bool valid = (k >= 0) && (k < P.n2) && (kr < Tk + 4);
            zzv[a] = P.zz[min(max(k, 0), P.n2 - 1)];
            colf[a] = (valid ? 1u : 0u) | ((k == 0) ? 2u : 0u) | ((k == P.n2 - 1) ? 4u : 0u) |
                      ((valid && kr >= 1 && kr <= Tk + 2) ? 8u : 0u) |
                      ((kr >= 2 && kr <= Tk + 1) ? 16u : 0u);
        }

        // ---- accumulators (registers) -------------------------------------------------
        double acc_pos = 0.0, acc_neg = 0.0;
        double sh_pos[NSH > 0 ? NSH : 1], sh_neg[NSH > 0 ? NSH : 1];
        int c_pos = 0, c_neg = 0, c_zero = 0, c_nan = 0;
        int sh_cnt[NSH > 0 ? NSH : 1];
#pragma unroll
        for (int s = 0; s < NSH; ++s) { sh_pos[s] = 0.0; sh_neg[s] = 0.0; sh_cnt[s] = 0; }

        // ---- produce plane p: phi(p) and the x-derivatives it completes -----------------
        const int w0 = min(max(i0 - 1, 0), P.n0 - 3);
        const int pstart = max(0, w0 - 1);
        auto produce = [&](int p) {
            const double xp = P.x[p], xxp = P.xx[p];
            double* const ph_p = phi_ring + (p % 3) * PLANE;
            const double* const ph_m1 = phi_ring + ((p + 2) % 3) * PLANE;   // plane p-1
            const double* const ph_m2 = phi_ring + ((p + 1) % 3) * PLANE;   // plane p-2
            const bool do_px = (p - 2 >= pstart);            // phi_x(p-1), interior formula
            const bool do_px0 = (p == 2) && (pstart == 0);   // phi_x(0), one-sided
#pragma unroll
            for (int bb = 0; bb < NB; bb += 2) {
                double f[4];
                if (DT == IWB_F64) {
                    double s[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) s[q] = (xxp + yyv[bb + (q >> 1)]) + zzv[q & 1];
                    phi_exact_f64<4>(s, xp, P.pc, f);
                } else {
                    float s[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        s[q] = __fadd_rn(__fadd_rn((float)xxp, (float)yyv[bb + (q >> 1)]), (float)zzv[q & 1]);
                    phi_exact_f32ref<4>(s, (float)xp, P.pc, f);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int idx = jr_[bb + (q >> 1)] * RK + tx + 32 * (q & 1);
                    ph_p[idx] = f[q];
                    if (do_px) px_ring[((p + 2) % 3) * PLANE + idx] = cdiv(f[q] - ph_m2[idx], cx.h2, cx.rh2);
                    if (do_px0) px_ring[idx] = edge3(cx.a0, ph_m2[idx], cx.b0, ph_m1[idx], cx.c0, f[q]);
                }
            }
        };

        // ---- phi_x(n0-1), one-sided.  Its ring slot still holds phi_x(n0-4) until plane n0-3 has been
        //      output, so it is computed separately, once the march reaches plane n0-2 (or at once if n0 == 3).
        auto produce_px_last = [&]() {
            const int p = P.n0 - 1;
            const double* const ph_p = phi_ring + (p % 3) * PLANE;
            const double* const ph_m1 = phi_ring + ((p + 2) % 3) * PLANE;
            const double* const ph_m2 = phi_ring + ((p + 1) % 3) * PLANE;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const int idx = jr_[b] * RK + tx + 32 * a;
                    px_ring[(p % 3) * PLANE + idx] = edge3(cx.a1, ph_m2[idx], cx.b1, ph_m1[idx], cx.c1, ph_p[idx]);
                }
            }
        };

        // ---- phi_y(i), phi_z(i) of the own columns from the phi(i) plane ------------------
        auto inplane = [&](int i) {
            const double* const ph = phi_ring + (i % 3) * PLANE;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int jr = jr_[b];
                const bool row_ok = (rowf[b] & 8u) != 0;
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const int kr = tx + 32 * a;
                    const int idx = jr * RK + kr;
                    if (row_ok && (colf[a] & 1u)) {
                        double d;
                        if (rowf[b] & 2u) d = edge3(cy.a0, ph[idx], cy.b0, ph[idx + RK], cy.c0, ph[idx + 2 * RK]);
                        else if (rowf[b] & 4u) d = edge3(cy.a1, ph[idx - 2 * RK], cy.b1, ph[idx - RK], cy.c1, ph[idx]);
                        else d = cdiv(ph[idx + RK] - ph[idx - RK], cy.h2, cy.rh2);
                        py_buf[idx] = d;
                    }
                    if ((rowf[b] & 1u) && (colf[a] & 8u)) {
                        double d;
                        if (colf[a] & 2u) d = edge3(cz.a0, ph[idx], cz.b0, ph[idx + 1], cz.c0, ph[idx + 2]);
                        else if (colf[a] & 4u) d = edge3(cz.a1, ph[idx - 2], cz.b1, ph[idx - 1], cz.c1, ph[idx]);
                        else d = cdiv(ph[idx + 1] - ph[idx - 1], cz.h2, cz.rh2);
                        pz_buf[idx] = d;
                    }
                }
            }
        };

        // d/dy and d/dz of a staged plane at an output column
        auto dY = [&](const double* B, int idx, unsigned rf) -> double {
            if (rf & 2u) return edge3(cy.a0, B[idx], cy.b0, B[idx + RK], cy.c0, B[idx + 2 * RK]);
            if (rf & 4u) return edge3(cy.a1, B[idx - 2 * RK], cy.b1, B[idx - RK], cy.c1, B[idx]);
            return cdiv(B[idx + RK] - B[idx - RK], cy.h2, cy.rh2);
        };
        auto dZ = [&](const double* B, int idx, unsigned cf) -> double {
            if (cf & 2u) return edge3(cz.a0, B[idx], cz.b0, B[idx + 1], cz.c0, B[idx + 2]);
            if (cf & 4u) return edge3(cz.a1, B[idx - 2], cz.b1, B[idx - 1], cz.c1, B[idx]);
            return cdiv(B[idx + 1] - B[idx - 1], cz.h2, cz.rh2);
        };

        // ---- second derivatives, rho, sums at the own output columns of plane i ---------
        auto output = [&](int i) {
            const double* const px_i = px_ring + (i % 3) * PLANE;
            const double xxi = P.xx[i];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                if (!(rowf[b] & 16u)) continue;
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    if (!(colf[a] & 16u)) continue;
                    const int idx = jr_[b] * RK + tx + 32 * a;
                    double pxx;
                    if (i == 0)
                        pxx = edge3(cx.a0, px_ring[idx], cx.b0, px_ring[PLANE + idx], cx.c0, px_ring[2 * PLANE + idx]);
                    else if (i == P.n0 - 1)
                        pxx = edge3(cx.a1, px_ring[((i - 2) % 3) * PLANE + idx], cx.b1,
                                    px_ring[((i - 1) % 3) * PLANE + idx], cx.c1, px_i[idx]);
                    else
                        pxx = cdiv(px_ring[((i + 1) % 3) * PLANE + idx] - px_ring[((i + 2) % 3) * PLANE + idx],
                                   cx.h2, cx.rh2);
                    const double pxy = dY(px_i, idx, rowf[b]);
                    const double pxz = dZ(px_i, idx, colf[a]);
                    const double pyy = dY(py_buf, idx, rowf[b]);
                    const double pyz = dZ(py_buf, idx, colf[a]);
                    const double pzz = dZ(pz_buf, idx, colf[a]);
                    // K_ij = -phi_ij: the sign cancels in K^2 and in K_ij K^ij (reproduce_rodal_exact.py:198-212)
                    const double tr = (pxx + pyy) + pzz;
                    const double diag = (pxx * pxx + pyy * pyy) + pzz * pzz;
                    const double off = (pxy * pxy + pxz * pxz) + pyz * pyz;
                    const double kk = fma(2.0, off, diag);     // 2.0*off is exact -> same rounding as diag + 2.0*off
                    const double rho_adm = cdiv(tr * tr - kk, P.c16pi, P.rc16pi);
                    const double e = rho_adm * P.dV;
                    if (EMIT) {
                        const long long j = j0 - 2 + jr_[b], k = k0 - 2 + tx + 32 * a;
                        P.rho_out[((long long)(i - P.i_begin) * P.n1 + j) * P.n2 + k] = rho_adm;
                    }
                    const bool pos = e > 0.0, neg = e < 0.0;
                    acc_pos += pos ? e : 0.0;
                    acc_neg += neg ? e : 0.0;
                    c_pos += pos; c_neg += neg; c_zero += (e == 0.0); c_nan += (e != e);
                    if (NSH > 0) {
                        bool in_[NSH > 0 ? NSH : 1];
                        if (DT == IWB_F64) {
                            const double s = (xxi + yyv[b]) + zzv[a];
#pragma unroll
                            for (int sh = 0; sh < NSH; ++sh) in_[sh] = s <= P.shell_thr[sh];
                        } else {
                            const float s = __fadd_rn(__fadd_rn((float)xxi, (float)yyv[b]), (float)zzv[a]);
#pragma unroll
                            for (int sh = 0; sh < NSH; ++sh) in_[sh] = s <= (float)P.shell_thr[sh];
                        }
#pragma unroll
                        for (int sh = 0; sh < NSH; ++sh) {
                            sh_pos[sh] += (in_[sh] && pos) ? e : 0.0;
                            sh_neg[sh] += (in_[sh] && neg) ? e : 0.0;
                            sh_cnt[sh] += in_[sh];
                        }
                    }
                }
            }
        };

        // ---- block-level flush of the register sums into partial[fb][tile] ----------------
        auto flush = [&](int fb) {
            constexpr int NV = 6 + 3 * NSH;    // values actually reduced
            double vals[NV];
            vals[0] = acc_pos; vals[1] = acc_neg;
            vals[2] = __longlong_as_double((long long)c_pos);
            vals[3] = __longlong_as_double((long long)c_neg);
            vals[4] = __longlong_as_double((long long)c_zero);
            vals[5] = __longlong_as_double((long long)c_nan);
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
            double* scratch = pz_buf;           // free after the trailing barrier of output()
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
                }
                P.partial[((size_t)fb * ntiles + tile) * ROW + slot] = acc;
            }
            __syncthreads();
            acc_pos = 0.0; acc_neg = 0.0; c_pos = c_neg = c_zero = c_nan = 0;
#pragma unroll
            for (int s = 0; s < NSH; ++s) { sh_pos[s] = 0.0; sh_neg[s] = 0.0; sh_cnt[s] = 0; }
        };

        // ---- march ------------------------------------------------------------------------
        __syncthreads();                 // previous item's shared memory is dead
        int produced = pstart - 1;
        bool px_last_done = false;
        for (int i = i0; i < i1; ++i) {
            const int target = min(max(i + 2, 3), P.n0 - 1);
            const int early = min(target, i + 2);
            while (produced < early) produce(++produced);
            if (i == i0) __syncthreads();            // prologue planes visible to neighbours
            inplane(i);
            if (produced < target) {                 // only i == 0: phi(3) reuses the slot of phi(0)
                __syncthreads();
                produce(++produced);
            }
            if (!px_last_done && max(i - 1, 0) >= P.n0 - 3 && produced == P.n0 - 1) {
                produce_px_last();
                px_last_done = true;
            }
            __syncthreads();
            output(i);
            __syncthreads();
            if (((i + 1) % FLUSH) == 0 || i == i1 - 1) flush(i / FLUSH);
        }
    }
}

// ---- fixed-order sum over tiles: partial[fb][tile][ROW] -> table[fb][ROW] ----------------------
// (static: every translation unit that includes this header gets its own copy)
static __global__ void __launch_bounds__(128) k_reduce_tiles(const double* __restrict__ partial,
                                                             double* __restrict__ table, int fb_begin, int ntiles)
{
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

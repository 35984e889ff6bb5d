// Planar siblings of the 3D evaluator.  Compiled with -fmad=false (IEEE replay).
//
//   iwb_energy_axisym <- compute_energy_axisymmetric
//                        (/root/reference/scripts/reproduce_rodal_exact.py:76-154)
//   iwb_slice_z0      <- compute_slice_z0 + central_diff_2d + integrate_signed
//                        (/root/reference/src/irrotational_warp/adm.py:27-97, fd.py:6-25,
//                         potential.py:6-34)
//
// Both grids are [j (y), i (x)] with x contiguous.  The axisymmetric kernel is one fused
// pass per 28 x 60 tile (phi, first and second derivatives staged in shared memory); the
// slice evaluator has to return phi, beta and rho to the caller, so it is three small
// kernels over HBM-resident fields.
#include <math.h>
#include <string.h>

#include <chrono>
#include <string>

#include "energy3d.cuh"
#include "host_common.h"

namespace iwb {

// =================================================================================================
// axisymmetric
// =================================================================================================
struct AxParams {
    int nx, ny, ntj, ntk;
    const double* x;     // [nx]
    const double* xx;    // [nx] x*x
    const double* y;     // [ny]
    const double* yy;    // [ny] y*y
    const double* dV;    // [ny] ((2 pi * y) * dx) * dy, rounded like the reference's array expression
    PhiConst pc;
    AxisCoef cx, cy;
    double c16pi, rc16pi;
    double shell_thr[IWB_MAX_SHELLS];
    double* partial;     // [ntiles][ROW]
    double* rho_out;     // NULL or [ny][nx]
};

constexpr int AX_RJ = 32, AX_NW = 8, AX_NB = AX_RJ / AX_NW, AX_PLANE = AX_RJ * RK;

template <int DT, int NSH, bool EMIT>
__global__ void __launch_bounds__(AX_NW * 32, 3) k_axisym(const __grid_constant__ AxParams P)
{
    __shared__ double ph[AX_PLANE];
    __shared__ double px[AX_PLANE];
    __shared__ double py[AX_PLANE];
    constexpr int NB = AX_NB, NW = AX_NW;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int tile = blockIdx.x;
    const int tj = tile / P.ntk, tk = tile - tj * P.ntk;
    const int j0 = tile_begin(tj, P.ntj, P.ny), Tj = tile_begin(tj + 1, P.ntj, P.ny) - j0;
    const int i0 = tile_begin(tk, P.ntk, P.nx), Tk = tile_begin(tk + 1, P.ntk, P.nx) - i0;
    const AxisCoef cx = P.cx, cy = P.cy;

    double yyv[NB], yv[NB], dVv[NB], xxv[2], xv[2];
    int jr_[NB];
    unsigned rowf[NB], colf[2];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        int jr = ty + NW * b, j = j0 - 2 + jr;
        bool valid = (j >= 0) && (j < P.ny) && (jr < Tj + 4);
        int jc = min(max(j, 0), P.ny - 1);
        jr_[b] = jr; yyv[b] = P.yy[jc]; yv[b] = P.y[jc]; dVv[b] = P.dV[jc];
        rowf[b] = (valid ? 1u : 0u) | ((j == 0) ? 2u : 0u) | ((j == P.ny - 1) ? 4u : 0u) |
                  ((valid && jr >= 1 && jr <= Tj + 2) ? 8u : 0u) | ((jr >= 2 && jr <= Tj + 1) ? 16u : 0u);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        int kr = tx + 32 * a, i = i0 - 2 + kr;
        bool valid = (i >= 0) && (i < P.nx) && (kr < Tk + 4);
        int ic = min(max(i, 0), P.nx - 1);
        xxv[a] = P.xx[ic]; xv[a] = P.x[ic];
        colf[a] = (valid ? 1u : 0u) | ((i == 0) ? 2u : 0u) | ((i == P.nx - 1) ? 4u : 0u) |
                  ((valid && kr >= 1 && kr <= Tk + 2) ? 8u : 0u) | ((kr >= 2 && kr <= Tk + 1) ? 16u : 0u);
    }

    // ---- phase 1: phi on the region (z = 0: r = sqrt((x*x + y*y) + 0)) ---------------------------
    // The NB rows of one column share x: NB independent evaluations per call, so their sqrt / division chains overlap.
    // The host's operand-range proof (phi_fast_path_is_safe, as for the 3D evaluator) covers every point but the origin
    // r == 0; the thread that holds it, and every thread of an exotic grid, takes the reference flavour.  (Measured on
    // B200, 2400 x 1200 / 4800 x 2400, event time with the tile reduction: one point at a time with per-point range tests
    // 0.079 / 0.270 ms; grouped with per-point range tests 0.114 / 0.405 ms; grouped + host proof + seeded divisions
    // 0.074 / 0.249 ms; the same with three CTAs per SM 0.067 / 0.223 ms -- profiles/r2_ab17_axisym.log.)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        double f[NB];
        bool force_ref = !P.pc.prechecked;
#pragma unroll
        for (int b = 0; b < NB; ++b) force_ref = force_ref || (xxv[a] == 0.0 && yyv[b] == 0.0);
        if (DT == IWB_F64) {
            double s[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) s[b] = xxv[a] + yyv[b];
            phi_exact_f64<NB, true>(s, xv[a], P.pc, f, force_ref);
        } else {
            float s[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) s[b] = __fadd_rn((float)xxv[a], (float)yyv[b]);
            phi_exact_f32ref<NB, true>(s, (float)xv[a], P.pc, f, force_ref);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) ph[jr_[b] * RK + tx + 32 * a] = f[b];
    }
    __syncthreads();

    auto dY = [&](const double* B, int idx, unsigned rf) -> double {
        if (rf & 2u) return edge3(cy.a0, B[idx], cy.b0, B[idx + RK], cy.c0, B[idx + 2 * RK]);
        if (rf & 4u) return edge3(cy.a1, B[idx - 2 * RK], cy.b1, B[idx - RK], cy.c1, B[idx]);
        return cdiv(B[idx + RK] - B[idx - RK], cy.h2, cy.rh2);
    };
    auto dX = [&](const double* B, int idx, unsigned cf) -> double {
        if (cf & 2u) return edge3(cx.a0, B[idx], cx.b0, B[idx + 1], cx.c0, B[idx + 2]);
        if (cf & 4u) return edge3(cx.a1, B[idx - 2], cx.b1, B[idx - 1], cx.c1, B[idx]);
        return cdiv(B[idx + 1] - B[idx - 1], cx.h2, cx.rh2);
    };

    // ---- phase 2: first derivatives where the second ones will need them ---------------------------
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int idx = jr_[b] * RK + tx + 32 * a;
            if ((rowf[b] & 8u) && (colf[a] & 1u)) py[idx] = dY(ph, idx, rowf[b]);
            if ((rowf[b] & 1u) && (colf[a] & 8u)) px[idx] = dX(ph, idx, colf[a]);
        }
    }
    __syncthreads();

    // ---- phase 3: K_ij, rho, sums (reproduce_rodal_exact.py:115-143) -----------------------------------
    double acc_pos = 0.0, acc_neg = 0.0;
    double sh_pos[NSH > 0 ? NSH : 1], sh_neg[NSH > 0 ? NSH : 1];
    int c_pos = 0, c_neg = 0, c_zero = 0, c_nan = 0, sh_cnt[NSH > 0 ? NSH : 1];
#pragma unroll
    for (int s = 0; s < NSH; ++s) { sh_pos[s] = 0.0; sh_neg[s] = 0.0; sh_cnt[s] = 0; }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (!(rowf[b] & 16u)) continue;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            if (!(colf[a] & 16u)) continue;
            const int idx = jr_[b] * RK + tx + 32 * a;
            const int j = j0 - 2 + jr_[b], i = i0 - 2 + tx + 32 * a;
            const double kxx = -dX(px, idx, colf[a]);
            const double kyy = -dY(py, idx, rowf[b]);
            const double kxy = -dX(py, idx, colf[a]);          // d/dx of phi_y (line 113)
            const double kzz = (yv[b] != 0.0 && j != 0) ? (-py[idx]) / yv[b] : kyy;
            const double tr = (kxx + kyy) + kzz;
            const double diag = (kxx * kxx + kyy * kyy) + kzz * kzz;
            const double kk = fma(2.0, kxy * kxy, diag);       // (2.0*kxy)*kxy == 2*RN(kxy^2), exact doubling
            const double rho_adm = cdiv(tr * tr - kk, P.c16pi, P.rc16pi);
            const double e = rho_adm * dVv[b];
            if (EMIT) P.rho_out[(size_t)j * P.nx + i] = rho_adm;
            const bool pos = e > 0.0, neg = e < 0.0;
            acc_pos += pos ? e : 0.0;
            acc_neg += neg ? e : 0.0;
            c_pos += pos; c_neg += neg; c_zero += (e == 0.0); c_nan += (e != e);
            if (NSH > 0) {
#pragma unroll
                for (int sh = 0; sh < NSH; ++sh) {
                    bool in;
                    if (DT == IWB_F64) in = (xxv[a] + yyv[b]) <= P.shell_thr[sh];
                    else in = __fadd_rn((float)xxv[a], (float)yyv[b]) <= (float)P.shell_thr[sh];
                    sh_pos[sh] += (in && pos) ? e : 0.0;
                    sh_neg[sh] += (in && neg) ? e : 0.0;
                    sh_cnt[sh] += in;
                }
            }
        }
    }

    // ---- block reduction into partial[tile] (fixed order) ------------------------------------------------
    constexpr int NV = 6 + 3 * NSH;
    double vals[NV];
    vals[0] = acc_pos; vals[1] = acc_neg;
    vals[2] = __longlong_as_double((long long)c_pos); vals[3] = __longlong_as_double((long long)c_neg);
    vals[4] = __longlong_as_double((long long)c_zero); vals[5] = __longlong_as_double((long long)c_nan);
#pragma unroll
    for (int s = 0; s < NSH; ++s) {
        vals[6 + s] = sh_pos[s]; vals[6 + NSH + s] = sh_neg[s];
        vals[6 + 2 * NSH + s] = __longlong_as_double((long long)sh_cnt[s]);
    }
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
    double* const scratch = ph;     // phi plane is dead after phase 2 (barrier above)
    if (tx == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) scratch[ty * NV + q] = vals[q];
    }
    __syncthreads();
    if (threadIdx.x < ROW) {
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
        P.partial[(size_t)tile * ROW + slot] = acc;
    }
}

// =================================================================================================
// 2D z = 0 slice, tanh-wall dipole, edge_order = 1
// =================================================================================================
struct SlParams {
    int n;
    const double *x, *y;       // [n]
    double rho, sigma, eps, v_rho;   // v_rho = v*rho (one Python-float product, potential.py:34)
    double hx, rhx, h2x, rh2x, hy, rhy, h2y, rh2y;
    double c16pi, rc16pi;
    double *phi, *bx, *by, *rho_adm;   // device [n][n]
    double* partial;           // [nblocks][ROW]
};

// numpy.gradient, edge_order=1 (fd.py:11-12): forward/backward difference over h at the ends
__device__ __forceinline__ double d1(const double* f, int p, int n, int stride, double h, double rh, double h2, double rh2)
{
    if (p == 0) return cdiv(f[stride] - f[0], h, rh);
    if (p == n - 1) return cdiv(f[0] - f[-stride], h, rh);
    return cdiv(f[stride] - f[-stride], h2, rh2);
}

__global__ void __launch_bounds__(256) k_slice_phi(const SlParams P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= P.n) return;
    const double x = P.x[i], y = P.y[j];
    const double r = sqrt(x * x + y * y)   /* + z*z with z == 0 is exact */;
    const double xi = r / P.rho;
    const double f = 0.5 * (1.0 + tanh(P.sigma * (1.0 - xi)));   // tanh_wall, potential.py:11-12
    const double c = x / (r + P.eps);
    P.phi[(size_t)j * P.n + i] = (P.v_rho * f) * c;
}

__global__ void __launch_bounds__(256) k_slice_beta(const SlParams P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= P.n) return;
    const size_t q = (size_t)j * P.n + i;
    P.bx[q] = d1(P.phi + q, i, P.n, 1, P.hx, P.rhx, P.h2x, P.rh2x);
    P.by[q] = d1(P.phi + q, j, P.n, P.n, P.hy, P.rhy, P.h2y, P.rh2y);
}

__global__ void __launch_bounds__(256) k_slice_rho(const SlParams P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    double pos = 0.0, neg = 0.0;
    int c_pos = 0, c_neg = 0, c_zero = 0;
    if (i < P.n) {
        const size_t q = (size_t)j * P.n + i;
        const double kxx = d1(P.bx + q, i, P.n, 1, P.hx, P.rhx, P.h2x, P.rh2x);
        const double bx_y = d1(P.bx + q, j, P.n, P.n, P.hy, P.rhy, P.h2y, P.rh2y);
        const double by_x = d1(P.by + q, i, P.n, 1, P.hx, P.rhx, P.h2x, P.rh2x);
        const double kyy = d1(P.by + q, j, P.n, P.n, P.hy, P.rhy, P.h2y, P.rh2y);
        const double kxy = 0.5 * (bx_y + by_x);
        const double tr = kxx + kyy;
        const double kk = fma(2.0, kxy * kxy, kxx * kxx + kyy * kyy);   // (2.0*kxy)*kxy is an exact doubling
        const double rho_adm = cdiv(tr * tr - kk, P.c16pi, P.rc16pi);
        P.rho_adm[q] = rho_adm;
        if (rho_adm > 0.0) { pos = rho_adm; c_pos = 1; }
        if (rho_adm < 0.0) { neg = -rho_adm; c_neg = 1; }
        c_zero = (rho_adm == 0.0);
    }
    // block reduction (fixed order) -> partial[block]
    __shared__ double sp[8], sn[8];
    __shared__ int sc[8][3];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        pos += __shfl_xor_sync(0xffffffffu, pos, off);
        neg += __shfl_xor_sync(0xffffffffu, neg, off);
        c_pos += __shfl_xor_sync(0xffffffffu, c_pos, off);
        c_neg += __shfl_xor_sync(0xffffffffu, c_neg, off);
        c_zero += __shfl_xor_sync(0xffffffffu, c_zero, off);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sp[w] = pos; sn[w] = neg; sc[w][0] = c_pos; sc[w][1] = c_neg; sc[w][2] = c_zero; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        long long c0 = 0, c1 = 0, c2 = 0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { a += sp[q]; b += sn[q]; c0 += sc[q][0]; c1 += sc[q][1]; c2 += sc[q][2]; }
        double* row = P.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * ROW;
        for (int q = 0; q < ROW; ++q) row[q] = 0.0;
        row[0] = a; row[1] = b;
        row[2] = __longlong_as_double(c0); row[3] = __longlong_as_double(c1); row[4] = __longlong_as_double(c2);
    }
}

template <int DT, bool EMIT>
static void launch_axisym(int nshell, const AxParams& P, int ntiles, cudaStream_t st)
{
    if (nshell <= 2) k_axisym<DT, 2, EMIT><<<ntiles, AX_NW * 32, 0, st>>>(P);
    else k_axisym<DT, IWB_MAX_SHELLS, EMIT><<<ntiles, AX_NW * 32, 0, st>>>(P);
}

static void unpack_row(const double* row, int nshell, iwb_result* out)
{
    auto as_i64 = [](double d) { int64_t v; memcpy(&v, &d, 8); return v; };
    out->e_pos = row[0]; out->e_neg = row[1];
    out->cnt_pos = as_i64(row[2]); out->cnt_neg = as_i64(row[3]);
    out->cnt_zero = as_i64(row[4]); out->cnt_nan = as_i64(row[5]);
    for (int s = 0; s < IWB_MAX_SHELLS; ++s) {
        const bool used = s < nshell;
        out->shell_pos[s] = used ? row[6 + s] : 0.0;
        out->shell_neg[s] = used ? row[6 + IWB_MAX_SHELLS + s] : 0.0;
        out->shell_cnt[s] = used ? as_i64(row[6 + 2 * IWB_MAX_SHELLS + s]) : 0;
    }
}

}  // namespace iwb

using namespace iwb;

extern "C" {

int iwb_energy_axisym(iwb_handle* h, const iwb_params_axisym* p, iwb_result* out)
{
    if (!h || !p || !out) return fail(IWB_ERR_INVALID, "iwb_energy_axisym: null argument");
    if (p->dtype != IWB_F64 && p->dtype != IWB_F32_REF) return fail(IWB_ERR_INVALID, "axisym: dtype must be IWB_F64 or IWB_F32_REF");
    if (p->mode != IWB_MODE_EXACT) return fail(IWB_ERR_INVALID, "axisym: only IWB_MODE_EXACT is built");
    if (p->nx < 3 || p->ny < 3)
        return fail(IWB_ERR_GRID_TOO_SMALL, "Shape of array too small to calculate a numerical gradient, "
                                            "at least (edge_order + 1) elements are required.");
    if (!p->x || !p->y) return fail(IWB_ERR_INVALID, "axisym: coordinate arrays are NULL");
    if (p->nshell < 0 || p->nshell > IWB_MAX_SHELLS) return fail(IWB_ERR_INVALID, "axisym: nshell out of range");
    if (!(p->dx > 0.0) || !(p->dy > 0.0)) return fail(IWB_ERR_INVALID, "axisym: spacings must be > 0");
    auto t0 = std::chrono::steady_clock::now();
    IWB_CUDA(cudaSetDevice(h->device));
    Scratch& c = h->scr[0];
    const int nx = p->nx, ny = p->ny;
    const size_t ncoord = (size_t)2 * nx + 3 * ny;
    IWB_CUDA(c.h_coords.reserve(ncoord * sizeof(double)));
    IWB_CUDA(c.coords.reserve(ncoord * sizeof(double)));
    IWB_CUDA(cudaStreamSynchronize(c.stream));
    double* hc = (double*)c.h_coords.p;
    double *hx = hc, *hxx = hc + nx, *hy = hc + 2 * nx, *hyy = hy + ny, *hdV = hyy + ny;
    const double two_pi = 2.0 * 3.141592653589793;
    if (p->dtype == IWB_F64) {
        for (int i = 0; i < nx; ++i) { hx[i] = p->x[i]; hxx[i] = p->x[i] * p->x[i]; }
        for (int j = 0; j < ny; ++j) {
            hy[j] = p->y[j]; hyy[j] = p->y[j] * p->y[j];
            hdV[j] = two_pi * p->y[j] * p->dx * p->dy;          // 2.0*np.pi*Y*dx*dy, left to right (line 133)
        }
    } else {
        for (int i = 0; i < nx; ++i) { float f = (float)p->x[i]; hx[i] = f; hxx[i] = (double)(f * f); }
        for (int j = 0; j < ny; ++j) {
            float f = (float)p->y[j];
            hy[j] = f; hyy[j] = (double)(f * f);
            hdV[j] = (double)((float)two_pi * f * (float)p->dx * (float)p->dy);   // stays float32 (weak scalars)
        }
    }
    IWB_CUDA(cudaMemcpyAsync(c.coords.p, hc, ncoord * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    AxParams P;
    memset(&P, 0, sizeof(P));
    P.nx = nx; P.ny = ny;
    P.ntj = (ny + (AX_RJ - 4) - 1) / (AX_RJ - 4);
    P.ntk = (nx + TK_MAX - 1) / TK_MAX;
    const int ntiles = P.ntj * P.ntk;
    double* dc = (double*)c.coords.p;
    P.x = dc; P.xx = dc + nx; P.y = dc + 2 * nx; P.yy = P.y + ny; P.dV = P.yy + ny;
    P.pc = make_phi_const(p->rho, p->sigma, p->v, p->eps, p->sinh_rs, p->cosh_rs);
    {   // operand ranges of the branch-free phi on this half-plane (z == 0), as for the 3D evaluator
        const double zero = 0.0;
        const double* const coords[3] = {p->x, p->y, &zero};
        const int nn[3] = {nx, ny, 1};
        int zero_index[3];
        P.pc.prechecked = phi_fast_path_is_safe(coords, nn, p->rho, p->sigma, p->v, p->eps, p->sinh_rs, p->cosh_rs,
                                                zero_index, p->dtype != IWB_F64) ? 1 : 0;
    }
    P.cx = make_axis_coef(p->dx); P.cy = make_axis_coef(p->dy);
    P.c16pi = 16.0 * 3.141592653589793; P.rc16pi = 1.0 / P.c16pi;
    for (int s = 0; s < IWB_MAX_SHELLS; ++s)
        P.shell_thr[s] = (s < p->nshell) ? ((p->dtype == IWB_F64) ? shell_threshold_f64(p->shell_r[s])
                                                                   : shell_threshold_f32(p->shell_r[s])) : -1.0;
    IWB_CUDA(c.partial.reserve((size_t)ntiles * ROW * sizeof(double)));
    IWB_CUDA(c.out.reserve(ROW_BYTES_MIN));
    IWB_CUDA(c.h_out.reserve(ROW_BYTES_MIN));
    P.partial = (double*)c.partial.p;
    const size_t rho_bytes = (size_t)nx * ny * sizeof(double);
    if (p->rho_out) { IWB_CUDA(c.rho.reserve(rho_bytes)); P.rho_out = (double*)c.rho.p; }
    IWB_CUDA(cudaEventRecord(c.ev0, c.stream));
    if (p->dtype == IWB_F64) {
        if (p->rho_out) launch_axisym<IWB_F64, true>(p->nshell, P, ntiles, c.stream);
        else launch_axisym<IWB_F64, false>(p->nshell, P, ntiles, c.stream);
    } else {
        if (p->rho_out) launch_axisym<IWB_F32_REF, true>(p->nshell, P, ntiles, c.stream);
        else launch_axisym<IWB_F32_REF, false>(p->nshell, P, ntiles, c.stream);
    }
    IWB_CUDA(cudaGetLastError());
    k_reduce_tiles<<<1, 1024, 0, c.stream>>>(P.partial, (double*)c.out.p, 0, ntiles);
    IWB_CUDA(cudaGetLastError());
    IWB_CUDA(cudaMemcpyAsync(c.h_out.p, c.out.p, ROW_BYTES_MIN, cudaMemcpyDeviceToHost, c.stream));
    IWB_CUDA(cudaEventRecord(c.ev1, c.stream));
    if (p->rho_out) IWB_CUDA(cudaMemcpyAsync(p->rho_out, P.rho_out, rho_bytes, cudaMemcpyDeviceToHost, c.stream));
    IWB_CUDA(cudaStreamSynchronize(c.stream));
    memset(out, 0, sizeof(*out));
    unpack_row((const double*)c.h_out.p, p->nshell, out);
    apply_reference_nonfinite(out, p->nshell);
    float ms = 0;
    IWB_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
    out->kernel_ms = ms;
    out->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return IWB_OK;
}

int iwb_slice_z0(iwb_handle* h, const iwb_params_slice* p, iwb_slice_result* out)
{
    if (!h || !p || !out) return fail(IWB_ERR_INVALID, "iwb_slice_z0: null argument");
    if (p->n < 2)
        return fail(IWB_ERR_GRID_TOO_SMALL, "Shape of array too small to calculate a numerical gradient, "
                                            "at least (edge_order + 1) elements are required.");
    if (!p->x || !p->y || !p->phi || !p->beta_x || !p->beta_y || !p->rho_adm)
        return fail(IWB_ERR_INVALID, "iwb_slice_z0: all coordinate and field buffers must be non-NULL");
    if (!(p->dx > 0.0) || !(p->dy > 0.0)) return fail(IWB_ERR_INVALID, "slice: spacings must be > 0");
    auto t0 = std::chrono::steady_clock::now();
    IWB_CUDA(cudaSetDevice(h->device));
    Scratch& c = h->scr[0];
    const int n = p->n;
    const size_t nn = (size_t)n * n;
    IWB_CUDA(c.h_coords.reserve((size_t)2 * n * sizeof(double)));
    IWB_CUDA(c.coords.reserve((size_t)2 * n * sizeof(double)));
    IWB_CUDA(c.rho.reserve(4 * nn * sizeof(double)));
    IWB_CUDA(cudaStreamSynchronize(c.stream));
    double* hc = (double*)c.h_coords.p;
    memcpy(hc, p->x, n * sizeof(double));
    memcpy(hc + n, p->y, n * sizeof(double));
    IWB_CUDA(cudaMemcpyAsync(c.coords.p, hc, (size_t)2 * n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    SlParams P;
    memset(&P, 0, sizeof(P));
    P.n = n;
    P.x = (double*)c.coords.p; P.y = P.x + n;
    P.rho = p->rho; P.sigma = p->sigma; P.eps = p->eps; P.v_rho = p->v * p->rho;
    P.hx = p->dx; P.rhx = 1.0 / P.hx; P.h2x = 2. * p->dx; P.rh2x = 1.0 / P.h2x;
    P.hy = p->dy; P.rhy = 1.0 / P.hy; P.h2y = 2. * p->dy; P.rh2y = 1.0 / P.h2y;
    P.c16pi = 16.0 * 3.141592653589793; P.rc16pi = 1.0 / P.c16pi;
    double* f = (double*)c.rho.p;
    P.phi = f; P.bx = f + nn; P.by = f + 2 * nn; P.rho_adm = f + 3 * nn;
    dim3 grid((n + 255) / 256, n);
    const int nblocks = grid.x * grid.y;
    IWB_CUDA(c.partial.reserve((size_t)nblocks * ROW * sizeof(double)));
    IWB_CUDA(c.out.reserve(ROW_BYTES_MIN));
    IWB_CUDA(c.h_out.reserve(ROW_BYTES_MIN));
    P.partial = (double*)c.partial.p;
    IWB_CUDA(cudaEventRecord(c.ev0, c.stream));
    k_slice_phi<<<grid, 256, 0, c.stream>>>(P);
    k_slice_beta<<<grid, 256, 0, c.stream>>>(P);
    k_slice_rho<<<grid, 256, 0, c.stream>>>(P);
    k_reduce_tiles<<<1, 1024, 0, c.stream>>>(P.partial, (double*)c.out.p, 0, nblocks);
    IWB_CUDA(cudaGetLastError());
    IWB_CUDA(cudaMemcpyAsync(c.h_out.p, c.out.p, ROW_BYTES_MIN, cudaMemcpyDeviceToHost, c.stream));
    IWB_CUDA(cudaEventRecord(c.ev1, c.stream));
    IWB_CUDA(cudaMemcpyAsync(p->phi, P.phi, nn * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IWB_CUDA(cudaMemcpyAsync(p->beta_x, P.bx, nn * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IWB_CUDA(cudaMemcpyAsync(p->beta_y, P.by, nn * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IWB_CUDA(cudaMemcpyAsync(p->rho_adm, P.rho_adm, nn * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IWB_CUDA(cudaStreamSynchronize(c.stream));
    const double* row = (const double*)c.h_out.p;
    auto as_i64 = [](double d) { int64_t v; memcpy(&v, &d, 8); return v; };
    memset(out, 0, sizeof(*out));
    const double dA = p->dx * p->dy;                 // integrate_signed, adm.py:93-96
    out->pos = row[0] * dA; out->neg = row[1] * dA; out->net = out->pos - out->neg;
    out->cnt_pos = as_i64(row[2]); out->cnt_neg = as_i64(row[3]); out->cnt_zero = as_i64(row[4]);
    float ms = 0;
    IWB_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
    out->kernel_ms = ms;
    out->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return IWB_OK;
}

}  // extern "C"

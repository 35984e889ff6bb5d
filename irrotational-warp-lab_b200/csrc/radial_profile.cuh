// Radial-shell profile of rho_adm on the 3D grid: per-bin sum and point count over the spherical shells
// edges[k] <= r < edges[k+1].  This is the binning rule of compute_radial_average_z0
// (/root/reference/src/irrotational_warp/tail.py:46-80: equal-width bins from 0 to max r, half-open, so the farthest
// point belongs to no bin) applied to the 3D field of compute_energy_3d instead of the z = 0 slice.
//
// The fused kernel emits rho for a slab of planes (8 B/point, the only time the field reaches HBM); k_radial_hist reads
// it back once.  Deterministic by construction: a flush block of IWB_FLUSH_PLANES planes is split over PROF_SUB (64) CTAs
// and PROF_NW warps by row index; every warp owns a private histogram that only its lane 0 updates, one bin at a time
// (the lanes of one bin are added by a fixed shuffle tree); warps, CTAs and flush blocks are then added in index
// order.  The result therefore depends neither on scheduling nor on how the planes were cut into passes or slabs.
#pragma once
#include "energy3d.cuh"

namespace iwb {

constexpr int PROF_MAX_BINS = 256;
constexpr int PROF_SUB = 64;      // CTAs per flush block (fixed: the summation order must not depend on the pass size)
constexpr int PROF_NW = 8;        // warps per CTA

struct ProfParams {
    const double* rho;            // [i_end - i_begin][n1][n2]
    const double *xx, *yy, *zz;   // squared coordinates as staged for k_energy3d
    const double* edges;          // [nb + 1]
    int n1, n2, i_begin, i_end, nb, dtype;
    double inv_width;             // nb / (edges[nb] - edges[0]): first guess of the bin, corrected against the edges
    double* part_sum;             // [flush block][PROF_SUB][nb]
    long long* part_cnt;
};

static __global__ void __launch_bounds__(PROF_NW * 32) k_radial_hist(const ProfParams P)
{
    __shared__ double s_edges[PROF_MAX_BINS + 1];
    __shared__ double s_sum[PROF_NW][PROF_MAX_BINS];
    __shared__ int s_cnt[PROF_NW][PROF_MAX_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = P.nb;
    for (int q = threadIdx.x; q <= nb; q += blockDim.x) s_edges[q] = P.edges[q];
    for (int q = lane; q < nb; q += 32) { s_sum[warp][q] = 0.0; s_cnt[warp][q] = 0; }
    __syncthreads();

    const int sub = blockIdx.x;
    const int fb = P.i_begin / FLUSH + blockIdx.y;
    const int ia = fb * FLUSH, ib = min(ia + FLUSH, P.i_end);
    const int nrows = (ib - ia) * P.n1;
    for (int row = sub * PROF_NW + warp; row < nrows; row += PROF_SUB * PROF_NW) {
        const int i = ia + row / P.n1, j = row - (row / P.n1) * P.n1;
        const double* src = P.rho + ((size_t)(i - P.i_begin) * P.n1 + j) * P.n2;
        const double sxy = P.xx[i] + P.yy[j];
        const float sxy_f = __fadd_rn((float)P.xx[i], (float)P.yy[j]);
        for (int k0 = 0; k0 < P.n2; k0 += 32) {
            const int k = k0 + lane;
            const bool valid = k < P.n2;
            double v = 0.0, r = 0.0;
            if (valid) {
                v = src[k];
                // the same r as the evaluator's (reproduce_rodal_exact.py:179, 218)
                r = (P.dtype == IWB_F64) ? sqrt(sxy + P.zz[k]) : (double)__fsqrt_rn(__fadd_rn(sxy_f, (float)P.zz[k]));
            }
            // first guess from the mean bin width, then corrected against the edges themselves; the masks of
            // tail.py:75 are (R >= e[k]) & (R < e[k+1]), so a point below the first edge or on / beyond the last
            // one belongs to no bin
            int kb = max(0, min((int)((r - s_edges[0]) * P.inv_width), nb - 1));
            while (kb > 0 && r < s_edges[kb]) --kb;
            while (kb < nb && r >= s_edges[kb + 1]) ++kb;     // kb == nb: on or beyond the last edge, in no bin
            const bool in_bin = valid && kb < nb && r >= s_edges[0];
            unsigned todo = __ballot_sync(0xffffffffu, in_bin);
            while (todo) {
                const int b = __shfl_sync(0xffffffffu, kb, __ffs(todo) - 1);
                const bool mine = in_bin && kb == b;
                const unsigned same = __ballot_sync(0xffffffffu, mine);
                double t = mine ? v : 0.0;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
                if (lane == 0) { s_sum[warp][b] += t; s_cnt[warp][b] += __popc(same); }
                todo &= ~same;
            }
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nb; q += blockDim.x) {
        double a = s_sum[0][q];
        long long c = s_cnt[0][q];
        for (int w = 1; w < PROF_NW; ++w) { a += s_sum[w][q]; c += s_cnt[w][q]; }
        const size_t o = ((size_t)fb * PROF_SUB + sub) * nb + q;
        P.part_sum[o] = a;
        P.part_cnt[o] = c;
    }
}

// ---- the fast flavour (n_bins <= PROF_FAST_BINS): lane-private histograms, thresholds on s instead of a square root ----
// Round 1's kernel above spends ~10 ms on the 8.6 GB of an n = 1024^3 field (0.86 TB/s): a float64 sqrt per point and, per
// group of 32 points, one ballot + 10-shuffle reduction for every bin present in the group.  Here
//   * bin k is decided on s = (x*x + y*y) + z*z itself:  slo[k] <= s < slo[k+1]  with  slo[k] = min{s : RN(sqrt(s)) >= e[k]}
//     found by the host (the same points as the reference's masks (R >= e[k]) & (R < e[k+1]), tail.py:75; a float32 sqrt
//     only supplies the first guess), and
//   * every LANE owns a private histogram in shared memory ([warp][bin][lane]: conflict-free for any mix of bins), so a
//     point is one shared-memory read-modify-write of its sum and one of its 16-bit count, and no lane talks to another
//     (a first version counted with shared-memory atomics: 32 lanes on one address, no faster than round 1's kernel).
//     The lanes of a warp are folded by a fixed shuffle tree at the end, warps / CTAs / flush blocks in index order as
//     before: same bits on every run, for every pass size and slab decomposition.
constexpr int PROF_FAST_BINS = 64;
constexpr int PROF_FAST_NW = 10;       // 10 x 64 x 32 lane-private (double sum, 16-bit count) pairs = 200 KB
constexpr int PROF_FAST_UNROLL = 8;    // columns in flight per lane (10 warps x 32 lanes x 8 x 8 B = 20 KB per SM)
constexpr size_t PROF_FAST_SMEM = (size_t)PROF_FAST_NW * PROF_FAST_BINS * 32 * (sizeof(double) + sizeof(unsigned short));
// points one lane sees in one CTA (its 16-bit counts must hold them): rows per warp x 32-column groups per row
__host__ __device__ inline long long prof_fast_points_per_lane(int n1, int n2)
{
    const long long rows_per_warp = ((long long)FLUSH * n1 + PROF_SUB * PROF_FAST_NW - 1) / (PROF_SUB * PROF_FAST_NW);
    return rows_per_warp * ((n2 + 31) / 32);
}

struct ProfFastParams {
    const double* rho;            // [i_end - i_begin][n1][n2]
    const double *xx, *yy, *zz;   // squared coordinates as staged for k_energy3d
    const double* slo;            // [nb + 1] thresholds on s (float32 values on a float32 grid)
    int n1, n2, i_begin, i_end, nb, dtype;
    float e0, inv_width;          // first edge and nb / (e[nb] - e[0]): first guess of the bin from a float32 sqrt
    double* part_sum;             // [flush block][PROF_SUB][nb]
    long long* part_cnt;
};

static __global__ void __launch_bounds__(PROF_FAST_NW * 32) k_radial_hist_fast(const ProfFastParams P)
{
    extern __shared__ double h_sum[];                       // [PROF_FAST_NW][PROF_FAST_BINS][32], then the counts
    __shared__ double s_slo[PROF_FAST_BINS + 1];
    __shared__ double s_wsum[PROF_FAST_NW][PROF_FAST_BINS];
    __shared__ int s_wcnt[PROF_FAST_NW][PROF_FAST_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = P.nb;
    for (int q = threadIdx.x; q <= nb; q += blockDim.x) s_slo[q] = P.slo[q];
    double* const mine = h_sum + (size_t)warp * PROF_FAST_BINS * 32 + lane;        // bin b at mine[b * 32]
    unsigned short* const h_cnt = (unsigned short*)(h_sum + (size_t)PROF_FAST_NW * PROF_FAST_BINS * 32);
    unsigned short* const mine_c = h_cnt + (size_t)warp * PROF_FAST_BINS * 32 + lane;
    for (int b = 0; b < nb; ++b) { mine[b * 32] = 0.0; mine_c[b * 32] = 0; }
    __syncthreads();

    const int sub = blockIdx.x;
    const int fb = P.i_begin / FLUSH + blockIdx.y;
    const int ia = fb * FLUSH, ib = min(ia + FLUSH, P.i_end);
    const int nrows = (ib - ia) * P.n1;
    const bool f64 = (P.dtype == IWB_F64);
    for (int row = sub * PROF_FAST_NW + warp; row < nrows; row += PROF_SUB * PROF_FAST_NW) {
        const int i = ia + row / P.n1, j = row - (row / P.n1) * P.n1;
        const double* src = P.rho + ((size_t)(i - P.i_begin) * P.n1 + j) * P.n2;
        const double sxy = P.xx[i] + P.yy[j];
        const float sxy_f = __fadd_rn((float)P.xx[i], (float)P.yy[j]);
        for (int k0 = 0; k0 < P.n2; k0 += 32 * PROF_FAST_UNROLL) {
            double v[PROF_FAST_UNROLL], zz[PROF_FAST_UNROLL];
#pragma unroll
            for (int u = 0; u < PROF_FAST_UNROLL; ++u) {             // all loads of the group first
                const int k = k0 + u * 32 + lane;
                const bool valid = k < P.n2;
                v[u] = valid ? __ldcs(src + k) : 0.0;                // streamed: the field is read exactly once
                zz[u] = valid ? P.zz[k] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < PROF_FAST_UNROLL; ++u) {
                const int k = k0 + u * 32 + lane;
                // the evaluator's own s (reproduce_rodal_exact.py:179, 218: float32 arithmetic on a float32 grid)
                const double s = f64 ? sxy + zz[u] : (double)__fadd_rn(sxy_f, (float)zz[u]);
                int kb = max(0, min((int)((sqrtf((float)s) - P.e0) * P.inv_width), nb - 1));
                while (kb > 0 && s < s_slo[kb]) --kb;
                while (kb < nb && s >= s_slo[kb + 1]) ++kb;          // kb == nb: on or beyond the last edge, in no bin
                if (k < P.n2 && kb < nb && s >= s_slo[0]) {
                    mine[kb * 32] += v[u];
                    mine_c[kb * 32] += 1;
                }
            }
        }
    }
    __syncwarp();
    // lanes -> warp (fixed tree for the sums), then warps in index order
    for (int b = 0; b < nb; ++b) {
        double t = mine[b * 32];
        int c = mine_c[b * 32];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            t += __shfl_xor_sync(0xffffffffu, t, off);
            c += __shfl_xor_sync(0xffffffffu, c, off);
        }
        if (lane == 0) { s_wsum[warp][b] = t; s_wcnt[warp][b] = c; }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nb; q += blockDim.x) {
        double a = s_wsum[0][q];
        long long c = s_wcnt[0][q];
        for (int w = 1; w < PROF_FAST_NW; ++w) { a += s_wsum[w][q]; c += s_wcnt[w][q]; }
        const size_t o = ((size_t)fb * PROF_SUB + sub) * nb + q;
        P.part_sum[o] = a;
        P.part_cnt[o] = c;
    }
}

// ---- run-length flavour (n_bins <= PROF_FAST_BINS): the default ---------------------------------------------------------
// The lane-private kernel above is latency-bound (200 KB of bins leave 10 warps per SM; ncu: issue slots 27 % busy, a chain
// of dependent shared-memory read-modify-writes per lane; 1.19 ms per GiB = 0.9 TB/s).  Here a lane walks a CONTIGUOUS
// chunk of the row (columns [l * c, (l + 1) * c)), where r changes by at most dz per step, so its bin changes about once in
// bin_width / dz points: the lane keeps the current bin, that bin's two thresholds and the running (sum, count) in
// registers and touches nothing else until the bin changes.  Then the finished run is handed to the warp's histogram --
// 64 x 12 bytes of shared memory per warp, so an SM holds 64 warps -- in a fixed order: points in walking order, lanes in
// ascending order within a step, and the runs still open at the end of the CTA's rows in lane order.  Same bits on every
// run; warps / CTAs / flush blocks are added in index order as before.
constexpr int PROF_RLE_NW = 8;

static __global__ void __launch_bounds__(PROF_RLE_NW * 32) k_radial_hist_rle(const ProfFastParams P)
{
    __shared__ double s_slo[PROF_FAST_BINS + 2];
    __shared__ double s_sum[PROF_RLE_NW][PROF_FAST_BINS];
    __shared__ int s_cnt[PROF_RLE_NW][PROF_FAST_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = P.nb;
    for (int q = threadIdx.x; q <= nb; q += blockDim.x) s_slo[q] = P.slo[q];
    for (int q = lane; q < PROF_FAST_BINS; q += 32) { s_sum[warp][q] = 0.0; s_cnt[warp][q] = 0; }
    __syncthreads();

    const int sub = blockIdx.x;
    const int fb = P.i_begin / FLUSH + blockIdx.y;
    const int ia = fb * FLUSH, ib = min(ia + FLUSH, P.i_end);
    const int nrows = (ib - ia) * P.n1;
    const bool f64 = (P.dtype == IWB_F64);
    const int chunk = (P.n2 + 31) / 32;
    const int c0 = lane * chunk, c1 = min(c0 + chunk, P.n2);

    // the open run of this lane: bin cur (-1: none; nb: "in no bin"), its thresholds lo <= s < hi, and what it holds
    int cur = -1, run_cnt = 0;
    double lo = 0.0, hi = -1.0, run_sum = 0.0;
    // hand the runs of the lanes in `mask` to the warp's histogram, lowest lane first (fixed order)
    auto hand_over = [&](unsigned mask, int b_out, double v_out, int c_out) {
        while (mask) {
            const int src = __ffs(mask) - 1;
            const int b = __shfl_sync(0xffffffffu, b_out, src);
            const double v = __shfl_sync(0xffffffffu, v_out, src);
            const int c = __shfl_sync(0xffffffffu, c_out, src);
            if (lane == 0) { s_sum[warp][b] += v; s_cnt[warp][b] += c; }
            mask &= mask - 1;
        }
    };
    for (int row = sub * PROF_RLE_NW + warp; row < nrows; row += PROF_SUB * PROF_RLE_NW) {
        const int i = ia + row / P.n1, j = row - (row / P.n1) * P.n1;
        const double* src = P.rho + ((size_t)(i - P.i_begin) * P.n1 + j) * P.n2;
        const double sxy = P.xx[i] + P.yy[j];
        const float sxy_f = __fadd_rn((float)P.xx[i], (float)P.yy[j]);
        for (int u = 0; u < chunk; ++u) {                        // warp-uniform trip count
            const int k = c0 + u;
            const bool valid = k < c1;
            double v = 0.0, s = 0.0;
            if (valid) {
                v = src[k];                                      // 32-byte sectors are shared by four consecutive steps (L1)
                const double zz = P.zz[k];
                s = f64 ? sxy + zz : (double)__fadd_rn(sxy_f, (float)zz);
            }
            const bool same = valid && s >= lo && s < hi;        // still inside the open run's bin
            int b_out = 0, c_out = 0;
            double v_out = 0.0;
            if (valid && !same) {
                // close the open run, find the bin of s against the thresholds (exactly the reference's masks), open a new one
                b_out = cur; v_out = run_sum; c_out = run_cnt;
                int kb = (cur < 0) ? 0 : min(cur, nb - 1);
                while (kb > 0 && s < s_slo[kb]) --kb;
                while (kb < nb && s >= s_slo[kb + 1]) ++kb;
                if (s < s_slo[0]) { cur = nb; lo = -INFINITY; hi = s_slo[0]; }            // below the first edge: in no bin
                else if (kb >= nb) { cur = nb; lo = s_slo[nb]; hi = INFINITY; }             // on / beyond the last edge
                else { cur = kb; lo = s_slo[kb]; hi = s_slo[kb + 1]; }
                run_sum = 0.0; run_cnt = 0;
            }
            if (valid) { run_sum += v; ++run_cnt; }
            const unsigned mask = __ballot_sync(0xffffffffu, valid && !same && c_out > 0 && b_out >= 0 && b_out < nb);
            if (mask) hand_over(mask, b_out, v_out, c_out);
        }
    }
    hand_over(__ballot_sync(0xffffffffu, run_cnt > 0 && cur >= 0 && cur < nb), cur, run_sum, run_cnt);
    __syncthreads();
    for (int q = threadIdx.x; q < nb; q += blockDim.x) {
        double a = s_sum[0][q];
        long long c = s_cnt[0][q];
        for (int w = 1; w < PROF_RLE_NW; ++w) { a += s_sum[w][q]; c += s_cnt[w][q]; }
        const size_t o = ((size_t)fb * PROF_SUB + sub) * nb + q;
        P.part_sum[o] = a;
        P.part_cnt[o] = c;
    }
}

// partials [flush block][PROF_SUB][nb] -> [nb]: first the PROF_SUB CTAs of every flush block in index order (one block per
// flush block, one thread per bin), then the flush blocks in index order.  Two short dependent chains (64 + nfb additions per
// bin) instead of round 1's single one of 64 x nfb, which took 0.6 ms at n = 1024; the order is fixed either way.
static __global__ void k_radial_reduce_sub(const double* __restrict__ part_sum, const long long* __restrict__ part_cnt,
                                           int fb0, int nb, double* __restrict__ fb_sum, long long* __restrict__ fb_cnt)
{
    const int fb = fb0 + blockIdx.x;
    for (int q = threadIdx.x; q < nb; q += blockDim.x) {
        const size_t s0 = (size_t)fb * PROF_SUB;
        double a = 0.0;
        long long c = 0;
        for (int u = 0; u < PROF_SUB; u += 16) {
            double v[16];
            long long w[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) { v[e] = part_sum[(s0 + u + e) * nb + q]; w[e] = part_cnt[(s0 + u + e) * nb + q]; }
#pragma unroll
            for (int e = 0; e < 16; ++e) { a += v[e]; c += w[e]; }
        }
        fb_sum[(size_t)fb * nb + q] = a;
        fb_cnt[(size_t)fb * nb + q] = c;
    }
}

static __global__ void k_radial_reduce_fb(const double* __restrict__ fb_sum, const long long* __restrict__ fb_cnt,
                                          int fb0, int fb1, int nb, double* __restrict__ sum, long long* __restrict__ cnt)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nb) return;
    double a = 0.0;
    long long c = 0;
    for (int fb = fb0; fb < fb1; ++fb) { a += fb_sum[(size_t)fb * nb + q]; c += fb_cnt[(size_t)fb * nb + q]; }
    sum[q] = a;
    cnt[q] = c;
}

}  // namespace iwb

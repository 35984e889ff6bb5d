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
//
// Round 2 built and measured two replacements for k_radial_hist -- lane-private histograms with thresholds on s instead of
// a square root (200 KB of shared memory, 10 warps per SM: latency-bound, 1.19 ms per GiB), and per-lane run-length
// accumulation over contiguous chunks of a row (the strided loads cost 32 wavefronts each: 1.45 ms per GiB) -- against
// 1.08 ms per GiB for this kernel; neither is in the source (profiles/r2_f_rows_probe.log, git history).
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

// Fused path (k_energy3d_prof): per-unit windows [flush block][tile][PROF_WIN] with their first bin -> bins of every
// flush block, the tiles added in schedule order (one block per flush block, one thread per bin; eight loads in flight).
static __global__ void __launch_bounds__(PROF_FUSED_MAX_BINS) k_prof_reduce_windows(
    const double* __restrict__ wsum, const int* __restrict__ wcnt, const int* __restrict__ wbase, int fb0, int ntiles,
    int nb, double* __restrict__ fb_sum, long long* __restrict__ fb_cnt)
{
    extern __shared__ int s_base[];          // [ntiles]
    const int fb = fb0 + blockIdx.x;
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) s_base[t] = wbase[(size_t)fb * ntiles + t];
    __syncthreads();
    const int q = threadIdx.x;
    if (q >= nb) return;
    double a = 0.0;
    long long c = 0;
    for (int t0 = 0; t0 < ntiles; t0 += 8) {
        double v[8];
        int w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int t = t0 + u;
            const int rel = (t < ntiles) ? q - s_base[t] : -1;
            const bool in = rel >= 0 && rel < PROF_WIN;
            const size_t o = ((size_t)fb * ntiles + (in ? t : 0)) * PROF_WIN + (in ? rel : 0);
            v[u] = in ? wsum[o] : 0.0;
            w[u] = in ? wcnt[o] : -1;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (w[u] >= 0) { a += v[u]; c += w[u]; }
    }
    fb_sum[(size_t)fb * nb + q] = a;
    fb_cnt[(size_t)fb * nb + q] = c;
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

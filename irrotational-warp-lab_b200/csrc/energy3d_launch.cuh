// Launch plumbing of k_energy3d shared by the exact (-fmad=false) and fast (-fmad=true) translation units.
// Each unit instantiates the kernels for ONE value of MODE, so the two sets of host stubs never collide.
#pragma once
#include "energy3d.cuh"

namespace iwb {

// ---- kernel geometry variants -----------------------------------------------------------------
//   config 0: region 32 x 64, 16 warps (180 KB smem)
//   config 1: region 40 x 64, 20 warps (225 KB smem)  -- default
// One CTA per SM.  (Measured alternatives, n = 1024^3 fp64: 36x64/18 warps 72, 32x64/8 warps 52,
// two CTAs per SM of 20x64/10 warps 64 Gpoints/s against 77 for config 1; profiles/r1_k_energy3d_history.txt.)
struct K3Config { int rj, nw; };
static const K3Config kConfigs[] = {{32, 16}, {40, 20}};
constexpr int kNumConfigs = 2;

template <int DT, int NSH, int RJ, int NW, bool EMIT, int MINB, int MODE>
static cudaError_t launch_one(const K3Params& P, int sm_count, cudaStream_t st)
{
    auto kern = k_energy3d<DT, NSH, RJ, NW, EMIT, MINB, MODE>;
    constexpr size_t smem = Geo<RJ, NW>::SMEM_BYTES;
    // attribute and occupancy are properties of (instantiation, device): asked once per device, not per launch
    static int per_sm_cached[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int per_sm = (dev >= 0 && dev < 64) ? per_sm_cached[dev] : 0;
    if (per_sm == 0) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        if (dev >= 0 && dev < 64) per_sm_cached[dev] = per_sm;
    }
    int grid = sm_count * per_sm;
    if (grid > P.nspans) grid = P.nspans;
    kern<<<grid, NW * 32, smem, st>>>(P);
    return cudaGetLastError();
}

static bool config_is_built(int cfg, int dtype, int nshell, bool emit)
{
    (void)dtype; (void)nshell; (void)emit;
    return cfg == 0 || cfg == 1;       // every variant exists in both geometries: the tiling (and with it the
                                       // summation order) must not depend on emit_rho or on the number of shells
}

template <int DT, int NSH, bool EMIT, int MODE>
static cudaError_t launch_cfg(int cfg, const K3Params& P, int sm_count, cudaStream_t st)
{
    if (cfg == 1) return launch_one<DT, NSH, 40, 20, EMIT, 1, MODE>(P, sm_count, st);
    return launch_one<DT, NSH, 32, 16, EMIT, 1, MODE>(P, sm_count, st);
}

template <int DT, bool EMIT, int MODE>
static cudaError_t launch_nsh(int nshell, int cfg, const K3Params& P, int sm_count, cudaStream_t st)
{
    if (nshell <= 2) return launch_cfg<DT, 2, EMIT, MODE>(cfg, P, sm_count, st);
    return launch_cfg<DT, IWB_MAX_SHELLS, EMIT, MODE>(cfg, P, sm_count, st);
}

// fused radial profile: default geometry, <= 2 shells, exact mode, no rho emission (everything else takes the two-pass path)
template <int DT>
static cudaError_t launch_prof(const K3Params& P, int sm_count, cudaStream_t st)
{
    auto kern = k_energy3d_prof<DT, 2, 40, 20>;
    constexpr size_t smem = Geo<40, 20>::SMEM_BYTES + PROF_THR_S * sizeof(double);
    static int ready[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !ready[dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) ready[dev] = 1;
    }
    int grid = sm_count;
    if (grid > P.nspans) grid = P.nspans;
    kern<<<grid, 20 * 32, smem, st>>>(P);
    return cudaGetLastError();
}

// defined in energy3d_fast.cu (IWB_MODE_FAST, float64 only)
cudaError_t launch_k3_fast(int nshell, bool emit, int cfg, const K3Params& P, int sm_count, cudaStream_t st);

}  // namespace iwb

// IWB_MODE_FAST instantiations of the fused 3D kernel.  Compiled WITH FMA contraction (-fmad=true): the same
// source as the exact kernels with MODE = 1, i.e. divisions by constants become one multiply by the rounded
// reciprocal, the two IEEE divisions of phi share one Newton-refined reciprocal, and a*b+c contracts.
// Energies agree with the reference to ~1e-12 (bar: 1e-10); the sign of rho inside the rounding-noise core of
// the bubble -- and therefore the sign counts -- may differ from the reference, which is why
// this is a separate, opt-in mode and not the default.
#include "energy3d_launch.cuh"

namespace iwb {

cudaError_t launch_k3_fast(int nshell, bool emit, int cfg, const K3Params& P, int sm_count, cudaStream_t st)
{
    return emit ? launch_nsh<IWB_F64, true, 1>(nshell, cfg, P, sm_count, st)
                : launch_nsh<IWB_F64, false, 1>(nshell, cfg, P, sm_count, st);
}

}  // namespace iwb

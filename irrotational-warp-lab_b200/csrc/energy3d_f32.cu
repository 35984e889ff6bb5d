// IWB_F32_STRICT instantiations of the fused 3D kernel (phi evaluated entirely in float32 on the FP32 pipe, float64
// stencils on the float32-valued field).  A translation unit of its own so that it compiles in parallel with the float64 /
// float32-reference kernels; -fmad=false like them.
#include "energy3d_launch.cuh"

namespace iwb {

cudaError_t launch_k3_f32strict(int nshell, bool emit, int cfg, const K3Params& P, int sm_count, cudaStream_t st)
{
    return emit ? launch_nsh<IWB_F32_STRICT, true, 0>(nshell, cfg, P, sm_count, st)
                : launch_nsh<IWB_F32_STRICT, false, 0>(nshell, cfg, P, sm_count, st);
}

}  // namespace iwb

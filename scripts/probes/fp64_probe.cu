// FP64 pipe probe (development aid): dependent-issue latency and throughput of DFMA / DADD / DMUL on one SM
// as a function of independent chains per thread and warps per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu && ./fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int NCH, int OP>
__global__ void k(double* sink, long long* cyc, int iters, double a, double b)
{
    double v[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) v[c] = threadIdx.x + c;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (OP == 0) v[c] = fma(v[c], a, b);
                else if (OP == 1) v[c] = __dadd_rn(v[c], b);
                else v[c] = __dmul_rn(v[c], a);
            }
        }
    }
    const long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) r += v[c];
    if (r == 123456789.0) sink[0] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NCH, int OP>
static void run(int warps, double* sink, long long* cyc, int nsm)
{
    const int iters = 2000;
    k<NCH, OP><<<nsm, warps * 32>>>(sink, cyc, iters, 1.0000000001, 1e-12);
    cudaDeviceSynchronize();
    k<NCH, OP><<<nsm, warps * 32>>>(sink, cyc, iters, 1.0000000001, 1e-12);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nsm; ++i) avg += (double)h[i];
    avg /= nsm;
    const double ops_per_warp = (double)iters * 8 * NCH;
    // cycles per warp-instruction slot on one SMSP: (warps / 4 warps per SMSP)
    const double per_smsp_ops = ops_per_warp * warps / 4.0;
    printf("op=%s chains=%d warps/SM=%2d: %.2f cycles per dependent step, %.2f cycles per warp-instr per SMSP (pipe busy %.0f %% at 2 cyc/instr)\n",
           OP == 0 ? "DFMA" : (OP == 1 ? "DADD" : "DMUL"), NCH, warps, avg / (iters * 8.0), avg / per_smsp_ops,
           200.0 * per_smsp_ops / avg);
}

int main()
{
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double* sink; long long* cyc;
    cudaMalloc(&sink, 64); cudaMalloc(&cyc, sizeof(long long) * 256);
    const int ws[] = {4, 8, 12, 16, 20, 32};
    for (int w : ws) { run<1, 0>(w, sink, cyc, nsm); run<2, 0>(w, sink, cyc, nsm); run<4, 0>(w, sink, cyc, nsm); run<8, 0>(w, sink, cyc, nsm); }
    run<1, 1>(4, sink, cyc, nsm); run<4, 1>(20, sink, cyc, nsm);
    run<1, 2>(4, sink, cyc, nsm); run<4, 2>(20, sink, cyc, nsm);
    return 0;
}

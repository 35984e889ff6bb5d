// Host build of the per-thread eigenvalue routine (irrotational-warp-lab_b200/csrc/small_eig.cuh) so that the
// CPU test suite can drive the very code the CUDA kernel runs.  Built on demand by tests/test_small_eig_cpu.py.
#include "../../irrotational-warp-lab_b200/csrc/small_eig.cuh"

template <int N>
static void run(const double* mats, int count, double* wr, double* wi, int* ok)
{
    for (int m = 0; m < count; ++m) {
        double a[N][N], r[N], im[N];
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) a[i][j] = mats[(size_t)m * N * N + i * N + j];
        ok[m] = iwb::eig_general<N>(a, r, im) ? 1 : 0;
        for (int i = 0; i < N; ++i) { wr[(size_t)m * N + i] = r[i]; wi[(size_t)m * N + i] = im[i]; }
    }
}

extern "C" void small_eig3(const double* mats, int count, double* wr, double* wi, int* ok) { run<3>(mats, count, wr, wi, ok); }
extern "C" void small_eig4(const double* mats, int count, double* wr, double* wi, int* ok) { run<4>(mats, count, wr, wi, ok); }

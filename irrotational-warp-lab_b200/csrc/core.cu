// Handle lifetime, error reporting, host-side scalar preparation, FMA-peak microbenchmark.
// Interface replaced: _resolve_backend (/root/reference/scripts/reproduce_rodal_exact.py:16-30).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "host_common.h"

namespace iwb {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

// numpy.gradient's scalars for a uniform axis (numpy/lib/_function_base_impl.py):
// 2.*h, and a,b,c of both one-sided second-order formulas, each one IEEE operation.
AxisCoef make_axis_coef(double h)
{
    AxisCoef c;
    c.h2 = 2. * h;
    c.rh2 = 1.0 / c.h2;
    c.a0 = -1.5 / h; c.b0 = 2. / h; c.c0 = -0.5 / h;
    c.a1 = 0.5 / h;  c.b1 = -2. / h; c.c1 = 1.5 / h;
    return c;
}

// Largest s with RN(sqrt(s)) <= R: the reference's inclusive mask r_grid <= R
// (reproduce_rodal_exact.py:221) applied to s = (x*x + y*y) + z*z before the square root.
// RN(sqrt) is monotone, so the two tests select exactly the same points.
double shell_threshold_f64(double R)
{
    if (!(R >= 0.0)) return -1.0;
    if (isinf(R)) return INFINITY;
    double s = R * R;
    while (!isinf(s) && sqrt(nextafter(s, INFINITY)) <= R) s = nextafter(s, INFINITY);
    while (s > 0.0 && sqrt(s) > R) s = nextafter(s, -INFINITY);
    return s;
}

// float32 grid: r and the comparison are float32 (r_lim is a weak Python scalar -> float32)
double shell_threshold_f32(double R)
{
    float Rf = (float)R;
    if (!(Rf >= 0.0f)) return -1.0;
    if (isinf(Rf)) return INFINITY;
    float s = Rf * Rf;
    while (!isinf(s) && sqrtf(nextafterf(s, INFINITY)) <= Rf) s = nextafterf(s, INFINITY);
    while (s > 0.0f && sqrtf(s) > Rf) s = nextafterf(s, -INFINITY);
    return (double)s;
}

// Smallest s >= 0 with RN(sqrt(s)) >= e: the profile's lower-edge test r >= edges[q] (tail.py:75) applied to
// s = (x*x + y*y) + z*z before the square root (fused radial profile, k_energy3d_prof).
double profile_threshold_f64(double e)
{
    if (!(e > 0.0)) return 0.0;
    double s = e * e;
    if (isinf(s)) {
        s = 1.7976931348623157e308;
        if (sqrt(s) < e) return INFINITY;
    }
    while (s > 0.0 && sqrt(nextafter(s, -INFINITY)) >= e) s = nextafter(s, -INFINITY);
    while (sqrt(s) < e) s = nextafter(s, INFINITY);
    return s;
}

// float32 grid: s and r are float32, the comparison with the (float64) edge is made in float64 (radial_profile.cuh)
double profile_threshold_f32(double e)
{
    if (!(e > 0.0)) return 0.0;
    float s = (float)(e * e);
    if (isinf(s)) {
        s = 3.4028234663852886e38f;
        if ((double)sqrtf(s) < e) return INFINITY;
    }
    while (s > 0.0f && (double)sqrtf(nextafterf(s, -INFINITY)) >= e) s = nextafterf(s, -INFINITY);
    while ((double)sqrtf(s) < e) s = nextafterf(s, INFINITY);
    return (double)s;
}

// true when |v| is +0 or 2^lo <= |v| <= 2^hi
static bool zero_or_in_range(double v, int lo, int hi)
{
    if (v == 0.0) return !signbit(v);
    if (!isfinite(v)) return false;
    int e;
    frexp(v, &e);                       // |v| = m * 2^e, m in [0.5, 1)
    return e - 1 >= lo && e - 1 <= hi;
}

// Operand-range proof obligations of phi_exact_f64<.., PRECHECKED = true> (iwb_device.cuh).  `zero_index[a]`
// receives the index of the (single) exactly-zero coordinate of each axis, -1 if none.
bool phi_fast_path_is_safe(const double* const coords[3], const int n[3], double rho, double sigma, double v, double eps,
                           double sinh_rs, double cosh_rs, int zero_index[3], bool float32_grid)
{
    bool ok = true;
    const int lim = float32_grid ? 49 : 240;       // |c| in [2^-lim, 2^lim]
    for (int a = 0; a < 3; ++a) {
        zero_index[a] = -1;
        for (int i = 0; i < n[a]; ++i) {
            const double c = coords[a][i];
            if (!zero_or_in_range(c, -lim, lim)) ok = false;
            if (c == 0.0) {
                if (zero_index[a] >= 0) ok = false;      // two zeros on one axis: not a linspace, stay general
                zero_index[a] = i;
            }
        }
    }
    const double K = (2.0 * sigma) * sinh_rs;
    if (!isfinite(K) || K == 0.0 || !zero_or_in_range(fabs(K), -150, 150)) ok = false;
    // smallest non-zero r of the grid >= smallest non-zero |coordinate|
    double cmin = INFINITY;
    for (int a = 0; a < 3; ++a)
        for (int i = 0; i < n[a]; ++i)
            if (coords[a][i] != 0.0 && fabs(coords[a][i]) < cmin) cmin = fabs(coords[a][i]);
    // the reference zeroes g where |t1| <= 1e-12, t1 = (2 r sigma) sinh(rho sigma) (potential.py:73-77); the branch-free
    // flavour has no such test, so every non-zero r must be clear of the threshold (tiny sigma: K ~ 1e-13 and below)
    if (isfinite(cmin) && !(fabs(K) * cmin > 1.0000001e-12)) ok = false;
    // reciprocal seed of x / (r + eps) from 1/r (div_seeded): (eps / r)^2 < 2^-56
    if (isfinite(cmin) && !(eps <= cmin * 0x1p-28)) ok = false;
    if (!isfinite(cosh_rs) || !isfinite(rho) || !isfinite(sigma) || !isfinite(v)) ok = false;
    if (!(eps >= 0.0) || !(eps <= 1.0)) ok = false;
    if (!(sigma > 0.0) || !(rho >= 0.0)) ok = false;          // |sigma (r + rho)| >= |sigma (r - rho)|: one wall test per point
    if (float32_grid && !((float)sigma > 0.0f)) ok = false;
    if (float32_grid && !(fabs(rho) <= 0x1p40 && fabs(sigma) <= 0x1p40 && fabs(v) <= 0x1p40 && (eps == 0.0 || eps >= 0x1p-100))) ok = false;
    if (!(fabs(rho) <= 0x1p200) || !(fabs(sigma) <= 0x1p200) || !(fabs(v) <= 0x1p200)) ok = false;
    return ok;
}

PhiConst make_phi_const(double rho, double sigma, double v, double eps, double sinh_rs, double cosh_rs)
{
    PhiConst c;
    c.prechecked = 0;
    c.sigma = sigma; c.sigma2 = 2.0 * sigma; c.rho = rho; c.v = v; c.eps = eps;
    c.S = sinh_rs; c.C = cosh_rs;
    c.rK = (double)(1.0L / ((long double)c.sigma2 * (long double)sinh_rs));
    c.sigf = (float)sigma; c.sig2f = 2.0f * (float)sigma; c.rhof = (float)rho; c.vf = (float)v; c.epsf = (float)eps;
    return c;
}

// ---- register-resident FMA throughput (roofline denominator) -------------------------------------
// One CTA of 16 warps per SM, 8 independent chains per thread: 4 warps x 8 chains per scheduler against a dependent-issue
// latency of 8 cycles (DFMA) / 4 cycles (FFMA), so the pipe never waits for an operand.  Thread 0 of every CTA reports
// the clock64 cycles of its loop: cycles / event time is the SM clock the benchmark actually ran at, and
// FMAs / cycles the pipe occupancy (1.000 = one warp instruction every 2 cycles per scheduler for FP64).
// (The round-1 version -- 8 CTAs of 256 threads per SM, 4.7 ms -- reached 92 % of the pipe under ncu,
// profiles/r2_fma_peak_ncu_summary.txt; scripts/probes/fp64_probe.cu is the latency / throughput study.)
template <typename T>
__global__ void __launch_bounds__(512) k_fma_peak(T* sink, long long* cycles, int iters, T a, T b)
{
    T v0 = (T)threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
            v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
        }
    }
    const long long t1 = clock64();
    T r = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
    if (r == (T)123456789) sink[0] = r;   // never true; keeps the chain alive
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---- constant-divisor division self-test -----------------------------------------------------------
__global__ void k_selftest_constdiv(double d, double rd, uint64_t count, uint64_t seed, unsigned long long* mism)
{
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (uint64_t n = tid; n < count; n += stride) {
        // splitmix64 -> random sign/exponent in [2^-40, 2^40) / mantissa
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (n + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        uint64_t mant = z & 0x000FFFFFFFFFFFFFull;
        uint64_t expo = 1023 - 40 + ((z >> 52) % 80);
        uint64_t sign = z >> 63;
        double a = __longlong_as_double((long long)((sign << 63) | (expo << 52) | mant));
        if (cdiv(a, d, rd) != a / d) ++bad;
    }
    if (bad) atomicAdd(mism, bad);
}

// ---- branch-free division / sqrt self-test ---------------------------------------------------------
__global__ void k_selftest_divsqrt(uint64_t count, uint64_t seed, int wide, unsigned long long* out)
{
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long bad_div = 0, bad_sqrt = 0, skipped = 0;
    for (uint64_t n = tid; n < count; n += stride) {
        double v[2];
        for (int k = 0; k < 2; ++k) {
            uint64_t z = seed + 0x9E3779B97F4A7C15ull * (2 * n + k + 1);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            uint64_t mant = z & 0x000FFFFFFFFFFFFFull;
            uint64_t span = wide ? 2046 : 120;
            uint64_t expo = (wide ? 1 : 1023 - 60) + ((z >> 52) % span);
            uint64_t sign = z >> 63;
            v[k] = __longlong_as_double((long long)((sign << 63) | (expo << 52) | mant));
        }
        if (exponent_mid_range(v[0]) && exponent_mid_range(v[1])) {
            if (div_fast(v[0], v[1]) != v[0] / v[1]) ++bad_div;
        } else ++skipped;
        const double s = fabs(v[0]);
        if (exponent_mid_range(s)) {
            if (sqrt_fast(s) != sqrt(s)) ++bad_sqrt;
        }
        // float32 flavours on operands derived from the same bits (exponents folded into the guarded range)
        {
            const unsigned ua = (unsigned)(__double_as_longlong(v[0]) >> 9), ub = (unsigned)(__double_as_longlong(v[1]) >> 11);
            const float fa = __int_as_float((ua & 0x807fffffu) | ((77u + (ua >> 23) % 100u) << 23));
            const float fb = __int_as_float((ub & 0x807fffffu) | ((77u + (ub >> 23) % 100u) << 23));
            if (divf_fast(fa, fb) != __fdiv_rn(fa, fb)) ++bad_div;
            if (sqrtf_fast(fabsf(fa)) != __fsqrt_rn(fabsf(fa))) ++bad_sqrt;
        }
    }
    if (bad_div) atomicAdd(out, bad_div);
    if (bad_sqrt) atomicAdd(out + 1, bad_sqrt);
    if (skipped) atomicAdd(out + 2, skipped);
}

}  // namespace iwb

using namespace iwb;

extern "C" {

int iwb_version(void) { return IWB_VERSION; }

const char* iwb_last_error(void) { return g_last_error.c_str(); }

int iwb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int iwb_create(int device_id, iwb_handle** out)
{
    if (!out) return fail(IWB_ERR_INVALID, "iwb_create: out is NULL");
    *out = nullptr;
    int n = iwb_device_count();
    if (n <= 0)
        return fail(IWB_ERR_NO_DEVICE, "b200 backend requested but no CUDA device is visible "
                                       "(there is no CPU fallback; use backend='numpy' for the reference path)");
    if (device_id < 0 || device_id >= n) return fail(IWB_ERR_INVALID, "iwb_create: device_id out of range");
    iwb_handle* h = new iwb_handle();
    h->device = device_id;
    IWB_CUDA(cudaSetDevice(device_id));
    IWB_CUDA(cudaGetDeviceProperties(&h->prop, device_id));
    if (h->prop.major != 10) {
        std::string nm = h->prop.name;
        delete h;
        return fail(IWB_ERR_NO_DEVICE, "b200 backend needs an sm_100 device, found " + nm);
    }
    h->sm_count = h->prop.multiProcessorCount;
    for (int s = 0; s < NSCRATCH; ++s) {
        IWB_CUDA(cudaStreamCreateWithFlags(&h->scr[s].stream, cudaStreamNonBlocking));
        IWB_CUDA(cudaEventCreate(&h->scr[s].ev0));
        IWB_CUDA(cudaEventCreate(&h->scr[s].ev1));
    }
    const char* cfg = getenv("IWB_K3_CONFIG");
    h->k3_config = cfg ? atoi(cfg) : 1;
    const char* sp = getenv("IWB_K3_STATIC_PCT");
    h->k3_static_pct = sp ? atoi(sp) : -1;
    if (h->k3_static_pct < -1 || h->k3_static_pct > 100) h->k3_static_pct = -1;
    const char* pf = getenv("IWB_PROFILE_FUSED");
    h->profile_fused = pf ? atoi(pf) : 1;
    *out = h;
    return IWB_OK;
}

int iwb_destroy(iwb_handle* h)
{
    if (!h) return IWB_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int s = 0; s < NASYNC; ++s) {
        AsyncSlot& a = h->aslot[s];
        a.h_coords.release(); a.coords.release(); a.partial.release(); a.counter.release();
        if (a.done) cudaEventDestroy(a.done);
    }
    for (int s = 0; s < NSCRATCH; ++s) {
        Scratch& c = h->scr[s];
        if (c.stream) cudaStreamSynchronize(c.stream);
        c.coords.release(); c.partial.release(); c.table.release(); c.out.release(); c.rho.release(); c.counter.release();
        c.h_coords.release(); c.h_out.release(); c.prof.release(); c.h_prof.release();
        if (c.ev0) cudaEventDestroy(c.ev0);
        if (c.ev1) cudaEventDestroy(c.ev1);
        if (c.stream) cudaStreamDestroy(c.stream);
    }
    if (h->prep && h->prep_free) h->prep_free(h->prep);
    delete h;
    return IWB_OK;
}

int iwb_device_info(iwb_handle* h, int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len)
{
    if (!h) return fail(IWB_ERR_INVALID, "null handle");
    if (sm_count) *sm_count = h->sm_count;
    if (cc_major) *cc_major = h->prop.major;
    if (cc_minor) *cc_minor = h->prop.minor;
    if (name && name_len > 0) {
        strncpy(name, h->prop.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    return IWB_OK;
}

int iwb_measure_fma_peak(iwb_handle* h, int use_fp32, double* tflops, double* sm_mhz_effective)
{
    if (!h || !tflops) return fail(IWB_ERR_INVALID, "iwb_measure_fma_peak: null argument");
    IWB_CUDA(cudaSetDevice(h->device));
    Scratch& c = h->scr[0];
    IWB_CUDA(c.out.reserve(ROW_BYTES_MIN + 8 * 1024));
    double* sink = (double*)c.out.p;
    long long* cyc = (long long*)((char*)c.out.p + ROW_BYTES_MIN);
    const int grid = h->sm_count < 1024 ? h->sm_count : 1024, block = 512, iters = use_fp32 ? 60000 : 30000;
    double best_ms = 1e30, best_cycles = 0.0;
    long long hc[1024];
    for (int rep = 0; rep < 5; ++rep) {
        IWB_CUDA(cudaEventRecord(c.ev0, c.stream));
        if (use_fp32) k_fma_peak<float><<<grid, block, 0, c.stream>>>((float*)sink, cyc, iters, 1.0000001f, 1e-9f);
        else k_fma_peak<double><<<grid, block, 0, c.stream>>>(sink, cyc, iters, 1.0000000001, 1e-12);
        IWB_CUDA(cudaEventRecord(c.ev1, c.stream));
        IWB_CUDA(cudaMemcpyAsync(hc, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost, c.stream));
        IWB_CUDA(cudaStreamSynchronize(c.stream));
        IWB_CUDA(cudaGetLastError());
        float ms = 0;
        IWB_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
        if (rep > 0 && ms < best_ms) {
            best_ms = ms;
            double mx = 0.0;
            for (int q = 0; q < grid; ++q) mx = hc[q] > mx ? (double)hc[q] : mx;
            best_cycles = mx;
        }
    }
    const double fmas = (double)grid * block * (double)iters * 64.0;
    *tflops = 2.0 * fmas / (best_ms * 1e-3) / 1e12;
    // the clock the SMs ran at: loop cycles of the slowest CTA over the event time (which adds the launch, ~0.1 %)
    if (sm_mhz_effective) *sm_mhz_effective = best_cycles / (best_ms * 1e-3) / 1e6;
    return IWB_OK;
}

int iwb_selftest_constdiv(iwb_handle* h, double divisor, uint64_t count, uint64_t seed, uint64_t* mismatches)
{
    if (!h || !mismatches) return fail(IWB_ERR_INVALID, "iwb_selftest_constdiv: null argument");
    IWB_CUDA(cudaSetDevice(h->device));
    Scratch& c = h->scr[0];
    IWB_CUDA(c.out.reserve(ROW_BYTES_MIN));
    IWB_CUDA(cudaMemsetAsync(c.out.p, 0, 8, c.stream));
    k_selftest_constdiv<<<h->sm_count * 8, 256, 0, c.stream>>>(divisor, 1.0 / divisor, count, seed,
                                                              (unsigned long long*)c.out.p);
    IWB_CUDA(cudaGetLastError());
    unsigned long long bad = 0;
    IWB_CUDA(cudaMemcpyAsync(&bad, c.out.p, 8, cudaMemcpyDeviceToHost, c.stream));
    IWB_CUDA(cudaStreamSynchronize(c.stream));
    *mismatches = bad;
    return IWB_OK;
}

int iwb_selftest_divsqrt(iwb_handle* h, uint64_t count, uint64_t seed, int wide_exponents, uint64_t* mismatches_div,
                         uint64_t* mismatches_sqrt, uint64_t* skipped)
{
    if (!h || !mismatches_div || !mismatches_sqrt) return fail(IWB_ERR_INVALID, "iwb_selftest_divsqrt: null argument");
    IWB_CUDA(cudaSetDevice(h->device));
    Scratch& c = h->scr[0];
    IWB_CUDA(c.out.reserve(ROW_BYTES_MIN));
    IWB_CUDA(cudaMemsetAsync(c.out.p, 0, 24, c.stream));
    k_selftest_divsqrt<<<h->sm_count * 8, 256, 0, c.stream>>>(count, seed, wide_exponents, (unsigned long long*)c.out.p);
    IWB_CUDA(cudaGetLastError());
    unsigned long long res[3] = {0, 0, 0};
    IWB_CUDA(cudaMemcpyAsync(res, c.out.p, 24, cudaMemcpyDeviceToHost, c.stream));
    IWB_CUDA(cudaStreamSynchronize(c.stream));
    *mismatches_div = res[0]; *mismatches_sqrt = res[1];
    if (skipped) *skipped = res[2];
    return IWB_OK;
}

}  // extern "C"

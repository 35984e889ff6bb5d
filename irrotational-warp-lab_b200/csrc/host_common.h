// Host-side plumbing shared by the C-ABI translation units (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/iwb200.h"
#include "iwb_device.cuh"

namespace iwb {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define IWB_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t err__ = (call);                                                            \
        if (err__ != cudaSuccess)                                                              \
            return ::iwb::fail(IWB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__)); \
    } while (0)

// grow-only device / pinned-host buffers
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// One in-flight evaluation's worth of scratch (a batch uses several, round-robin).
struct Scratch {
    cudaStream_t stream = nullptr;
    DevBuf coords;     // x, xx, yy, zz
    DevBuf partial;    // [nfb][ntiles][ROW]
    DevBuf table;      // [nfb][ROW]
    DevBuf out;        // [ROW]
    DevBuf rho;        // staging for a host rho_out
    DevBuf counter;    // dynamic work-item counter of k_energy3d
    DevBuf prof;       // radial profile: edges | per-(flush block, CTA) partials | result
    PinBuf h_prof;
    PinBuf h_coords;
    PinBuf h_out;      // [ROW]
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

// The reference multiplies by the masks (e * (e > 0), reproduce_rodal_exact.py:57-61, 220-224), so a non-finite
// point poisons sums it does not belong to: NaN * 0 and inf * 0 are NaN.  The kernels select instead of multiplying
// (no extra work per point); this restores the reference's result for a field with non-finite points.  A shell sum
// that is itself infinite keeps its value (the infinite point lies inside that shell).
inline void apply_reference_nonfinite(iwb_result* r, int nshell)
{
    const double qnan = __builtin_nan("");
    const bool any_nan = r->cnt_nan > 0;
    const bool pinf = __builtin_isinf(r->e_pos), ninf = __builtin_isinf(r->e_neg);
    if (!any_nan && !pinf && !ninf) return;
    if (any_nan || ninf) r->e_pos = qnan;
    if (any_nan || pinf) r->e_neg = qnan;
    for (int s = 0; s < nshell; ++s) {
        if (any_nan || ninf || (pinf && !__builtin_isinf(r->shell_pos[s]))) r->shell_pos[s] = qnan;
        if (any_nan || pinf || (ninf && !__builtin_isinf(r->shell_neg[s]))) r->shell_neg[s] = qnan;
    }
}

// Staging of one non-blocking launch (iwb_energy3d_slab_async / _tiles_async / _tiles_exchange_async): the pinned
// coordinate buffer, its device copy, the per-tile partial rows and the work counter.  The handle keeps a small ring of
// them; a slot is reused only after the event recorded behind its previous launch has completed, so calls on different
// streams never share staging memory and no call has to synchronise its stream.
struct AsyncSlot {
    PinBuf h_coords;
    DevBuf coords, partial, counter;
    cudaEvent_t done = nullptr;
    bool in_flight = false;
};
constexpr int NASYNC = 4;

constexpr int NSCRATCH = 4;
constexpr size_t ROW_BYTES_MIN = IWB_ROW_WIDTH * sizeof(double);

}  // namespace iwb

struct iwb_handle {
    int device = -1;
    int sm_count = 0;
    cudaDeviceProp prop;
    iwb::Scratch scr[iwb::NSCRATCH];
    iwb::AsyncSlot aslot[iwb::NASYNC];
    int next_aslot = 0;
    void* prep = nullptr;      // iwb::PrepCache (energy3d.cu): the last few host preparations, see prepare_host
    void (*prep_free)(void*) = nullptr;
    int k3_config = 0;         // kernel geometry variant (IWB_K3_CONFIG env override)
    int k3_static_pct = -1;    // share of the work units handed out as one large span per CTA; -1 = automatic (80 / 70 %),
                               // IWB_K3_STATIC_PCT env override
    int last_profile_path = 0; // iwb_debug_profile_path
    int profile_fused = 1;     // iwb_radial_profile3d bins inside the fused kernel: 1 = for large volumes (where it is the
                               // faster route), 2 = whenever the kernel can, 0 = never (emit rho, bin it in a second
                               // pass); IWB_PROFILE_FUSED env override
};

namespace iwb {
AxisCoef make_axis_coef(double h);
double shell_threshold_f64(double R);
double shell_threshold_f32(double R);
double profile_threshold_f64(double e);
double profile_threshold_f32(double e);
PhiConst make_phi_const(double rho, double sigma, double v, double eps, double sinh_rs, double cosh_rs);
bool phi_fast_path_is_safe(const double* const coords[3], const int n[3], double rho, double sigma, double v, double eps,
                           double sinh_rs, double cosh_rs, int zero_index[3], bool float32_grid);
}  // namespace iwb

// C-ABI entry points of the fused 3D evaluator.  Compiled with -fmad=false (IEEE replay).
// Interfaces replaced: compute_energy_3d + the _e_at_r closure + _split_pos_neg
// (/root/reference/scripts/reproduce_rodal_exact.py:157-235, 57-61) and the loops of
// sweep_velocity (/root/reference/scripts/sweep_superluminal.py:59-106).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "energy3d_launch.cuh"
#include "radial_profile.cuh"
#include "host_common.h"

namespace iwb {

static cudaError_t launch_k3(int dtype, int mode, int nshell, bool emit, int cfg, const K3Params& P, int sm_count, cudaStream_t st)
{
    if (mode == IWB_MODE_FAST) return launch_k3_fast(nshell, emit, cfg, P, sm_count, st);
    if (dtype == IWB_F64)
        return emit ? launch_nsh<IWB_F64, true, 0>(nshell, cfg, P, sm_count, st)
                    : launch_nsh<IWB_F64, false, 0>(nshell, cfg, P, sm_count, st);
    return emit ? launch_nsh<IWB_F32_REF, true, 0>(nshell, cfg, P, sm_count, st)
                : launch_nsh<IWB_F32_REF, false, 0>(nshell, cfg, P, sm_count, st);
}

static int validate(const iwb_params3d* p)
{
    if (!p) return fail(IWB_ERR_INVALID, "params is NULL");
    if (p->dtype != IWB_F64 && p->dtype != IWB_F32_REF)
        return fail(IWB_ERR_INVALID, "energy3d: dtype must be IWB_F64 or IWB_F32_REF");
    if (p->mode != IWB_MODE_EXACT && p->mode != IWB_MODE_FAST) return fail(IWB_ERR_INVALID, "energy3d: unknown mode");
    if (p->mode == IWB_MODE_FAST && p->dtype != IWB_F64)
        return fail(IWB_ERR_INVALID, "energy3d: IWB_MODE_FAST exists for IWB_F64 only");
    if (p->n0 < 3 || p->n1 < 3 || p->n2 < 3)
        return fail(IWB_ERR_GRID_TOO_SMALL, "Shape of array too small to calculate a numerical gradient, "
                                            "at least (edge_order + 1) elements are required.");
    if (!p->x || !p->y || !p->z) return fail(IWB_ERR_INVALID, "energy3d: coordinate arrays are NULL");
    if (p->i_begin < 0 || p->i_end > p->n0 || p->i_begin >= p->i_end)
        return fail(IWB_ERR_INVALID, "energy3d: need 0 <= i_begin < i_end <= n0");
    if (p->i_begin % IWB_FLUSH_PLANES != 0)
        return fail(IWB_ERR_INVALID, "energy3d: i_begin must be a multiple of IWB_FLUSH_PLANES");
    if (p->nshell < 0 || p->nshell > IWB_MAX_SHELLS) return fail(IWB_ERR_INVALID, "energy3d: nshell out of range");
    if (!(p->dx > 0.0) || !(p->dy > 0.0) || !(p->dz > 0.0)) return fail(IWB_ERR_INVALID, "energy3d: spacings must be > 0");
    {
        // e = rho * dV is what the reference classifies and sums; a cell volume outside this range makes that product
        // under- or overflow wholesale (an unusable grid), and the call refuses it instead of returning zeros / infinities
        const double dV = (p->dx * p->dy) * p->dz;
        if (!(dV >= 0x1p-400 && dV <= 0x1p400)) return fail(IWB_ERR_INVALID, "energy3d: dx*dy*dz outside [2^-400, 2^400]");
    }
    const int ts = p->tile_stride;
    if (!(ts == 0 || ts == 1 || ts == 2 || ts == 4 || ts == 8 || ts == 16))
        return fail(IWB_ERR_INVALID, "energy3d: tile_stride must be 0, 1, 2, 4, 8 or 16");
    if (p->tile_first < 0 || p->tile_first >= (ts > 1 ? ts : 1)) return fail(IWB_ERR_INVALID, "energy3d: need 0 <= tile_first < tile_stride");
    if (ts > 1 && p->rho_out) return fail(IWB_ERR_INVALID, "energy3d: rho_out is not available for an interleaved tile shard");
    return IWB_OK;
}

// Everything a launch needs besides the output table.
struct Launch3 {
    K3Params P;
    int dtype, mode, nshell, cfg_id, ntiles, fb0, fb1;
    int tile_stride, tile_first;       // tile_stride > 1: interleaved tile shard (chain sums instead of table rows)
    bool emit;
};

// One cached host preparation (see prepare_host) and the handle's small LRU set of them.
struct PrepEntry {
    bool valid = false, emit = false;
    int static_pct = -1, cfg = 0;
    unsigned long long stamp = 0;
    iwb_params3d key;                  // scalars only (pointers nulled)
    std::vector<double> coords;        // x | y | z as given
    std::vector<double> stage;         // x | x*x | y*y | z*z | span table
    size_t stage_bytes = 0, ncoord = 0;
    Launch3 L;                         // everything but the device pointers
};
struct PrepCache {
    static constexpr int N = 6;
    PrepEntry e[N];
    unsigned long long clock = 0;
};

// Span boundaries over `units` work units for `ctas` persistent CTAs: `static_pct` % of the units in one equal span
// per CTA (taken first, in index order), the rest in spans of geometrically decreasing size (guided self-scheduling,
// half of what an even split of the remainder would give, down to one unit) that level out the cost differences
// between tiles.  Every span boundary is a prologue of the kernel, so few and large spans are cheap; the last spans
// bound the scheduling tail, so they are single units.
static std::vector<int> make_spans(long long units, int ctas, int static_pct)
{
    std::vector<int> b;
    b.push_back(0);
    if (ctas < 1) ctas = 1;
    long long u = 0;
    if (units >= 4LL * ctas && static_pct > 0) {
        const long long us = units * static_pct / 100;
        for (int s = 1; s <= ctas; ++s) {
            const long long e = us * s / ctas;
            if (e > u) { b.push_back((int)e); u = e; }
        }
    }
    while (u < units) {
        long long c = (units - u + 2LL * ctas - 1) / (2LL * ctas);
        if (c < 1) c = 1;
        u += c;
        if (u > units) u = units;
        b.push_back((int)u);
    }
    return b;
}

// Host preparation of one launch: the staged block (x, x*x, y*y, z*z, span table) and the launch record with every
// derived constant (one-sided coefficients, shell thresholds on s, the operand-range proof of the branch-free phi, ...).
// It depends on nothing but the parameters, so the handle keeps the last few (PrepCache): a driver that evaluates the same
// grid again -- a benchmark loop, the _e_at_r closure, the ranks of a sweep -- pays a 30 KB compare instead.  The
// coordinates are still copied to the device on every call.
static int prepare_host(iwb_handle* h, const iwb_params3d* p, bool emit, PrepEntry& E)
{
    const int n0 = p->n0, n1 = p->n1, n2 = p->n2;
    const int cfg_id = config_is_built(h->k3_config, p->dtype, p->nshell, emit) ? h->k3_config : 0;
    const K3Config cfg = kConfigs[cfg_id];

    // work spans of the fused kernel (see k_energy3d), staged behind the coordinates
    const int ntj_ = tiles_j(n1, cfg.rj), ntk_ = tiles_k(n2);
    const int tstride = p->tile_stride > 1 ? p->tile_stride : 1, tfirst = p->tile_stride > 1 ? p->tile_first : 0;
    const int ntiles_local = (ntj_ * ntk_ - tfirst + tstride - 1) / tstride;      // may be 0 on a tiny grid
    // share of the units in the large spans: 80 % when the launch covers the whole of axis 0 (every CTA's span then
    // mixes wall / shell / plain planes), 70 % for a slab (the slabs around the bubble wall are markedly heavier per unit)
    // (an interleaved shard of 8 or more: 60 % -- with a larger share one shard in eight draws a static span full of wall
    // tiles and finishes 3 % late, and a multi-GPU step waits for the slowest rank; profiles/r2_shard_static_sweep.log)
    const int static_pct = h->k3_static_pct >= 0 ? h->k3_static_pct
                           : ((p->i_begin == 0 && p->i_end == n0) ? (tstride >= 8 ? 60 : 80) : 70);
    const std::vector<int> spans = make_spans((long long)(ntiles_local > 0 ? ntiles_local : 0) *
                                                  ((p->i_end - p->i_begin + FLUSH - 1) / FLUSH),
                                              h->sm_count, static_pct);
    // coordinates: x, x*x, y*y, z*z (squares rounded once in the coordinate dtype, as x*x is in NumPy)
    const size_t ncoord = (size_t)2 * n0 + n1 + n2;
    E.ncoord = ncoord;
    E.stage.assign((ncoord * sizeof(double) + spans.size() * sizeof(int) + 7) / 8, 0.0);
    double* hc = E.stage.data();
    double *hx = hc, *hxx = hc + n0, *hyy = hc + 2 * n0, *hzz = hc + 2 * n0 + n1;
    if (p->dtype == IWB_F64) {
        for (int i = 0; i < n0; ++i) { hx[i] = p->x[i]; hxx[i] = p->x[i] * p->x[i]; }
        for (int j = 0; j < n1; ++j) hyy[j] = p->y[j] * p->y[j];
        for (int k = 0; k < n2; ++k) hzz[k] = p->z[k] * p->z[k];
    } else {
        for (int i = 0; i < n0; ++i) { float f = (float)p->x[i]; hx[i] = f; hxx[i] = (double)(f * f); }
        for (int j = 0; j < n1; ++j) { float f = (float)p->y[j]; hyy[j] = (double)(f * f); }
        for (int k = 0; k < n2; ++k) { float f = (float)p->z[k]; hzz[k] = (double)(f * f); }
    }
    memcpy(hc + ncoord, spans.data(), spans.size() * sizeof(int));
    E.stage_bytes = ncoord * sizeof(double) + spans.size() * sizeof(int);

    Launch3* L = &E.L;
    K3Params& P = L->P;
    memset(&P, 0, sizeof(P));
    P.n0 = n0; P.n1 = n1; P.n2 = n2;
    P.i_begin = p->i_begin; P.i_end = p->i_end;
    P.ntj = tiles_j(n1, cfg.rj);
    P.ntk = tiles_k(n2);
    const int ntiles = P.ntj * P.ntk;
    const int planes = p->i_end - p->i_begin;
    const int nfb_slab = (planes + FLUSH - 1) / FLUSH;
    P.nfb_slab = nfb_slab;
    P.nspans = (int)spans.size() - 1;
    P.tile_first = tfirst; P.tile_stride = tstride;
    P.pc = make_phi_const(p->rho, p->sigma, p->v, p->eps, p->sinh_rs, p->cosh_rs);
    {
        const double* const coords[3] = {p->x, p->y, p->z};
        const int nn[3] = {n0, n1, n2};
        int zero_index[3];
        const bool safe = phi_fast_path_is_safe(coords, nn, p->rho, p->sigma, p->v, p->eps, p->sinh_rs, p->cosh_rs, zero_index, p->dtype != IWB_F64);
        P.pc.prechecked = safe ? 1 : 0;
        P.i_zero = zero_index[0];
        P.j_zero = zero_index[1];
    }
    P.cx = make_axis_coef(p->dx); P.cy = make_axis_coef(p->dy); P.cz = make_axis_coef(p->dz);
    P.dV = (p->dx * p->dy) * p->dz;
    P.c16pi = 16.0 * 3.141592653589793;
    P.rc16pi = 1.0 / P.c16pi;
    for (int s = 0; s < IWB_MAX_SHELLS; ++s) {
        if (s < p->nshell)
            P.shell_thr[s] = (p->dtype == IWB_F64) ? shell_threshold_f64(p->shell_r[s]) : shell_threshold_f32(p->shell_r[s]);
        else
            P.shell_thr[s] = -1.0;
    }
    P.shell_thr_max = -1.0;
    for (int s = 0; s < p->nshell; ++s) P.shell_thr_max = fmax(P.shell_thr_max, P.shell_thr[s]);
    L->dtype = p->dtype; L->mode = p->mode; L->nshell = p->nshell; L->cfg_id = cfg_id; L->ntiles = ntiles; L->emit = emit;
    L->fb0 = p->i_begin / FLUSH; L->fb1 = (p->i_end + FLUSH - 1) / FLUSH;
    L->tile_stride = tstride; L->tile_first = tfirst;
    return IWB_OK;
}

// The cached preparation of `p` (computed on a miss).  The key is every value the preparation reads: the scalar
// parameters, and the three coordinate arrays by content.
static int prepared(iwb_handle* h, const iwb_params3d* p, bool emit, const PrepEntry** out)
{
    if (!h->prep) { h->prep = new PrepCache(); h->prep_free = [](void* q) { delete (PrepCache*)q; }; }
    PrepCache& C = *(PrepCache*)h->prep;
    auto same = [&](const PrepEntry& E) {
        const iwb_params3d& q = E.key;
        if (!(E.valid && E.emit == emit && E.static_pct == h->k3_static_pct && E.cfg == h->k3_config)) return false;
        if (!(q.dtype == p->dtype && q.mode == p->mode && q.n0 == p->n0 && q.n1 == p->n1 && q.n2 == p->n2 &&
              q.i_begin == p->i_begin && q.i_end == p->i_end && q.nshell == p->nshell &&
              q.tile_stride == p->tile_stride && q.tile_first == p->tile_first)) return false;
        // (memcmp: the doubles must be the same bits, and NaN parameters must not defeat the cache)
        if (memcmp(&q.dx, &p->dx, 3 * sizeof(double)) || memcmp(&q.rho, &p->rho, 6 * sizeof(double)) ||
            memcmp(q.shell_r, p->shell_r, sizeof(double) * (size_t)(p->nshell > 0 ? p->nshell : 0))) return false;
        const double* k = E.coords.data();
        return !memcmp(k, p->x, sizeof(double) * p->n0) && !memcmp(k + p->n0, p->y, sizeof(double) * p->n1) &&
               !memcmp(k + p->n0 + p->n1, p->z, sizeof(double) * p->n2);
    };
    for (int e = 0; e < PrepCache::N; ++e)
        if (same(C.e[e])) { C.e[e].stamp = ++C.clock; *out = &C.e[e]; return IWB_OK; }
    int victim = 0;
    for (int e = 1; e < PrepCache::N; ++e)
        if (!C.e[e].valid || (C.e[victim].valid && C.e[e].stamp < C.e[victim].stamp)) victim = e;
    PrepEntry& E = C.e[victim];
    E.valid = false;
    int rc = prepare_host(h, p, emit, E);
    if (rc) return rc;
    E.key = *p;
    E.key.x = E.key.y = E.key.z = nullptr; E.key.rho_out = nullptr;
    E.coords.resize((size_t)p->n0 + p->n1 + p->n2);
    memcpy(E.coords.data(), p->x, sizeof(double) * p->n0);
    memcpy(E.coords.data() + p->n0, p->y, sizeof(double) * p->n1);
    memcpy(E.coords.data() + p->n0 + p->n1, p->z, sizeof(double) * p->n2);
    E.emit = emit; E.static_pct = h->k3_static_pct; E.cfg = h->k3_config;
    E.valid = true; E.stamp = ++C.clock;
    *out = &E;
    return IWB_OK;
}

// Stages the prepared block into `h_coords` (pinned) -> `coords` (device) on `st` and completes the launch record with
// the buffers of this launch.
static int prepare_slab(iwb_handle* h, const iwb_params3d* p, PinBuf& h_coords, DevBuf& coords, DevBuf& partial,
                        DevBuf& counter, cudaStream_t st, double* rho_dev, Launch3* L)
{
    const PrepEntry* E = nullptr;
    int rc = prepared(h, p, rho_dev != nullptr, &E);
    if (rc) return rc;
    IWB_CUDA(h_coords.reserve(E->stage_bytes));
    IWB_CUDA(coords.reserve(E->stage_bytes));
    memcpy(h_coords.p, E->stage.data(), E->stage_bytes);
    IWB_CUDA(cudaMemcpyAsync(coords.p, h_coords.p, E->stage_bytes, cudaMemcpyHostToDevice, st));
    *L = E->L;
    K3Params& P = L->P;
    double* dc = (double*)coords.p;
    const int n0 = p->n0, n1 = p->n1;
    P.span_begin = (const int*)(dc + E->ncoord);
    P.x = dc; P.xx = dc + n0; P.yy = dc + 2 * n0; P.zz = dc + 2 * n0 + n1;
    const int nfb_total = (n0 + FLUSH - 1) / FLUSH;
    IWB_CUDA(partial.reserve((size_t)nfb_total * L->ntiles * ROW * sizeof(double)));
    P.partial = (double*)partial.p;
    P.rho_out = rho_dev;
    if (!counter.p) {
        IWB_CUDA(counter.reserve(256));
        IWB_CUDA(cudaMemsetAsync(counter.p, 0, 256, st));
    }
    P.counter = (unsigned int*)counter.p;
    return IWB_OK;
}

// Fused kernel + fixed-order tile reduction into rows [fb0, fb1) of `table`; no host<->device copy.
// `table`: rows [fb0, fb1) of the [nfb][ROW] table, or (tiles_out) the shard's chain sums [nfb][CHAINS / stride][ROW].
static int launch_slab(iwb_handle* h, const Launch3& L, double* table, cudaStream_t st, bool tiles_out = false)
{
    static_assert(IWB_TILE_CHAINS == 16, "k_reduce_tiles is launched with 16 chains");
    if (L.P.nspans > 0) {
        cudaError_t e = launch_k3(L.dtype, L.mode, L.nshell, L.emit, L.cfg_id, L.P, h->sm_count, st);
        if (e != cudaSuccess) return fail(IWB_ERR_CUDA, std::string("k_energy3d launch: ") + cudaGetErrorString(e));
    }
    if (tiles_out)
        k_chain_sums<<<L.fb1 - L.fb0, 32 * (IWB_TILE_CHAINS / L.tile_stride), 0, st>>>(
            L.P.partial, table, L.fb0, L.ntiles, L.tile_first, L.tile_stride, IWB_TILE_CHAINS, L.P.counter);
    else
        k_reduce_tiles<<<L.fb1 - L.fb0, 32 * IWB_TILE_CHAINS, 0, st>>>(L.P.partial, table, L.fb0, L.ntiles, L.P.counter);
    IWB_CUDA(cudaGetLastError());
    return IWB_OK;
}

static int enqueue_slab(iwb_handle* h, Scratch& c, const iwb_params3d* p, double* table, cudaStream_t st,
                        double* rho_dev, bool tiles_out = false)
{
    Launch3 L;
    int rc = prepare_slab(h, p, c.h_coords, c.coords, c.partial, c.counter, st, rho_dev, &L);
    if (rc) return rc;
    return launch_slab(h, L, table, st, tiles_out);
}

// Non-blocking launch with its own staging slot (see AsyncSlot): no stream synchronisation, safe across streams.
static int acquire_aslot(iwb_handle* h, AsyncSlot** out)
{
    AsyncSlot& a = h->aslot[h->next_aslot];
    h->next_aslot = (h->next_aslot + 1) % NASYNC;
    if (!a.done) IWB_CUDA(cudaEventCreateWithFlags(&a.done, cudaEventDisableTiming));
    if (a.in_flight) {
        IWB_CUDA(cudaEventSynchronize(a.done));      // only when NASYNC launches are still in flight
        a.in_flight = false;
    }
    *out = &a;
    return IWB_OK;
}

static int enqueue_slab_async(iwb_handle* h, const iwb_params3d* p, double* table, cudaStream_t st, double* rho_dev,
                              bool tiles_out, Launch3* L_out = nullptr, AsyncSlot** slot_out = nullptr)
{
    AsyncSlot* a = nullptr;
    int rc = acquire_aslot(h, &a);
    if (rc) return rc;
    Launch3 L;
    rc = prepare_slab(h, p, a->h_coords, a->coords, a->partial, a->counter, st, rho_dev, &L);
    if (rc) return rc;
    if (L_out) { *L_out = L; *slot_out = a; return IWB_OK; }       // the caller launches (and records the event) itself
    rc = launch_slab(h, L, table, st, tiles_out);
    if (rc) return rc;
    IWB_CUDA(cudaEventRecord(a->done, st));
    a->in_flight = true;
    return IWB_OK;
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

// enqueue a full evaluation on scratch c; results land in c.h_out after c.ev1
static int enqueue_full(iwb_handle* h, Scratch& c, const iwb_params3d* p)
{
    if (p->tile_stride > 1) return fail(IWB_ERR_INVALID, "an interleaved tile shard (tile_stride > 1) goes through iwb_energy3d_tiles_async");
    const int nfb_total = (p->n0 + FLUSH - 1) / FLUSH;
    IWB_CUDA(c.table.reserve((size_t)nfb_total * ROW * sizeof(double)));
    IWB_CUDA(c.out.reserve(ROW_BYTES_MIN));
    IWB_CUDA(c.h_out.reserve(ROW_BYTES_MIN));
    double* rho_dev = nullptr;
    const size_t rho_bytes = (size_t)(p->i_end - p->i_begin) * p->n1 * p->n2 * sizeof(double);
    if (p->rho_out) {
        if (p->rho_out_is_device) rho_dev = p->rho_out;
        else {
            IWB_CUDA(c.rho.reserve(rho_bytes));
            rho_dev = (double*)c.rho.p;
        }
    }
    IWB_CUDA(cudaEventRecord(c.ev0, c.stream));
    int rc = enqueue_slab(h, c, p, (double*)c.table.p, c.stream, rho_dev);
    if (rc) return rc;
    const int fb0 = p->i_begin / FLUSH, fb1 = (p->i_end + FLUSH - 1) / FLUSH;
    k_reduce_rows<<<1, 128, 0, c.stream>>>((const double*)c.table.p + (size_t)fb0 * ROW, (double*)c.out.p, fb1 - fb0);
    IWB_CUDA(cudaGetLastError());
    IWB_CUDA(cudaMemcpyAsync(c.h_out.p, c.out.p, ROW_BYTES_MIN, cudaMemcpyDeviceToHost, c.stream));
    IWB_CUDA(cudaEventRecord(c.ev1, c.stream));
    if (p->rho_out && !p->rho_out_is_device)
        IWB_CUDA(cudaMemcpyAsync(p->rho_out, rho_dev, rho_bytes, cudaMemcpyDeviceToHost, c.stream));
    return IWB_OK;
}

// ---- phi self-test: the branch-free flavour against the plain operators on random points -------------------------
// (lives here because this translation unit is compiled with -fmad=false like the kernels it checks)
__global__ void k_selftest_phi(PhiConst pc, double lo, double hi, uint64_t count, uint64_t seed, unsigned long long* out)
{
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0, wall = 0;
    for (uint64_t n = tid; n < count; n += stride) {
        double c[3];
        for (int k = 0; k < 3; ++k) {
            uint64_t z = seed + 0x9E3779B97F4A7C15ull * (3 * n + k + 1);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            const double u = (double)(z >> 11) * 0x1p-53;              // [0, 1)
            // log-uniform magnitude in [lo, hi), random sign
            const double mag = lo * exp(u * log(hi / lo));
            c[k] = (z & 1) ? -mag : mag;
        }
        const double s[1] = {(c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]};
        double f[1];
        phi_exact_f64<1, true>(s, c[0], pc, f);
        const double ref = phi_point_f64_ref(s[0], c[0], pc.sigma, pc.sigma2, pc.rho, pc.v, pc.eps, pc.S, pc.C);
        if (__double_as_longlong(f[0]) != __double_as_longlong(ref)) ++bad;
        const double r = sqrt(s[0]);
        if (fabs(pc.sigma * (r - pc.rho)) < LAE_SKIP_F64) ++wall;
    }
    if (bad) atomicAdd(out, bad);
    if (wall) atomicAdd(out + 1, wall);
}

}  // namespace iwb

constexpr size_t PEER_HEADER_BYTES = 512;      // flags [2][PEER_MAX] of 8 bytes = 256, padded
struct iwb_peer {
    iwb_handle* h = nullptr;
    int rank = 0, world = 1, nfb_max = 0;
    size_t buf_doubles = 0;
    void* base[iwb::PEER_MAX] = {};              // flags + two gathered buffers of every rank (own: cudaMalloc, others: IPC-mapped)
    bool opened[iwb::PEER_MAX] = {};
    bool connected = false;
    double* table = nullptr;                     // [nfb_max][ROW] scratch + the ticket word
    unsigned long long epoch = 0;
};

struct iwb_plan3d {
    iwb_handle* h = nullptr;
    iwb::Launch3 L;
    iwb::DevBuf coords, partial, counter;
    iwb::PinBuf h_coords;
};

using namespace iwb;

extern "C" {

int iwb_energy3d(iwb_handle* h, const iwb_params3d* p, iwb_result* out)
{
    if (!h || !out) return fail(IWB_ERR_INVALID, "iwb_energy3d: null argument");
    int rc = validate(p);
    if (rc) return rc;
    auto t0 = std::chrono::steady_clock::now();
    IWB_CUDA(cudaSetDevice(h->device));
    Scratch& c = h->scr[0];
    rc = enqueue_full(h, c, p);
    if (rc) return rc;
    IWB_CUDA(cudaStreamSynchronize(c.stream));
    memset(out, 0, sizeof(*out));
    unpack_row((const double*)c.h_out.p, p->nshell, out);
    out->cnt_nan = (int64_t)(p->i_end - p->i_begin) * p->n1 * p->n2 - out->cnt_pos - out->cnt_neg - out->cnt_zero;
    apply_reference_nonfinite(out, p->nshell);
    float ms = 0;
    IWB_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
    out->kernel_ms = ms;
    out->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return IWB_OK;
}

int iwb_energy3d_slab_async(iwb_handle* h, const iwb_params3d* p, double* table_device, void* stream)
{
    if (!h || !table_device) return fail(IWB_ERR_INVALID, "iwb_energy3d_slab_async: null argument");
    int rc = validate(p);
    if (rc) return rc;
    if (p->rho_out && !p->rho_out_is_device)
        return fail(IWB_ERR_INVALID, "iwb_energy3d_slab_async: rho_out must be a device pointer");
    if (p->tile_stride > 1) return fail(IWB_ERR_INVALID, "iwb_energy3d_slab_async: tile shards go through iwb_energy3d_tiles_async");
    IWB_CUDA(cudaSetDevice(h->device));
    return enqueue_slab_async(h, p, table_device, (cudaStream_t)stream /* NULL is CUDA's default stream */, p->rho_out, false);
}

int iwb_reduce_table(iwb_handle* h, const double* table_device, int nrows, int nshell, void* stream, iwb_result* out)
{
    if (!h || !table_device || !out || nrows < 1) return fail(IWB_ERR_INVALID, "iwb_reduce_table: bad argument");
    IWB_CUDA(cudaSetDevice(h->device));
    Scratch& c = h->scr[0];
    cudaStream_t st = (cudaStream_t)stream;      // NULL is CUDA's default stream
    IWB_CUDA(c.out.reserve(ROW_BYTES_MIN));
    IWB_CUDA(c.h_out.reserve(ROW_BYTES_MIN));
    k_reduce_rows<<<1, 128, 0, st>>>(table_device, (double*)c.out.p, nrows);
    IWB_CUDA(cudaGetLastError());
    IWB_CUDA(cudaMemcpyAsync(c.h_out.p, c.out.p, ROW_BYTES_MIN, cudaMemcpyDeviceToHost, st));
    IWB_CUDA(cudaStreamSynchronize(st));
    memset(out, 0, sizeof(*out));
    unpack_row((const double*)c.h_out.p, nshell, out);
    return IWB_OK;
}

int iwb_reduce_table_async(iwb_handle* h, const double* table_device, int nrows, void* stream, double* row_device)
{
    if (!h || !table_device || !row_device || nrows < 1) return fail(IWB_ERR_INVALID, "iwb_reduce_table_async: bad argument");
    IWB_CUDA(cudaSetDevice(h->device));
    k_reduce_rows<<<1, 128, 0, (cudaStream_t)stream>>>(table_device, row_device, nrows);
    IWB_CUDA(cudaGetLastError());
    return IWB_OK;
}

int iwb_reduce_chains_async(iwb_handle* h, const double* gathered_device, int tile_stride, int nfb, void* stream,
                            double* table_scratch_device, double* row_device)
{
    if (!h || !gathered_device || !table_scratch_device || !row_device || nfb < 1)
        return fail(IWB_ERR_INVALID, "iwb_reduce_chains_async: bad argument");
    if (!(tile_stride == 1 || tile_stride == 2 || tile_stride == 4 || tile_stride == 8 || tile_stride == 16))
        return fail(IWB_ERR_INVALID, "iwb_reduce_chains_async: tile_stride must be 1, 2, 4, 8 or 16");
    IWB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    k_combine_chains<<<nfb, 32 * IWB_TILE_CHAINS, 0, st>>>(gathered_device, table_scratch_device, nfb, tile_stride);
    k_reduce_rows<<<1, 128, 0, st>>>(table_scratch_device, row_device, nfb);
    IWB_CUDA(cudaGetLastError());
    return IWB_OK;
}

int iwb_selftest_phi(iwb_handle* h, double rho, double sigma, double v, double eps, double coord_lo, double coord_hi,
                     uint64_t count, uint64_t seed, uint64_t* mismatches, uint64_t* wall_points)
{
    if (!h || !mismatches) return fail(IWB_ERR_INVALID, "iwb_selftest_phi: null argument");
    if (!(coord_lo > 0.0) || !(coord_hi > coord_lo)) return fail(IWB_ERR_INVALID, "iwb_selftest_phi: need 0 < coord_lo < coord_hi");
    const double sinh_rs = sinh(rho * sigma), cosh_rs = cosh(rho * sigma);
    PhiConst pc = make_phi_const(rho, sigma, v, eps, sinh_rs, cosh_rs);
    {   // the host's operand-range proof on the two extreme coordinates (every |coordinate| lies between them)
        const double ax[2] = {coord_lo, coord_hi};
        const double* const coords[3] = {ax, ax, ax};
        const int nn[3] = {2, 2, 2};
        int zero_index[3];
        if (!phi_fast_path_is_safe(coords, nn, rho, sigma, v, eps, sinh_rs, cosh_rs, zero_index, false))
            return fail(IWB_ERR_INVALID, "iwb_selftest_phi: these parameters fail the operand-range proof (the engine "
                                         "would use the reference flavour throughout)");
    }
    pc.prechecked = 1;
    IWB_CUDA(cudaSetDevice(h->device));
    Scratch& c = h->scr[0];
    IWB_CUDA(c.out.reserve(ROW_BYTES_MIN));
    IWB_CUDA(cudaMemsetAsync(c.out.p, 0, 16, c.stream));
    k_selftest_phi<<<h->sm_count * 8, 256, 0, c.stream>>>(pc, coord_lo, coord_hi, count, seed, (unsigned long long*)c.out.p);
    IWB_CUDA(cudaGetLastError());
    unsigned long long res[2] = {0, 0};
    IWB_CUDA(cudaMemcpyAsync(res, c.out.p, 16, cudaMemcpyDeviceToHost, c.stream));
    IWB_CUDA(cudaStreamSynchronize(c.stream));
    *mismatches = res[0];
    if (wall_points) *wall_points = res[1];
    return IWB_OK;
}

int iwb_debug_spans(long long units, int ctas, int static_pct, int* out, int cap)
{
    const std::vector<int> b = make_spans(units, ctas, static_pct);
    for (size_t q = 0; q < b.size() && (int)q < cap; ++q)
        if (out) out[q] = b[q];
    return (int)b.size();
}

int iwb_debug_tiles_k(int n, int* out, int cap)
{
    const int nt = tiles_k(n);
    for (int t = 0; t <= nt && t < cap; ++t)
        if (out) out[t] = tile_begin_k(t, nt, n);
    return nt;
}

int iwb_debug_profile_path(iwb_handle* h) { return h ? h->last_profile_path : 0; }

int iwb_debug_tiles_j(int n, int* out, int cap)
{
    const int rj = kConfigs[1].rj;       // the default geometry (region of 40 rows)
    const int nt = tiles_j(n, rj);
    for (int t = 0; t <= nt && t < cap; ++t)
        if (out) out[t] = tile_begin_j(t, nt, n);
    return nt;
}

int iwb_energy3d_tiles_async(iwb_handle* h, const iwb_params3d* p, double* chains_device, void* stream)
{
    if (!h || !chains_device) return fail(IWB_ERR_INVALID, "iwb_energy3d_tiles_async: null argument");
    int rc = validate(p);
    if (rc) return rc;
    if (p->rho_out) return fail(IWB_ERR_INVALID, "iwb_energy3d_tiles_async: rho_out must be NULL");
    IWB_CUDA(cudaSetDevice(h->device));
    return enqueue_slab_async(h, p, chains_device, (cudaStream_t)stream /* NULL is CUDA's default stream */, nullptr, true);
}

int iwb_read_row(iwb_handle* h, const double* row_device, int nshell, void* stream, iwb_result* out)
{
    if (!h || !row_device || !out) return fail(IWB_ERR_INVALID, "iwb_read_row: null argument");
    if (nshell < 0 || nshell > IWB_MAX_SHELLS) return fail(IWB_ERR_INVALID, "iwb_read_row: nshell out of range");
    IWB_CUDA(cudaSetDevice(h->device));
    Scratch& c = h->scr[0];
    cudaStream_t st = (cudaStream_t)stream;
    IWB_CUDA(c.h_out.reserve(ROW_BYTES_MIN));
    IWB_CUDA(cudaMemcpyAsync(c.h_out.p, row_device, ROW_BYTES_MIN, cudaMemcpyDeviceToHost, st));
    IWB_CUDA(cudaStreamSynchronize(st));
    memset(out, 0, sizeof(*out));
    unpack_row((const double*)c.h_out.p, nshell, out);
    return IWB_OK;
}

// ---- peer exchange: the chain sums of the tile-shard path cross NVLink as plain stores (no NCCL launch) ---------------
int iwb_peer_create(iwb_handle* h, int rank, int world, int nfb_max, iwb_peer** out, void* handle_out)
{
    if (!h || !out || !handle_out) return fail(IWB_ERR_INVALID, "iwb_peer_create: null argument");
    *out = nullptr;
    if (!(world == 1 || world == 2 || world == 4 || world == 8 || world == 16) || rank < 0 || rank >= world)
        return fail(IWB_ERR_INVALID, "iwb_peer_create: world must be 1, 2, 4, 8 or 16 and 0 <= rank < world");
    if (nfb_max < 1) return fail(IWB_ERR_INVALID, "iwb_peer_create: nfb_max must be >= 1");
    IWB_CUDA(cudaSetDevice(h->device));
    iwb_peer* pr = new iwb_peer();
    pr->h = h; pr->rank = rank; pr->world = world; pr->nfb_max = nfb_max;
    pr->buf_doubles = (size_t)nfb_max * IWB_TILE_CHAINS * ROW;          // [S][nfb][NCH / S][ROW]
    const size_t bytes = PEER_HEADER_BYTES + 2 * pr->buf_doubles * sizeof(double);
    cudaError_t e = cudaMalloc(&pr->base[rank], bytes);
    if (e == cudaSuccess) e = cudaMemset(pr->base[rank], 0, bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pr->table, ((size_t)nfb_max * ROW + 32) * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(pr->table, 0, ((size_t)nfb_max * ROW + 32) * sizeof(double));
    cudaIpcMemHandle_t hd;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hd, pr->base[rank]);
    if (e != cudaSuccess) {
        if (pr->base[rank]) cudaFree(pr->base[rank]);
        if (pr->table) cudaFree(pr->table);
        delete pr;
        return fail(IWB_ERR_CUDA, std::string("iwb_peer_create: ") + cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == IWB_PEER_HANDLE_BYTES, "IPC handle size");
    memcpy(handle_out, &hd, sizeof(hd));
    *out = pr;
    return IWB_OK;
}

int iwb_peer_connect(iwb_peer* pr, const void* all_handles)
{
    if (!pr || !all_handles) return fail(IWB_ERR_INVALID, "iwb_peer_connect: null argument");
    IWB_CUDA(cudaSetDevice(pr->h->device));
    for (int r = 0; r < pr->world; ++r) {
        if (r == pr->rank) continue;
        cudaIpcMemHandle_t hd;
        memcpy(&hd, (const char*)all_handles + (size_t)r * IWB_PEER_HANDLE_BYTES, sizeof(hd));
        IWB_CUDA(cudaIpcOpenMemHandle(&pr->base[r], hd, cudaIpcMemLazyEnablePeerAccess));
        pr->opened[r] = true;
    }
    pr->connected = true;
    return IWB_OK;
}

int iwb_debug_peer_connect_local(iwb_peer* pr, iwb_peer* const* all_peers)
{
    if (!pr || !all_peers) return fail(IWB_ERR_INVALID, "iwb_debug_peer_connect_local: null argument");
    for (int r = 0; r < pr->world; ++r) {
        if (!all_peers[r] || all_peers[r]->world != pr->world || all_peers[r]->rank != r || all_peers[r]->nfb_max != pr->nfb_max)
            return fail(IWB_ERR_INVALID, "iwb_debug_peer_connect_local: peers do not form one group");
        if (r != pr->rank) pr->base[r] = all_peers[r]->base[r];
    }
    pr->connected = true;
    return IWB_OK;
}

int iwb_peer_destroy(iwb_peer* pr)
{
    if (!pr) return IWB_OK;
    cudaSetDevice(pr->h->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < pr->world; ++r)
        if (pr->opened[r]) cudaIpcCloseMemHandle(pr->base[r]);
    if (pr->base[pr->rank]) cudaFree(pr->base[pr->rank]);
    if (pr->table) cudaFree(pr->table);
    delete pr;
    return IWB_OK;
}

// k_energy3d of launch record L, then the push of the chain sums into every rank's gathered buffer and the combine
static int launch_exchange(iwb_handle* h, iwb_peer* pr, const Launch3& L, int nfb, cudaStream_t st, double* row_device)
{
    const int S = pr->world, first = pr->rank;
    if (L.P.nspans > 0) {
        cudaError_t e = launch_k3(L.dtype, L.mode, L.nshell, L.emit, L.cfg_id, L.P, h->sm_count, st);
        if (e != cudaSuccess) return fail(IWB_ERR_CUDA, std::string("k_energy3d launch: ") + cudaGetErrorString(e));
    }
    const unsigned long long epoch = ++pr->epoch;
    const int par = (int)(epoch & 1ull);
    ExchParams X;
    memset(&X, 0, sizeof(X));
    for (int r = 0; r < S; ++r) {
        char* b = (char*)pr->base[r];
        X.peers.flag[r] = (unsigned long long*)(b + (size_t)par * PEER_MAX * sizeof(unsigned long long));
        X.peers.gath[r] = (double*)(b + PEER_HEADER_BYTES) + (size_t)par * pr->buf_doubles;
    }
    X.gath_local = X.peers.gath[pr->rank];
    X.flag_local = X.peers.flag[pr->rank];
    X.epoch = epoch; X.rank = pr->rank; X.stride = S; X.nfb = nfb;
    X.table = pr->table; X.row = row_device;
    X.ticket = (unsigned int*)(pr->table + (size_t)pr->nfb_max * ROW);
    {
        static const double timeout_s = [] {
            const char* e = getenv("IWB_EXCHANGE_TIMEOUT_S");
            const double v = e ? atof(e) : 60.0;
            return (v > 0.0 && v < 86400.0) ? v : 60.0;
        }();
        X.timeout_ns = (unsigned long long)(timeout_s * 1e9);
    }
    k_chain_sums_push<<<nfb, 32 * (IWB_TILE_CHAINS / S), 0, st>>>(L.P.partial, X.peers, 0, nfb, L.ntiles, first, S,
                                                                  IWB_TILE_CHAINS, L.P.counter);
    k_exchange_combine<<<nfb, 32 * IWB_TILE_CHAINS, 0, st>>>(X);
    IWB_CUDA(cudaGetLastError());
    return IWB_OK;
}

int iwb_energy3d_tiles_exchange_async(iwb_handle* h, iwb_peer* pr, const iwb_params3d* p, void* stream, double* row_device)
{
    if (!h || !pr || !row_device) return fail(IWB_ERR_INVALID, "iwb_energy3d_tiles_exchange_async: null argument");
    int rc = validate(p);
    if (rc) return rc;
    if (!pr->connected) return fail(IWB_ERR_INVALID, "iwb_energy3d_tiles_exchange_async: iwb_peer_connect has not been called");
    if (p->rho_out) return fail(IWB_ERR_INVALID, "iwb_energy3d_tiles_exchange_async: rho_out must be NULL");
    const int S = p->tile_stride > 1 ? p->tile_stride : 1, first = p->tile_stride > 1 ? p->tile_first : 0;
    if (S != pr->world || first != pr->rank)
        return fail(IWB_ERR_INVALID, "iwb_energy3d_tiles_exchange_async: tile_stride / tile_first must be the group's world size / this rank");
    if (p->i_begin != 0 || p->i_end != p->n0)
        return fail(IWB_ERR_INVALID, "iwb_energy3d_tiles_exchange_async: a tile shard covers the whole of axis 0");
    const int nfb = (p->n0 + FLUSH - 1) / FLUSH;
    if (nfb > pr->nfb_max) return fail(IWB_ERR_INVALID, "iwb_energy3d_tiles_exchange_async: grid has more flush blocks than the peer buffers hold");
    IWB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    Launch3 L;
    AsyncSlot* a = nullptr;
    rc = enqueue_slab_async(h, p, nullptr, st, nullptr, true, &L, &a);
    if (rc) return rc;
    rc = launch_exchange(h, pr, L, nfb, st, row_device);
    if (rc) return rc;
    IWB_CUDA(cudaEventRecord(a->done, st));
    a->in_flight = true;
    return IWB_OK;
}

int iwb_reduce_chains(iwb_handle* h, const double* gathered_device, int tile_stride, int nfb, int nshell, void* stream,
                      iwb_result* out)
{
    if (!h || !gathered_device || !out || nfb < 1) return fail(IWB_ERR_INVALID, "iwb_reduce_chains: bad argument");
    if (!(tile_stride == 1 || tile_stride == 2 || tile_stride == 4 || tile_stride == 8 || tile_stride == 16))
        return fail(IWB_ERR_INVALID, "iwb_reduce_chains: tile_stride must be 1, 2, 4, 8 or 16");
    IWB_CUDA(cudaSetDevice(h->device));
    Scratch& c = h->scr[0];
    cudaStream_t st = (cudaStream_t)stream;
    IWB_CUDA(c.table.reserve((size_t)nfb * ROW * sizeof(double)));
    IWB_CUDA(c.out.reserve(ROW_BYTES_MIN));
    IWB_CUDA(c.h_out.reserve(ROW_BYTES_MIN));
    k_combine_chains<<<nfb, 32 * IWB_TILE_CHAINS, 0, st>>>(gathered_device, (double*)c.table.p, nfb, tile_stride);
    k_reduce_rows<<<1, 128, 0, st>>>((const double*)c.table.p, (double*)c.out.p, nfb);
    IWB_CUDA(cudaGetLastError());
    IWB_CUDA(cudaMemcpyAsync(c.h_out.p, c.out.p, ROW_BYTES_MIN, cudaMemcpyDeviceToHost, st));
    IWB_CUDA(cudaStreamSynchronize(st));
    memset(out, 0, sizeof(*out));
    unpack_row((const double*)c.h_out.p, nshell, out);
    return IWB_OK;
}

}  // extern "C"

// Radial profile inside the fused kernel (k_energy3d_prof): one launch over [i_begin, i_end), no rho in HBM.  Every work
// unit (tile x flush block) leaves a window of PROF_WIN bins; k_prof_reduce_windows adds the tiles of a flush block in
// schedule order, k_radial_reduce_fb the flush blocks in index order -- fixed order, independent of scheduling and of how
// the planes are cut into slabs.  *done = false (nothing written) when a unit touched more bins than its window holds:
// the caller then takes the two-pass path.
static int radial_profile_fused(iwb_handle* h, const iwb_params3d* p, const iwb_profile3d* prof, iwb_result* out, bool* done)
{
    *done = false;
    auto t0 = std::chrono::steady_clock::now();
    const int nb = prof->n_bins;
    Scratch& c = h->scr[0];
    cudaStream_t st = c.stream;
    IWB_CUDA(cudaStreamSynchronize(st));
    Launch3 L;
    int rc = prepare_slab(h, p, c.h_coords, c.coords, c.partial, c.counter, st, nullptr, &L);
    if (rc) return rc;
    if (L.cfg_id != 1 || L.P.nspans <= 0) return IWB_OK;           // other geometry / empty grid: two-pass path
    const int nfb_total = (p->n0 + FLUSH - 1) / FLUSH;
    const int fb0 = L.fb0, fb1 = L.fb1, planes = p->i_end - p->i_begin;
    const size_t nunits = (size_t)nfb_total * L.ntiles;
    // device (doubles): thr[nb+1] | wsum[nunits][WIN] | fb_sum[nfb][nb] | fb_cnt[nfb][nb] | sum[nb] | cnt[nb] | wcnt (int) | wbase (int) | overflow
    const size_t off_wsum = (size_t)nb + 1, off_fsum = off_wsum + nunits * PROF_WIN, off_fcnt = off_fsum + (size_t)nfb_total * nb,
                 off_sum = off_fcnt + (size_t)nfb_total * nb, off_cnt = off_sum + nb, off_wcnt = off_cnt + nb,
                 off_wbase = off_wcnt + (nunits * PROF_WIN + 1) / 2, off_ovf = off_wbase + (nunits + 1) / 2;
    IWB_CUDA(c.prof.reserve((off_ovf + 1) * sizeof(double)));
    IWB_CUDA(c.h_prof.reserve(((size_t)3 * nb + 4) * sizeof(double)));
    IWB_CUDA(c.table.reserve((size_t)nfb_total * ROW * sizeof(double)));
    IWB_CUDA(c.out.reserve(ROW_BYTES_MIN));
    IWB_CUDA(c.h_out.reserve(ROW_BYTES_MIN));
    double* dprof = (double*)c.prof.p;
    double* hprof = (double*)c.h_prof.p;
    for (int q = 0; q <= nb; ++q)
        hprof[q] = (p->dtype == IWB_F64) ? profile_threshold_f64(prof->edges[q]) : profile_threshold_f32(prof->edges[q]);
    IWB_CUDA(cudaMemcpyAsync(dprof, hprof, ((size_t)nb + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
    IWB_CUDA(cudaMemsetAsync(dprof + off_ovf, 0, sizeof(double), st));
    K3Params& P = L.P;
    P.prof_thr = dprof; P.prof_nb = nb;
    P.prof_sum = dprof + off_wsum;
    P.prof_cnt = (int*)(dprof + off_wcnt);
    P.prof_base = (int*)(dprof + off_wbase);
    P.prof_overflow = (unsigned int*)(dprof + off_ovf);
    IWB_CUDA(cudaEventRecord(c.ev0, st));
    cudaError_t e = (p->dtype == IWB_F64) ? launch_prof<IWB_F64>(P, h->sm_count, st) : launch_prof<IWB_F32_REF>(P, h->sm_count, st);
    if (e != cudaSuccess) return fail(IWB_ERR_CUDA, std::string("k_energy3d_prof launch: ") + cudaGetErrorString(e));
    k_reduce_tiles<<<fb1 - fb0, 32 * IWB_TILE_CHAINS, 0, st>>>(P.partial, (double*)c.table.p, fb0, L.ntiles, P.counter);
    k_prof_reduce_windows<<<fb1 - fb0, PROF_FUSED_MAX_BINS, (size_t)L.ntiles * sizeof(int), st>>>(
        P.prof_sum, P.prof_cnt, P.prof_base, fb0, L.ntiles, nb, dprof + off_fsum, (long long*)(dprof + off_fcnt));
    k_radial_reduce_fb<<<(nb + 63) / 64, 64, 0, st>>>(dprof + off_fsum, (const long long*)(dprof + off_fcnt), fb0, fb1, nb,
                                                      dprof + off_sum, (long long*)(dprof + off_cnt));
    k_reduce_rows<<<1, 128, 0, st>>>((const double*)c.table.p + (size_t)fb0 * ROW, (double*)c.out.p, fb1 - fb0);
    IWB_CUDA(cudaGetLastError());
    IWB_CUDA(cudaEventRecord(c.ev1, st));
    IWB_CUDA(cudaMemcpyAsync(hprof, dprof + off_sum, (size_t)2 * nb * sizeof(double), cudaMemcpyDeviceToHost, st));
    IWB_CUDA(cudaMemcpyAsync(hprof + 2 * nb, dprof + off_ovf, sizeof(double), cudaMemcpyDeviceToHost, st));
    IWB_CUDA(cudaMemcpyAsync(c.h_out.p, c.out.p, ROW_BYTES_MIN, cudaMemcpyDeviceToHost, st));
    IWB_CUDA(cudaStreamSynchronize(st));
    unsigned int overflow = 0;
    memcpy(&overflow, hprof + 2 * nb, sizeof(overflow));
    if (overflow) return IWB_OK;                                  // bins narrower than a unit's window can hold
    memcpy(prof->sum, hprof, (size_t)nb * sizeof(double));
    memcpy(prof->count, hprof + nb, (size_t)nb * sizeof(int64_t));
    if (out) {
        memset(out, 0, sizeof(*out));
        unpack_row((const double*)c.h_out.p, p->nshell, out);
        out->cnt_nan = (int64_t)planes * p->n1 * p->n2 - out->cnt_pos - out->cnt_neg - out->cnt_zero;
        apply_reference_nonfinite(out, p->nshell);
        float ms = 0;
        IWB_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
        out->kernel_ms = ms;
        out->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    *done = true;
    return IWB_OK;
}

extern "C" {

int iwb_radial_profile3d(iwb_handle* h, const iwb_params3d* p, const iwb_profile3d* prof, iwb_result* out)
{
    if (!h || !prof) return fail(IWB_ERR_INVALID, "iwb_radial_profile3d: null argument");
    int rc = validate(p);
    if (rc) return rc;
    if (p->rho_out) return fail(IWB_ERR_INVALID, "iwb_radial_profile3d: rho_out must be NULL (the field is staged internally)");
    if (p->tile_stride > 1) return fail(IWB_ERR_INVALID, "iwb_radial_profile3d: not available for an interleaved tile shard");
    const int nb = prof->n_bins;
    if (nb < 1 || nb > PROF_MAX_BINS) return fail(IWB_ERR_INVALID, "iwb_radial_profile3d: n_bins must be in [1, 256]");
    if (!prof->edges || !prof->sum || !prof->count) return fail(IWB_ERR_INVALID, "iwb_radial_profile3d: edges / sum / count are NULL");
    for (int q = 0; q <= nb; ++q)
        if (!isfinite(prof->edges[q])) return fail(IWB_ERR_INVALID, "iwb_radial_profile3d: edges must be finite");
    for (int q = 0; q < nb; ++q)
        if (!(prof->edges[q + 1] > prof->edges[q])) return fail(IWB_ERR_INVALID, "iwb_radial_profile3d: edges must be strictly increasing");
    auto t0 = std::chrono::steady_clock::now();
    IWB_CUDA(cudaSetDevice(h->device));
    // Bins inside the fused kernel when that is the faster route: measured on B200 (profiles/r2_f_rows_probe.log) the fused
    // kernel takes 18.2 ms at n = 1024^3 against 19.4 ms for emit-then-bin (and needs no 1 GiB staging buffer), but 3.24
    // against 2.56 ms at n = 512^3, and a work unit of a grid coarser than n ~ 400 spans more bins than a window holds.
    // profile_fused: 0 never, 1 automatic (large volumes), 2 whenever the kernel can (tests).
    const bool large = (double)(p->i_end - p->i_begin) * p->n1 * p->n2 >= 536870912.0;
    if ((h->profile_fused == 2 || (h->profile_fused == 1 && large)) && nb <= PROF_FUSED_MAX_BINS &&
        p->mode == IWB_MODE_EXACT && p->nshell <= 2) {
        bool done = false;
        rc = radial_profile_fused(h, p, prof, out, &done);
        if (!rc && done) h->last_profile_path = 1;
        if (rc || done) return rc;
    }
    h->last_profile_path = 2;
    Scratch& c = h->scr[0];
    cudaStream_t st = c.stream;

    // planes per pass: the emitted field of one pass stays <= 1 GiB (at least one flush block)
    const size_t plane_bytes = (size_t)p->n1 * p->n2 * sizeof(double);
    int pass = (int)(((size_t)1 << 30) / plane_bytes) / FLUSH * FLUSH;
    if (pass < FLUSH) pass = FLUSH;
    const int planes = p->i_end - p->i_begin;
    if (pass > planes) pass = (planes + FLUSH - 1) / FLUSH * FLUSH;
    const int nfb_total = (p->n0 + FLUSH - 1) / FLUSH;
    const int fb0 = p->i_begin / FLUSH, fb1 = (p->i_end + FLUSH - 1) / FLUSH;

    // device: edges[nb+1] | part_sum[nfb][SUB][nb] | part_cnt[...] | sum[nb] | cnt[nb] | fb_sum[nfb][nb] | fb_cnt[nfb][nb]
    const size_t npart = (size_t)nfb_total * PROF_SUB * nb;
    const size_t off_psum = (size_t)nb + 1, off_pcnt = off_psum + npart, off_sum = off_pcnt + npart,
                 off_cnt = off_sum + nb, off_fsum = off_cnt + nb, off_fcnt = off_fsum + (size_t)nfb_total * nb;
    IWB_CUDA(c.prof.reserve((off_fcnt + (size_t)nfb_total * nb) * sizeof(double)));
    IWB_CUDA(c.h_prof.reserve(((size_t)3 * nb + 2) * sizeof(double)));
    IWB_CUDA(c.rho.reserve((size_t)pass * plane_bytes));
    IWB_CUDA(c.table.reserve((size_t)nfb_total * ROW * sizeof(double)));
    IWB_CUDA(c.out.reserve(ROW_BYTES_MIN));
    IWB_CUDA(c.h_out.reserve(ROW_BYTES_MIN));
    double* dprof = (double*)c.prof.p;
    double* hprof = (double*)c.h_prof.p;
    memcpy(hprof, prof->edges, ((size_t)nb + 1) * sizeof(double));
    IWB_CUDA(cudaMemcpyAsync(dprof, hprof, ((size_t)nb + 1) * sizeof(double), cudaMemcpyHostToDevice, st));

    IWB_CUDA(cudaEventRecord(c.ev0, st));
    for (int ia = p->i_begin; ia < p->i_end; ia += pass) {
        iwb_params3d q = *p;
        q.i_begin = ia;
        q.i_end = (ia + pass < p->i_end) ? ia + pass : p->i_end;
        IWB_CUDA(cudaStreamSynchronize(st));       // the coordinate staging buffer and the rho pass buffer are reused
        Launch3 L;
        rc = prepare_slab(h, &q, c.h_coords, c.coords, c.partial, c.counter, st, (double*)c.rho.p, &L);
        if (rc) return rc;
        rc = launch_slab(h, L, (double*)c.table.p, st);
        if (rc) return rc;
        ProfParams R;
        R.rho = (const double*)c.rho.p;
        R.xx = L.P.xx; R.yy = L.P.yy; R.zz = L.P.zz;
        R.edges = dprof;
        R.n1 = p->n1; R.n2 = p->n2; R.i_begin = q.i_begin; R.i_end = q.i_end; R.nb = nb; R.dtype = p->dtype;
        R.inv_width = (double)nb / (prof->edges[nb] - prof->edges[0]);     // finite and > 0 (an overflowing span gives 0)
        R.part_sum = dprof + off_psum;
        R.part_cnt = (long long*)(dprof + off_pcnt);
        k_radial_hist<<<dim3(PROF_SUB, L.fb1 - L.fb0), PROF_NW * 32, 0, st>>>(R);
        IWB_CUDA(cudaGetLastError());
    }
    k_radial_reduce_sub<<<fb1 - fb0, 64, 0, st>>>(dprof + off_psum, (const long long*)(dprof + off_pcnt), fb0, nb,
                                                  dprof + off_fsum, (long long*)(dprof + off_fcnt));
    k_radial_reduce_fb<<<(nb + 63) / 64, 64, 0, st>>>(dprof + off_fsum, (const long long*)(dprof + off_fcnt), fb0, fb1, nb,
                                                      dprof + off_sum, (long long*)(dprof + off_cnt));
    IWB_CUDA(cudaGetLastError());
    k_reduce_rows<<<1, 128, 0, st>>>((const double*)c.table.p + (size_t)fb0 * ROW, (double*)c.out.p, fb1 - fb0);
    IWB_CUDA(cudaGetLastError());
    IWB_CUDA(cudaEventRecord(c.ev1, st));
    IWB_CUDA(cudaMemcpyAsync(hprof, dprof + off_sum, (size_t)2 * nb * sizeof(double), cudaMemcpyDeviceToHost, st));
    IWB_CUDA(cudaMemcpyAsync(c.h_out.p, c.out.p, ROW_BYTES_MIN, cudaMemcpyDeviceToHost, st));
    IWB_CUDA(cudaStreamSynchronize(st));
    memcpy(prof->sum, hprof, (size_t)nb * sizeof(double));
    memcpy(prof->count, hprof + nb, (size_t)nb * sizeof(int64_t));
    if (out) {
        memset(out, 0, sizeof(*out));
        unpack_row((const double*)c.h_out.p, p->nshell, out);
        out->cnt_nan = (int64_t)planes * p->n1 * p->n2 - out->cnt_pos - out->cnt_neg - out->cnt_zero;
        apply_reference_nonfinite(out, p->nshell);
        float ms = 0;
        IWB_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
        out->kernel_ms = ms;
        out->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return IWB_OK;
}

int iwb_plan3d_create(iwb_handle* h, const iwb_params3d* p, iwb_plan3d** out)
{
    if (!h || !out) return fail(IWB_ERR_INVALID, "iwb_plan3d_create: null argument");
    *out = nullptr;
    int rc = validate(p);
    if (rc) return rc;
    if (p->rho_out && !p->rho_out_is_device)
        return fail(IWB_ERR_INVALID, "iwb_plan3d_create: rho_out must be a device pointer (or NULL)");
    IWB_CUDA(cudaSetDevice(h->device));
    iwb_plan3d* plan = new iwb_plan3d();
    plan->h = h;
    cudaStream_t st = h->scr[0].stream;
    rc = prepare_slab(h, p, plan->h_coords, plan->coords, plan->partial, plan->counter, st, p->rho_out, &plan->L);
    if (rc == IWB_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = fail(IWB_ERR_CUDA, "iwb_plan3d_create: upload failed");
    if (rc) {
        plan->coords.release(); plan->partial.release(); plan->counter.release(); plan->h_coords.release();
        delete plan;
        return rc;
    }
    *out = plan;
    return IWB_OK;
}

int iwb_plan3d_run_async(iwb_plan3d* plan, double* table_device, void* stream)
{
    if (!plan || !table_device) return fail(IWB_ERR_INVALID, "iwb_plan3d_run_async: null argument");
    IWB_CUDA(cudaSetDevice(plan->h->device));
    cudaStream_t st = (cudaStream_t)stream;      // NULL is CUDA's default stream
    return launch_slab(plan->h, plan->L, table_device, st, plan->L.tile_stride > 1);
}

int iwb_plan3d_run_exchange_async(iwb_plan3d* plan, iwb_peer* pr, void* stream, double* row_device)
{
    if (!plan || !pr || !row_device) return fail(IWB_ERR_INVALID, "iwb_plan3d_run_exchange_async: null argument");
    if (!pr->connected) return fail(IWB_ERR_INVALID, "iwb_plan3d_run_exchange_async: iwb_peer_connect has not been called");
    const Launch3& L = plan->L;
    if (L.tile_stride != pr->world || L.tile_first != pr->rank || L.P.i_begin != 0 || L.P.i_end != L.P.n0)
        return fail(IWB_ERR_INVALID, "iwb_plan3d_run_exchange_async: the plan is not this rank's tile shard of the group");
    const int nfb = (L.P.n0 + FLUSH - 1) / FLUSH;
    if (nfb > pr->nfb_max) return fail(IWB_ERR_INVALID, "iwb_plan3d_run_exchange_async: grid has more flush blocks than the peer buffers hold");
    IWB_CUDA(cudaSetDevice(plan->h->device));
    return launch_exchange(plan->h, pr, L, nfb, (cudaStream_t)stream, row_device);
}

int iwb_plan3d_destroy(iwb_plan3d* plan)
{
    if (!plan) return IWB_OK;
    cudaSetDevice(plan->h->device);
    cudaDeviceSynchronize();
    plan->coords.release(); plan->partial.release(); plan->counter.release(); plan->h_coords.release();
    delete plan;
    return IWB_OK;
}

int iwb_energy3d_batch(iwb_handle* h, int npoints, const iwb_params3d* pts, iwb_result* outs)
{
    if (!h || npoints < 0 || (npoints > 0 && (!pts || !outs))) return fail(IWB_ERR_INVALID, "iwb_energy3d_batch: bad argument");
    for (int q = 0; q < npoints; ++q) {
        int rc = validate(&pts[q]);
        if (rc) return rc;
    }
    auto t0 = std::chrono::steady_clock::now();
    IWB_CUDA(cudaSetDevice(h->device));
    // (Giving each of the NSCRATCH concurrent evaluations of a batch of small grids its own quarter of the SMs -- four
    // times the units per CTA, a quarter of the phi prologues -- was measured and lost: 50 x n = 256: 72.5 against 78.5
    // Gpoints/s, n = 400: 95.1 against 99.3, two halves: no difference; profiles/r2_ab18_batch_split.log.)
    // rolling pipeline: slot q is re-armed with the next point as soon as its previous occupant has been read back
    auto collect = [&](int slot, int idx) -> int {
        Scratch& c = h->scr[slot];
        IWB_CUDA(cudaStreamSynchronize(c.stream));
        memset(&outs[idx], 0, sizeof(iwb_result));
        unpack_row((const double*)c.h_out.p, pts[idx].nshell, &outs[idx]);
        outs[idx].cnt_nan = (int64_t)(pts[idx].i_end - pts[idx].i_begin) * pts[idx].n1 * pts[idx].n2 -
                            outs[idx].cnt_pos - outs[idx].cnt_neg - outs[idx].cnt_zero;
        apply_reference_nonfinite(&outs[idx], pts[idx].nshell);
        float ms = 0;
        IWB_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
        outs[idx].kernel_ms = ms;
        return IWB_OK;
    };
    int rc = IWB_OK;
    for (int q = 0; q < npoints && rc == IWB_OK; ++q) {
        const int slot = q % NSCRATCH;
        if (q >= NSCRATCH) rc = collect(slot, q - NSCRATCH);
        if (rc == IWB_OK) rc = enqueue_full(h, h->scr[slot], &pts[q]);
    }
    for (int q = (npoints > NSCRATCH ? npoints - NSCRATCH : 0); q < npoints && rc == IWB_OK; ++q) rc = collect(q % NSCRATCH, q);
    if (rc) {
        for (int s2 = 0; s2 < NSCRATCH; ++s2) cudaStreamSynchronize(h->scr[s2].stream);    // leave no launch behind
        return rc;
    }
    const double total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (int q = 0; q < npoints; ++q) outs[q].total_ms = total;
    return IWB_OK;
}

}  // extern "C"

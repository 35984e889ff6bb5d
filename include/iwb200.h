/*
 * iwb200.h -- C-ABI of the B200-native energy-condition engine (libiwb200.so).
 *
 * The reference (DawsonInstitute/irrotational-warp-lab) is pure Python and has no
 * FFI; its seam for this path is the keyword-only contract of three functions and
 * one backend registry.  Each entry point below replaces the array work of one of
 * them; the Python host (irrotational-warp-lab_b200/backend.py) keeps the
 * reference's signatures and return objects and calls these through ctypes.
 *
 *   iwb_energy3d*        <- compute_energy_3d      scripts/reproduce_rodal_exact.py:157-235
 *                           + phi_exact_rodal      src/irrotational_warp/potential.py:80-99
 *                           + g_rodal_exact        src/irrotational_warp/potential.py:37-77
 *                           + _split_pos_neg       scripts/reproduce_rodal_exact.py:57-61
 *                           + the _e_at_r closure  scripts/reproduce_rodal_exact.py:220-224
 *   iwb_energy_axisym    <- compute_energy_axisymmetric  scripts/reproduce_rodal_exact.py:76-154
 *   iwb_slice_z0         <- compute_slice_z0       src/irrotational_warp/adm.py:27-88
 *                           + central_diff_2d      src/irrotational_warp/fd.py:6-25
 *                           + integrate_signed     src/irrotational_warp/adm.py:91-97
 *   iwb_radial_profile3d <- compute_radial_average_z0  src/irrotational_warp/tail.py:46-80 (its 3D generalisation)
 *   iwb_einstein_z0      <- compute_einstein_eigenvalues  src/irrotational_warp/einstein.py:256-309
 *   iwb_energy3d_batch   <- the serial loops of sweep_velocity
 *                           scripts/sweep_superluminal.py:59-106 and grid_search
 *                           src/irrotational_warp/optimize.py:100-114
 *   iwb_create/device_count/last_error <- _resolve_backend
 *                           scripts/reproduce_rodal_exact.py:16-30
 *
 * Conventions: plain C, no torch / CUDA types.  All pointers are caller-owned
 * HOST memory unless a field says "device"; the library copies what it needs and
 * keeps nothing past return.  Calls are blocking unless their name ends in _async
 * (those enqueue work on the caller's stream and return; each uses its own staging
 * slot, so calls on different streams do not interfere) and return IWB_OK (0) or an
 * error code; iwb_last_error() gives the thread-local message.  A handle is bound
 * to ONE device (one process per GPU is the multi-GPU model; the host layer does
 * the NCCL collective through torch.distributed) and is not thread-safe; use one
 * handle per thread.  There is no CPU fallback: without a usable sm_100 device
 * iwb_create fails.
 */
#ifndef IWB200_H
#define IWB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IWB_VERSION 100          /* 0.1.0 */
#define IWB_MAX_SHELLS 8

enum iwb_status {
    IWB_OK = 0,
    IWB_ERR_INVALID = 1,         /* bad argument (maps to Python ValueError)        */
    IWB_ERR_GRID_TOO_SMALL = 2,  /* < edge_order+1 points on an axis (numpy.gradient's ValueError) */
    IWB_ERR_NO_DEVICE = 3,       /* no usable GPU (maps to RuntimeError, like a failed CuPy import) */
    IWB_ERR_CUDA = 4,            /* CUDA runtime error, see iwb_last_error()        */
    IWB_ERR_NOMEM = 5
};

/* arithmetic variants */
enum iwb_dtype {
    IWB_F64 = 0,       /* reference --dtype float64                                          */
    IWB_F32_REF = 1    /* reference --dtype float32 as NumPy>=2 evaluates it: float32 grid,   */
                       /* r, log-cosh and cos(theta); float64 from sinh(rho*sigma) onwards    */
                       /* (an all-float32 phi was built and measured in round 2: slower -- IEEE float32 division */
                       /* and square root plus conversions -- and 8e-4 from this mode at n = 1024, so it was     */
                       /* removed; DESIGN.md section 2, profiles/r2_f32_error_table.log)                         */
};

enum iwb_mode {
    IWB_MODE_EXACT = 0, /* IEEE replay of the reference's operation order (bit-exact masks/counts) */
    IWB_MODE_FAST = 1   /* IWB_F64 only: FMA contraction + reciprocal multiplies; energies within 1e-10, */
                        /* sign counts may differ inside the rounding-noise core (3D evaluator only)     */
};

typedef struct iwb_handle iwb_handle;

/* One 3D evaluation.  The grid is n0 x n1 x n2 points, C order [i, j, k] like the
 * reference's meshgrid(indexing="ij"); x, y, z are the host linspace arrays and
 * dx, dy, dz the host-computed float(x[1]-x[0]) (reproduce_rodal_exact.py:171-177).
 * For IWB_F32_REF the arrays hold float32-rounded values widened to double. */
typedef struct {
    int32_t dtype;             /* enum iwb_dtype */
    int32_t mode;              /* enum iwb_mode  */
    int32_t n0, n1, n2;
    int32_t i_begin, i_end;    /* planes of axis 0 to evaluate: [i_begin, i_end) of the GLOBAL grid.
                                  (0, n0) = whole volume.  A slab gets its 2-plane phi halo by
                                  analytic recomputation from the global coordinates.            */
    const double *x, *y, *z;   /* host, lengths n0, n1, n2 */
    double dx, dy, dz;
    double rho, sigma, v, eps; /* eps = 1e-12 in the reference (potential.py:88) */
    double sinh_rs, cosh_rs;   /* np.sinh / np.cosh(rho*sigma) computed by the host (potential.py:50-52) */
    int32_t nshell;            /* <= IWB_MAX_SHELLS */
    double shell_r[IWB_MAX_SHELLS]; /* radii for the inclusive masks r <= R (the _e_at_r closure) */
    double *rho_out;           /* NULL, or [i_end-i_begin, n1, n2] doubles receiving rho_adm          */
    int32_t rho_out_is_device; /* 1: rho_out is a device pointer on the handle's device             */
    int32_t tile_stride;       /* 0 or 1: every tile of the (j, k) plane.  S in {2, 4, 8, 16}: an interleaved */
    int32_t tile_first;        /* tile shard, the tiles whose schedule index is tile_first (mod S) -- the     */
                               /* multi-GPU decomposition (iwb_energy3d_tiles_async); rho_out must be NULL   */
} iwb_params3d;

typedef struct {
    double e_pos, e_neg;                 /* sum e*[e>0], sum e*[e<0] (negative), e = rho_adm*dV */
    int64_t cnt_pos, cnt_neg, cnt_zero, cnt_nan;
    double shell_pos[IWB_MAX_SHELLS], shell_neg[IWB_MAX_SHELLS];
    int64_t shell_cnt[IWB_MAX_SHELLS];   /* points with r <= R */
    double kernel_ms;                    /* CUDA-event time of the fused kernel + final reduction */
    double total_ms;                     /* host wall time of the call */
} iwb_result;

/* Fixed-layout partial sums of one slab, for the shard-count-independent reduction:
 * one row per "flush block" of IWB_FLUSH_PLANES planes of axis 0, IWB_ROW_WIDTH doubles per row
 * (int64 counts are stored bit-cast in double slots).  See DESIGN.md §5. */
#ifndef IWB_FLUSH_PLANES
#define IWB_FLUSH_PLANES 16
#endif
#define IWB_ROW_WIDTH (6 + 3 * IWB_MAX_SHELLS)   /* e_pos e_neg cnt_pos cnt_neg cnt_zero cnt_nan | shell_pos[8] shell_neg[8] shell_cnt[8] */

int iwb_version(void);
int iwb_device_count(void);
const char *iwb_last_error(void);

/* device_id: CUDA ordinal.  Fails with IWB_ERR_NO_DEVICE when there is none or it is not sm_100. */
int iwb_create(int device_id, iwb_handle **out);
int iwb_destroy(iwb_handle *h);
int iwb_device_info(iwb_handle *h, int *sm_count, int *cc_major, int *cc_minor, char *name, int name_len);

/* Whole evaluation: kernel, deterministic final reduction, D2H of the scalars. */
int iwb_energy3d(iwb_handle *h, const iwb_params3d *p, iwb_result *out);

/* Slab evaluation for the multi-GPU path.  Writes the per-flush-block table
 * rows [i_begin/IWB_FLUSH_PLANES, ceil(i_end/IWB_FLUSH_PLANES)) x IWB_ROW_WIDTH into
 * `table_device` (DEVICE pointer to the full [ceil(n0/IWB_FLUSH_PLANES), IWB_ROW_WIDTH] table; rows of
 * other slabs are not touched) on `stream` (a cudaStream_t passed as void*; NULL = CUDA's default
 * stream) WITHOUT synchronising.  i_begin must be a multiple of IWB_FLUSH_PLANES. */
int iwb_energy3d_slab_async(iwb_handle *h, const iwb_params3d *p, double *table_device, void *stream);

/* Interleaved tile shards, the preferred multi-GPU decomposition: shard r of S (p->tile_first = r,
 * p->tile_stride = S in {1, 2, 4, 8, 16}) evaluates every S-th tile of the (j, k) plane along the WHOLE of
 * axis 0 -- no halo along axis 0, and every shard gets the same mix of face / wall / shell tiles.  The tile
 * sums of a flush block are formed as IWB_TILE_CHAINS chains (chain c adds the tiles c, c + 16, ... of the
 * schedule order) that a fixed binary tree combines; a shard owns the chains c = r (mod S) entirely and
 * writes them to `chains_device`, [ceil(n0/IWB_FLUSH_PLANES)][IWB_TILE_CHAINS / S][IWB_ROW_WIDTH] doubles.
 * After an all-gather (shard-major) iwb_reduce_chains applies the tree and the row sum: the same additions
 * in the same order as the single-GPU path, so the result is bit-identical for every S.  No synchronisation. */
#define IWB_TILE_CHAINS 16
int iwb_energy3d_tiles_async(iwb_handle *h, const iwb_params3d *p, double *chains_device, void *stream);
/* gathered_device: [S][nfb][IWB_TILE_CHAINS / S][IWB_ROW_WIDTH] (DEVICE).  Blocking: the scalars are on the host at return. */
int iwb_reduce_chains(iwb_handle *h, const double *gathered_device, int tile_stride, int nfb, int nshell, void *stream,
                      iwb_result *out);

/* The tile-shard evaluation with the exchange fused in: instead of an NCCL all-gather, every rank's chain sums cross
 * NVLink / NVSwitch as plain peer stores into the gathered buffer of every rank, followed by an epoch-flag handshake and
 * the same chain tree + row sum as the single-GPU path (bit-identical energies on every rank, for every S).  Set-up, once
 * per process group: every rank calls iwb_peer_create (allocates its buffers for grids of up to nfb_max flush blocks and
 * returns an IWB_PEER_HANDLE_BYTES IPC handle), the HOST layer all-gathers the handles (rank-major), every rank calls
 * iwb_peer_connect.  All ranks must reside on one node with peer access between their GPUs.
 * iwb_energy3d_tiles_exchange_async is COLLECTIVE: every rank of the group calls it the same number of times with
 * tile_stride = world, tile_first = rank; it enqueues k_energy3d, the push of the chain sums and the combine on `stream`
 * and leaves the final row (IWB_ROW_WIDTH doubles) in `row_device` without synchronising.  iwb_read_row copies such a row
 * to the host (blocking) and unpacks it.  A rank whose peers do not signal within 60 s (a peer process died, or did not
 * make the call) traps instead of spinning for ever: the next synchronisation on `stream` reports a CUDA error. */
#define IWB_PEER_HANDLE_BYTES 64
typedef struct iwb_peer iwb_peer;
int iwb_peer_create(iwb_handle *h, int rank, int world, int nfb_max, iwb_peer **out, void *handle_out);
int iwb_peer_connect(iwb_peer *peer, const void *all_handles);
int iwb_peer_destroy(iwb_peer *peer);
int iwb_energy3d_tiles_exchange_async(iwb_handle *h, iwb_peer *peer, const iwb_params3d *p, void *stream, double *row_device);
int iwb_read_row(iwb_handle *h, const double *row_device, int nshell, void *stream, iwb_result *out);
/* Test hook: connects peers that live in ONE process on one device (IPC handles cannot be opened by the process that
 * exported them) so that the single-GPU test-suite can run the exchange kernels with emulated ranks, one stream each. */
int iwb_debug_peer_connect_local(iwb_peer *peer, iwb_peer *const *all_peers);

/* Non-blocking halves of iwb_reduce_table / iwb_reduce_chains for callers that keep several evaluations in flight on one
 * stream: the final row (IWB_ROW_WIDTH doubles, layout of the table rows) is left in `row_device`; nothing is copied to
 * the host and nothing is synchronised.  `table_scratch_device` receives the [nfb][IWB_ROW_WIDTH] table of the chain tree. */
int iwb_reduce_table_async(iwb_handle *h, const double *table_device, int nrows, void *stream, double *row_device);
int iwb_reduce_chains_async(iwb_handle *h, const double *gathered_device, int tile_stride, int nfb, void *stream,
                            double *table_scratch_device, double *row_device);

/* Prepared evaluation: coordinates uploaded and constants derived once, then launched any number
 * of times with NO host<->device traffic (kernel + tile reduction into `table_device` only).
 * This is what a resident-input measurement and a repeated-evaluation driver use.  rho_out, if
 * set, must be a device pointer. */
typedef struct iwb_plan3d iwb_plan3d;
int iwb_plan3d_create(iwb_handle *h, const iwb_params3d *p, iwb_plan3d **out);
/* table_device: the slab table of iwb_energy3d_slab_async, or -- for a plan with tile_stride > 1 -- the chain
 * buffer of iwb_energy3d_tiles_async. */
int iwb_plan3d_run_async(iwb_plan3d *plan, double *table_device, void *stream);
/* the prepared tile shard of a peer group (tile_stride = world, tile_first = rank) with the fused exchange: collective,
 * no host<->device traffic, result row in `row_device` (see iwb_energy3d_tiles_exchange_async) */
int iwb_plan3d_run_exchange_async(iwb_plan3d *plan, iwb_peer *peer, void *stream, double *row_device);
int iwb_plan3d_destroy(iwb_plan3d *plan);

/* Fixed-order sum of a complete table (device pointer) into *out; blocking. */
int iwb_reduce_table(iwb_handle *h, const double *table_device, int nrows, int nshell, void *stream,
                     iwb_result *out);

/* Independent parameter points on one device; one host synchronisation at the end. */
int iwb_energy3d_batch(iwb_handle *h, int npoints, const iwb_params3d *pts, iwb_result *outs);

/* Radial-shell profile of rho_adm on the 3D grid: sum and point count over edges[k] <= r < edges[k+1], the binning
 * rule of compute_radial_average_z0 (src/irrotational_warp/tail.py:46-80) applied to the field of compute_energy_3d
 * (SURVEY.md §8 f3).  Large volumes (>= 2^29 points, <= 144 bins, <= 2 shells) are binned inside the fused kernel itself
 * (k_energy3d_prof); otherwise the field is emitted and re-read in passes of <= 1 GiB on the device.  It never reaches the host.
 * `out` (may be NULL) receives the energies / counts of the same evaluation.  Deterministic and independent of the
 * pass size; sums over [i_begin, i_end) only, so slabs can be added on the host. */
typedef struct {
    int32_t n_bins;            /* 1 .. 256 */
    const double *edges;       /* host [n_bins + 1], strictly increasing (np.linspace(0, r_max, n_bins + 1)) */
    double *sum;               /* host out [n_bins] */
    int64_t *count;            /* host out [n_bins] */
} iwb_profile3d;

int iwb_radial_profile3d(iwb_handle *h, const iwb_params3d *p, const iwb_profile3d *prof, iwb_result *out);

/* Axisymmetric half-plane, arrays [ny, nx] (x fastest), z = 0. */
typedef struct {
    int32_t dtype, mode;
    int32_t nx, ny;
    const double *x, *y;       /* host, lengths nx, ny */
    double dx, dy;
    double rho, sigma, v, eps;
    double sinh_rs, cosh_rs;
    int32_t nshell;
    double shell_r[IWB_MAX_SHELLS];
    double *rho_out;           /* NULL or host [ny, nx] */
} iwb_params_axisym;

int iwb_energy_axisym(iwb_handle *h, const iwb_params_axisym *p, iwb_result *out);

/* 2D z=0 slice with the tanh-wall dipole potential and edge_order=1 differences.
 * All four field buffers are host [n, n] doubles in [y, x] order and must be non-NULL
 * (SliceResult carries them).  pos / neg / net follow integrate_signed: neg >= 0. */
typedef struct {
    int32_t n;
    const double *x, *y;
    double dx, dy;
    double rho, sigma, v, eps;
    double *phi, *beta_x, *beta_y, *rho_adm;
} iwb_params_slice;

typedef struct {
    double pos, neg, net;
    int64_t cnt_pos, cnt_neg, cnt_zero;
    double kernel_ms, total_ms;
} iwb_slice_result;

int iwb_slice_z0(iwb_handle *h, const iwb_params_slice *p, iwb_slice_result *out);

/* Einstein-tensor eigenvalue diagnostic of the z = 0 slice (compute_einstein_eigenvalues,
 * src/irrotational_warp/einstein.py:256-309, with its helpers :41-253): G^mu_nu of the metric built from the shift
 * (beta_x, beta_y) -- host [ny, nx] doubles, e.g. the SliceResult fields -- and its four eigenvalues per grid point.
 * Outputs are host buffers: eig_re / eig_im [4, ny, nx] (eigenvalue order is not LAPACK's; entry 3 is the decoupled
 * G^3_3 = -R/2), eig_max and ricci_scalar [ny, nx], is_type_i [ny, nx] bytes (all |Im| < imag_tol; 1e-8 in the
 * reference).  Parity with the reference is to rounding (its inverse and eigenvalues are LAPACK's), see DESIGN.md. */
typedef struct {
    int32_t nx, ny;
    const double *beta_x, *beta_y;
    double dx, dy;
    double imag_tol;
    double *eig_re, *eig_im;
    double *eig_max, *ricci_scalar;
    uint8_t *is_type_i;
} iwb_params_einstein;

typedef struct {
    int64_t cnt_type_i;
    double kernel_ms, total_ms;
} iwb_einstein_result;

int iwb_einstein_z0(iwb_handle *h, const iwb_params_einstein *p, iwb_einstein_result *out);

/* Measured FP64 / FP32 FMA pipe peak of the handle's device (register-resident microbenchmark),
 * in TFLOP/s counting 2 flops per FMA; used as the roofline denominator by bench.py. */
int iwb_measure_fma_peak(iwb_handle *h, int use_fp32, double *tflops, double *sm_mhz_effective);

/* Test hook: compares the constant-divisor division used by the exact kernels with IEEE division
 * on `count` pseudo-random numerators; returns the number of mismatching results in *mismatches. */
int iwb_selftest_constdiv(iwb_handle *h, double divisor, uint64_t count, uint64_t seed, uint64_t *mismatches);

/* Host-only views of the fused kernel's work decomposition (no device needed; used by the CPU test-suite).
 * iwb_debug_spans: the span boundaries over `units` work units for `ctas` persistent CTAs with `static_pct` % of
 * the units in one large span per CTA (see k_energy3d); writes at most `cap` ints to `out`, returns the number of
 * boundaries (spans + 1).  iwb_debug_tiles_k: the first column of each tile of an n-column axis, then n; returns
 * the number of tiles (boundaries = tiles + 1), writing at most `cap` ints. */
int iwb_debug_spans(long long units, int ctas, int static_pct, int *out, int cap);
int iwb_debug_tiles_k(int n, int *out, int cap);
/* the same for the rows (j) of the default 40-row region: 36 output rows per interior tile, 38 in the tiles at the j faces */
int iwb_debug_tiles_j(int n, int *out, int cap);
/* Which path the last iwb_radial_profile3d call on this handle took: 1 = bins accumulated inside the fused kernel,
 * 2 = rho emitted and binned in a second pass (volumes below 2^29 points, where that is faster; more than 144 bins; more
 * than two shells; bins narrower than a work unit's window of 14; IWB_PROFILE_FUSED=0), 0 = no call yet.
 * IWB_PROFILE_FUSED=2 takes the fused kernel whenever it can.  Diagnostic for the tests. */
int iwb_debug_profile_path(iwb_handle *h);

/* Test hook: the branch-free division / square-root sequences of the exact kernels against the built-in
 * IEEE operators on `count` pseudo-random operand pairs (wide_exponents != 0 also exercises the range
 * check that routes exceptional operands to the built-in sequence; those pairs are counted in *skipped). */
int iwb_selftest_divsqrt(iwb_handle *h, uint64_t count, uint64_t seed, int wide_exponents, uint64_t *mismatches_div,
                         uint64_t *mismatches_sqrt, uint64_t *skipped);

/* Test hook: the branch-free phi evaluation of the exact 3D kernel (square root, the two divisions with reciprocal seeds
 * taken from the square root's 1/sqrt, the wall flavour with exp / log1p) against the plain IEEE operators on `count`
 * pseudo-random points whose coordinates are log-uniform in +-[coord_lo, coord_hi); *mismatches counts results that
 * differ in any bit, *wall_points (may be NULL) the points that took the exp / log1p flavour. */
int iwb_selftest_phi(iwb_handle *h, double rho, double sigma, double v, double eps, double coord_lo, double coord_hi,
                     uint64_t count, uint64_t seed, uint64_t *mismatches, uint64_t *wall_points);

#ifdef __cplusplus
}
#endif
#endif /* IWB200_H */

/*
 * CPU oracle (plain C + pthreads) for the 3D / axisymmetric energy-condition path.
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg -- never by the product library.
 *
 * Pointwise restatement of the reference's NumPy evaluation, one IEEE operation
 * per NumPy ufunc, in the reference's operand order (built with
 * -ffp-contract=off, see oracle/Makefile).  Reference lines followed
 * (/root/reference/...):
 *   src/irrotational_warp/potential.py:48-77   g(r)  (logaddexp -> glibc exp/log1p)
 *   src/irrotational_warp/potential.py:95-99   phi = v r g x/(r+eps)
 *   scripts/reproduce_rodal_exact.py:168-224   3D: gradient(gradient(phi)), K, rho, sums
 *   scripts/reproduce_rodal_exact.py:93-143    axisym: K_zz = -phi_y/y, dV = 2 pi y dx dy
 *   numpy.gradient, uniform spacing, edge_order=2 (numpy/lib/_function_base_impl.py)
 *   numpy.logaddexp (npy_math: x==y -> x+ln2; else larger + log1p(exp(-|x-y|)))
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks it against the
 * vectors frozen from the live reference (tests/golden/make_golden.py): sign
 * counts and shell point counts exactly, sampled e-density values bit-for-bit,
 * energies to 1e-13 (summation order differs from NumPy's pairwise sum).
 *
 * dtype_mode 0 = float64.  dtype_mode 1 = the reference's "--dtype float32"
 * behaviour under NumPy >= 2 (coordinates, r, sigma*(r-+rho), logaddexp and
 * cos(theta) in float32, everything downstream of the first np.float64 scalar
 * in float64).  Coordinates are always passed as doubles
 * holding the (possibly float32-rounded) linspace values.
 *
 * Slab entry point: planes [i_begin, i_end) of the global grid, phi evaluated on
 * a 2-plane overlap from GLOBAL coordinates, so disjoint slabs add up exactly
 * to the whole (integer counts) or to rounding (sums).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

#define ORACLE_MAX_SHELLS 8
#define LOGE2 0.693147180559945309417232121458176568
#define LOGE2F 0.693147180559945309417232121458176568F

typedef struct {
    double e_pos, e_neg;
    int64_t cnt_pos, cnt_neg, cnt_zero;
    double shell_pos[ORACLE_MAX_SHELLS], shell_neg[ORACLE_MAX_SHELLS];
    int64_t shell_cnt[ORACLE_MAX_SHELLS];
} oracle_result;

typedef struct {
    double rho, sigma, v;
    double sinh_rs, cosh_rs; /* np.sinh / np.cosh of rho*sigma, from the host */
    int dtype_mode;
} oracle_params;

/* ---- potential.py:62-63, np.logaddexp(a, -a) ------------------------------ */
static double lae_pm(double a)
{
    double b = -a;
    if (a == b) return a + LOGE2;
    double t = a - b;
    if (t > 0) return a + log1p(exp(-t));
    if (t <= 0) return b + log1p(exp(t));
    return t;
}

static float lae_pmf(float a)
{
    float b = -a;
    if (a == b) return a + LOGE2F;
    float t = a - b;
    if (t > 0) return a + log1pf(expf(-t));
    if (t <= 0) return b + log1pf(expf(t));
    return t;
}

/* ---- potential.py:96-99 with g from :48-77 --------------------------------- */
static double phi64(const oracle_params *p, double x, double y, double z)
{
    double r = sqrt(x * x + y * y + z * z);
    double t1 = 2 * r * p->sigma * p->sinh_rs;
    double am = p->sigma * (r - p->rho);
    double ap = p->sigma * (r + p->rho);
    double d = lae_pm(am) - lae_pm(ap);
    double t2 = p->cosh_rs * d;
    double num = t1 + t2;
    double den = 2 * r * p->sigma * p->sinh_rs;
    double g = (fabs(den) > 1e-12) ? num / den : 0.0;
    double c = x / (r + 1e-12);
    return p->v * r * g * c;
}

static double phi32mixed(const oracle_params *p, float x, float y, float z)
{
    float sig = (float)p->sigma, rho = (float)p->rho, vf = (float)p->v;
    float r = sqrtf(x * x + y * y + z * z);
    double t1 = (double)(2 * r * sig) * p->sinh_rs;
    float am = sig * (r - rho);
    float ap = sig * (r + rho);
    float d = lae_pmf(am) - lae_pmf(ap);
    double t2 = p->cosh_rs * (double)d;
    double num = t1 + t2;
    double den = (double)(2 * r * sig) * p->sinh_rs;
    double g = (fabs(den) > 1e-12) ? num / den : 0.0;
    float c = x / (r + (float)1e-12);
    return (double)(vf * r) * g * (double)c;
}

static double phi_at(const oracle_params *p, double x, double y, double z)
{
    return p->dtype_mode ? phi32mixed(p, (float)x, (float)y, (float)z) : phi64(p, x, y, z);
}

static int in_shell(const oracle_params *p, double x, double y, double z, double r_lim)
{
    if (p->dtype_mode) {
        float xf = (float)x, yf = (float)y, zf = (float)z;
        return sqrtf(xf * xf + yf * yf + zf * zf) <= (float)r_lim;
    }
    return sqrt(x * x + y * y + z * z) <= r_lim;
}

/* ---- numpy.gradient, one axis, uniform spacing, edge_order 2 --------------- */
typedef struct { double h2, a0, b0, c0, a1, b1, c1; } dcoef;

static dcoef make_coef(double h)
{
    dcoef c;
    c.h2 = 2. * h;
    c.a0 = -1.5 / h; c.b0 = 2. / h; c.c0 = -0.5 / h;
    c.a1 = 0.5 / h;  c.b1 = -2. / h; c.c1 = 1.5 / h;
    return c;
}

/* f points at element `p` of a line of length n with element stride s */
static inline double dline(const double *f, long p, long n, long s, const dcoef *c,
                           int lo_edge, int hi_edge)
{
    if (p == 0)
        return lo_edge ? c->a0 * f[0] + c->b0 * f[s] + c->c0 * f[2 * s] : NAN;
    if (p == n - 1)
        return hi_edge ? c->a1 * f[-2 * s] + c->b1 * f[-s] + c->c1 * f[0] : NAN;
    return (f[s] - f[-s]) / c->h2;
}

static void accumulate(oracle_result *acc, const oracle_result *part, int nshell)
{
    acc->e_pos += part->e_pos; acc->e_neg += part->e_neg;
    acc->cnt_pos += part->cnt_pos; acc->cnt_neg += part->cnt_neg; acc->cnt_zero += part->cnt_zero;
    for (int s = 0; s < nshell; ++s) {
        acc->shell_pos[s] += part->shell_pos[s];
        acc->shell_neg[s] += part->shell_neg[s];
        acc->shell_cnt[s] += part->shell_cnt[s];
    }
}

static const oracle_result ZERO_RESULT;

/* ---- tiny pthread parallel-for (libgomp is not in this image) --------------- */
typedef void (*range_fn)(void *ctx, long begin, long end, int tid);
typedef struct { range_fn fn; void *ctx; long begin, end; int tid; } pf_job;

static void *pf_trampoline(void *arg)
{
    pf_job *j = arg;
    j->fn(j->ctx, j->begin, j->end, j->tid);
    return NULL;
}

static int g_threads = 0;
void oracle_set_threads(int t) { g_threads = t; }
int oracle_get_threads(void)
{
    if (g_threads > 0) return g_threads;
    const char *e = getenv("ORACLE_THREADS");
    int t = e ? atoi(e) : (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (t < 1) t = 1;
    if (t > 256) t = 256;
    return t;
}

static void parallel_for(long n, range_fn fn, void *ctx)
{
    int t = oracle_get_threads();
    if (t > n) t = n > 0 ? (int)n : 1;
    pthread_t th[256];
    pf_job jobs[256];
    for (int i = 0; i < t; ++i) {
        jobs[i] = (pf_job){fn, ctx, n * i / t, n * (i + 1) / t, i};
        if (i > 0) pthread_create(&th[i], NULL, pf_trampoline, &jobs[i]);
    }
    pf_trampoline(&jobs[0]);
    for (int i = 1; i < t; ++i) pthread_join(th[i], NULL);
}

/* ---- reproduce_rodal_exact.py:168-224 on planes [i_begin, i_end) ------------ */
typedef struct {
    const oracle_params *p;
    long n1, n2, m0, lo, s0, s1, own0, own1;
    int e_lo, e_hi, nshell;
    const double *x, *y, *z, *shell_r;
    double dV;
    dcoef cx, cy, cz;
    double *phi, *px, *py, *pz, *rho_out;
    oracle_result *partial; /* one per thread */
} ctx3d;

static void k3_phi(void *v, long b, long e, int tid)
{
    ctx3d *c = v; (void)tid;
    for (long row = b; row < e; ++row) {          /* row = i * n1 + j */
        long i = row / c->n1, j = row % c->n1;
        for (long k = 0; k < c->n2; ++k)
            c->phi[row * c->s1 + k] = phi_at(c->p, c->x[c->lo + i], c->y[j], c->z[k]);
    }
}

static void k3_grad(void *v, long b, long e, int tid)
{
    ctx3d *c = v; (void)tid;
    for (long row = b; row < e; ++row) {
        long i = row / c->n1, j = row % c->n1;
        for (long k = 0; k < c->n2; ++k) {
            long q = row * c->s1 + k;
            c->px[q] = dline(c->phi + q, i, c->m0, c->s0, &c->cx, c->e_lo, c->e_hi);
            c->py[q] = dline(c->phi + q, j, c->n1, c->s1, &c->cy, 1, 1);
            c->pz[q] = dline(c->phi + q, k, c->n2, 1, &c->cz, 1, 1);
        }
    }
}

static void tally(oracle_result *loc, const oracle_params *p, double e, int nshell,
                  const double *shell_r, double x, double y, double z)
{
    int pos = e > 0.0, neg = e < 0.0;
    if (pos) { loc->e_pos += e; loc->cnt_pos++; }
    if (neg) { loc->e_neg += e; loc->cnt_neg++; }
    if (e == 0.0) loc->cnt_zero++;
    for (int s = 0; s < nshell; ++s)
        if (in_shell(p, x, y, z, shell_r[s])) {
            loc->shell_cnt[s]++;
            if (pos) loc->shell_pos[s] += e;
            if (neg) loc->shell_neg[s] += e;
        }
}

static void k3_rho(void *v, long b, long e, int tid)
{
    ctx3d *c = v;
    const double sixteen_pi = 16.0 * 3.141592653589793;
    oracle_result loc = ZERO_RESULT;
    for (long row = b; row < e; ++row) {          /* rows of the owned planes only */
        long i = c->own0 + row / c->n1, j = row % c->n1;
        for (long k = 0; k < c->n2; ++k) {
            long q = i * c->s0 + j * c->s1 + k;
            double kxx = -dline(c->px + q, i, c->m0, c->s0, &c->cx, c->e_lo, c->e_hi);
            double kxy = -dline(c->px + q, j, c->n1, c->s1, &c->cy, 1, 1);
            double kxz = -dline(c->px + q, k, c->n2, 1, &c->cz, 1, 1);
            double kyy = -dline(c->py + q, j, c->n1, c->s1, &c->cy, 1, 1);
            double kyz = -dline(c->py + q, k, c->n2, 1, &c->cz, 1, 1);
            double kzz = -dline(c->pz + q, k, c->n2, 1, &c->cz, 1, 1);
            double tr = kxx + kyy + kzz;
            double kk = kxx * kxx + kyy * kyy + kzz * kzz
                      + 2.0 * (kxy * kxy + kxz * kxz + kyz * kyz);
            double rho_adm = (tr * tr - kk) / sixteen_pi;
            double en = rho_adm * c->dV;
            if (c->rho_out) c->rho_out[(i - c->own0) * c->s0 + j * c->s1 + k] = rho_adm;
            tally(&loc, c->p, en, c->nshell, c->shell_r, c->x[c->lo + i], c->y[j], c->z[k]);
        }
    }
    c->partial[tid] = loc;
}

int oracle_energy3d_slab(const oracle_params *p, int n0, int n1, int n2,
                         const double *x, const double *y, const double *z,
                         double dx, double dy, double dz,
                         int i_begin, int i_end,
                         int nshell, const double *shell_r,
                         double *rho_out, /* [i_end-i_begin, n1, n2] or NULL */
                         oracle_result *out)
{
    if (n0 < 3 || n1 < 3 || n2 < 3) return 1; /* numpy.gradient raises ValueError */
    if (nshell > ORACLE_MAX_SHELLS || i_begin < 0 || i_end > n0 || i_begin >= i_end) return 2;
    long lo = i_begin - 2 < 0 ? 0 : i_begin - 2;
    long hi = i_end + 2 > n0 ? n0 : i_end + 2;
    if (lo == 0 && hi < 4) hi = n0 < 4 ? n0 : 4;
    if (hi == n0 && lo > n0 - 4) lo = n0 - 4 < 0 ? 0 : n0 - 4;
    ctx3d c;
    c.p = p; c.n1 = n1; c.n2 = n2; c.m0 = hi - lo; c.lo = lo;
    c.s0 = (long)n1 * n2; c.s1 = n2; c.own0 = i_begin - lo; c.own1 = i_end - lo;
    c.e_lo = (lo == 0); c.e_hi = (hi == n0); c.nshell = nshell;
    c.x = x; c.y = y; c.z = z; c.shell_r = shell_r;
    c.dV = dx * dy * dz;
    c.cx = make_coef(dx); c.cy = make_coef(dy); c.cz = make_coef(dz);
    size_t bytes = sizeof(double) * (size_t)c.m0 * (size_t)c.s0;
    c.phi = malloc(bytes); c.px = malloc(bytes); c.py = malloc(bytes); c.pz = malloc(bytes);
    c.rho_out = rho_out;
    oracle_result partial[256];
    for (int t = 0; t < 256; ++t) partial[t] = ZERO_RESULT;
    c.partial = partial;
    if (!c.phi || !c.px || !c.py || !c.pz) { free(c.phi); free(c.px); free(c.py); free(c.pz); return 3; }

    parallel_for(c.m0 * n1, k3_phi, &c);
    parallel_for(c.m0 * n1, k3_grad, &c);
    parallel_for((c.own1 - c.own0) * n1, k3_rho, &c);

    oracle_result tot = ZERO_RESULT;
    for (int t = 0; t < 256; ++t) accumulate(&tot, &partial[t], nshell);
    *out = tot;
    free(c.phi); free(c.px); free(c.py); free(c.pz);
    return 0;
}

/* ---- reproduce_rodal_exact.py:93-143: arrays [ny][nx], z = 0 ---------------- */
typedef struct {
    const oracle_params *p;
    long nx, ny;
    int nshell;
    const double *x, *y, *shell_r;
    double dx, dy;
    dcoef cx, cy;
    double *phi, *px, *py, *rho_out;
    oracle_result *partial;
} ctxax;

static void ka_phi(void *v, long b, long e, int tid)
{
    ctxax *c = v; (void)tid;
    for (long j = b; j < e; ++j)
        for (long i = 0; i < c->nx; ++i)
            c->phi[j * c->nx + i] = phi_at(c->p, c->x[i], c->y[j], 0.0);
}

static void ka_grad(void *v, long b, long e, int tid)
{
    ctxax *c = v; (void)tid;
    for (long j = b; j < e; ++j)
        for (long i = 0; i < c->nx; ++i) {
            long q = j * c->nx + i;
            c->py[q] = dline(c->phi + q, j, c->ny, c->nx, &c->cy, 1, 1);
            c->px[q] = dline(c->phi + q, i, c->nx, 1, &c->cx, 1, 1);
        }
}

static void ka_rho(void *v, long b, long e, int tid)
{
    ctxax *c = v;
    const double two_pi = 2.0 * 3.141592653589793;
    const double sixteen_pi = 16.0 * 3.141592653589793;
    oracle_result loc = ZERO_RESULT;
    for (long j = b; j < e; ++j)
        for (long i = 0; i < c->nx; ++i) {
            long q = j * c->nx + i;
            double kxx = -dline(c->px + q, i, c->nx, 1, &c->cx, 1, 1);
            double kyy = -dline(c->py + q, j, c->ny, c->nx, &c->cy, 1, 1);
            double kxy = -dline(c->py + q, i, c->nx, 1, &c->cx, 1, 1);   /* d/dx of phi_y (:113) */
            double kzz = (c->y[j] != 0.0 && j != 0) ? (-c->py[q]) / c->y[j] : kyy;
            double tr = kxx + kyy + kzz;
            double kk = kxx * kxx + kyy * kyy + kzz * kzz + 2.0 * kxy * kxy;
            double rho_adm = (tr * tr - kk) / sixteen_pi;
            double dV;
            if (c->p->dtype_mode) /* 2.0*pi*Y*dx*dy with a float32 Y stays float32 */
                dV = (double)((float)two_pi * (float)c->y[j] * (float)c->dx * (float)c->dy);
            else
                dV = two_pi * c->y[j] * c->dx * c->dy;
            double en = rho_adm * dV;
            if (c->rho_out) c->rho_out[q] = rho_adm;
            tally(&loc, c->p, en, c->nshell, c->shell_r, c->x[i], c->y[j], 0.0);
        }
    c->partial[tid] = loc;
}

int oracle_energy_axisym(const oracle_params *p, int nx, int ny,
                         const double *x, const double *y, double dx, double dy,
                         int nshell, const double *shell_r,
                         double *rho_out, /* [ny, nx] or NULL */
                         oracle_result *out)
{
    if (nx < 3 || ny < 3) return 1;
    if (nshell > ORACLE_MAX_SHELLS) return 2;
    ctxax c;
    c.p = p; c.nx = nx; c.ny = ny; c.nshell = nshell;
    c.x = x; c.y = y; c.shell_r = shell_r; c.dx = dx; c.dy = dy;
    c.cx = make_coef(dx); c.cy = make_coef(dy);
    size_t bytes = sizeof(double) * (size_t)nx * (size_t)ny;
    c.phi = malloc(bytes); c.px = malloc(bytes); c.py = malloc(bytes);
    c.rho_out = rho_out;
    oracle_result partial[256];
    for (int t = 0; t < 256; ++t) partial[t] = ZERO_RESULT;
    c.partial = partial;
    if (!c.phi || !c.px || !c.py) { free(c.phi); free(c.px); free(c.py); return 3; }
    parallel_for(ny, ka_phi, &c);
    parallel_for(ny, ka_grad, &c);
    parallel_for(ny, ka_rho, &c);
    oracle_result tot = ZERO_RESULT;
    for (int t = 0; t < 256; ++t) accumulate(&tot, &partial[t], nshell);
    *out = tot;
    free(c.phi); free(c.px); free(c.py);
    return 0;
}

int oracle_max_shells(void) { return ORACLE_MAX_SHELLS; }

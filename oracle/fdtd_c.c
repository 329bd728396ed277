/* fdtd_c.c -- plain-C restatement of the Phonomena FDTD time step.  TEST INFRASTRUCTURE ONLY.
 *
 * Second, independent CPU oracle next to oracle/fdtd_numpy.py: index loops instead of NumPy
 * slices, per-cell material looked up through (id -> 6x6 table) instead of the dense
 * C[x,y,z,6,6] array, so that grids the NumPy oracle cannot hold (256^3 needs ~17 GB there) can
 * still be checked.  Nothing under phonomena_b200/ links or loads this file.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks it bit-for-bit against the golden
 * fixtures produced by the unmodified reference (tests/golden/, oracle/gen_golden.py).  It is
 * compiled with -ffp-contract=off: every operation is a separately rounded IEEE double op in the
 * reference's order, which is what NumPy evaluates.
 *
 * Follows phonomena/simulation/base_solver.py of the reference:
 *   oc_step():  :245-256 (source, update_T, update_T_BC, update_u, update_u_BC, time_step)
 *   update_T :323-372, apply_T_tfbc :402-433, update_u :435-463, apply_u_tfbc :488-517,
 *   apply_u_abc :519-554, time_step :556-571.   Array shapes: grid.py:88-110.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int nx, ny, nz, nmat;
    double *fdx, *fdy, *fdz, *sdx, *sdy, *sdz;
    uint8_t *ids;          /* (nx, ny, nz) */
    double *tab;           /* nmat x 36 */
    double *rho;           /* nmat */
    double dt, d2;         /* d2 = dt**2 as evaluated by the caller in Python (base_solver.py:443) */
    double *ux, *uy, *uz, *uxn, *uyn, *uzn, *uxo, *uyo, *uzo;
    double *T1, *T2, *T3, *T4, *T5, *T6;
} oc_t;

#define UX(a, i, j, k) a[((size_t)(i) * ny + (j)) * nz + (k)]             /* (nx-1, ny,   nz  ) */
#define UY(a, i, j, k) a[((size_t)(i) * (ny - 1) + (j)) * nz + (k)]       /* (nx,   ny-1, nz  ) */
#define UZ(a, i, j, k) a[((size_t)(i) * ny + (j)) * (nz - 1) + (k)]       /* (nx,   ny,   nz-1) */
#define TN(a, i, j, k) a[((size_t)(i) * ny + (j)) * nz + (k)]             /* T1..T3 (nx, ny, nz) */
#define T4_(i, j, k) o->T4[((size_t)(i) * (ny - 1) + (j)) * (nz - 1) + (k)]
#define T5_(i, j, k) o->T5[((size_t)(i) * ny + (j)) * (nz - 1) + (k)]
#define T6_(i, j, k) o->T6[((size_t)(i) * (ny - 1) + (j)) * nz + (k)]
#define CC(i, j, k, r, c) o->tab[(size_t)o->ids[((size_t)(i) * ny + (j)) * nz + (k)] * 36 + (r) * 6 + (c)]
#define PP(i, j, k) o->rho[o->ids[((size_t)(i) * ny + (j)) * nz + (k)]]

static double *zalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

void *oc_create(int nx, int ny, int nz, const double *fdx, const double *fdy, const double *fdz, const double *sdx,
                const double *sdy, const double *sdz, const uint8_t *ids, int nmat, const double *tab36,
                const double *rho, double dt, double d2) {
    oc_t *o = (oc_t *)calloc(1, sizeof(oc_t));
    o->nx = nx; o->ny = ny; o->nz = nz; o->nmat = nmat; o->dt = dt; o->d2 = d2;
#define DUP(dst, src, n) dst = (double *)malloc(sizeof(double) * (n)); memcpy(dst, src, sizeof(double) * (n));
    DUP(o->fdx, fdx, nx - 1) DUP(o->fdy, fdy, ny - 1) DUP(o->fdz, fdz, nz - 1)
    DUP(o->sdx, sdx, nx - 2) DUP(o->sdy, sdy, ny - 2) DUP(o->sdz, sdz, nz - 2)
    DUP(o->tab, tab36, (size_t)nmat * 36) DUP(o->rho, rho, nmat)
    o->ids = (uint8_t *)malloc((size_t)nx * ny * nz);
    memcpy(o->ids, ids, (size_t)nx * ny * nz);
    const size_t sx = (size_t)(nx - 1) * ny * nz, sy = (size_t)nx * (ny - 1) * nz, sz = (size_t)nx * ny * (nz - 1);
    o->ux = zalloc(sx); o->uxn = zalloc(sx); o->uxo = zalloc(sx);
    o->uy = zalloc(sy); o->uyn = zalloc(sy); o->uyo = zalloc(sy);
    o->uz = zalloc(sz); o->uzn = zalloc(sz); o->uzo = zalloc(sz);
    o->T1 = zalloc((size_t)nx * ny * nz); o->T2 = zalloc((size_t)nx * ny * nz); o->T3 = zalloc((size_t)nx * ny * nz);
    o->T4 = zalloc((size_t)nx * (ny - 1) * (nz - 1));
    o->T5 = zalloc((size_t)(nx - 1) * ny * (nz - 1));
    o->T6 = zalloc((size_t)(nx - 1) * (ny - 1) * nz);
    return o;
}

void oc_destroy(void *h) {
    oc_t *o = (oc_t *)h;
    double *p[] = {o->fdx, o->fdy, o->fdz, o->sdx, o->sdy, o->sdz, o->tab, o->rho, o->ux, o->uy, o->uz, o->uxn, o->uyn,
                   o->uzn, o->uxo, o->uyo, o->uzo, o->T1, o->T2, o->T3, o->T4, o->T5, o->T6};
    for (size_t q = 0; q < sizeof p / sizeof *p; ++q) free(p[q]);
    free(o->ids);
    free(o);
}

/* which: 0 ux 1 uy 2 uz 3 ux_old 4 uy_old 5 uz_old 6..11 T1..T6 12 ux_new 13 uy_new 14 uz_new */
double *oc_array(void *h, int which) {
    oc_t *o = (oc_t *)h;
    double *a[] = {o->ux, o->uy, o->uz, o->uxo, o->uyo, o->uzo, o->T1, o->T2, o->T3, o->T4, o->T5, o->T6, o->uxn, o->uyn, o->uzn};
    return a[which];
}

static void update_T(oc_t *o) { /* base_solver.py:323-372 */
    const int nx = o->nx, ny = o->ny, nz = o->nz;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx - 2; ++i)
        for (int j = 1; j <= ny - 2; ++j)
            for (int k = 1; k <= nz - 2; ++k) {
                const double dxx = UX(o->ux, i, j, k) - UX(o->ux, i - 1, j, k);
                const double dyy = UY(o->uy, i, j, k) - UY(o->uy, i, j - 1, k);
                const double dzz = UZ(o->uz, i, j, k) - UZ(o->uz, i, j, k - 1);
                const double sx = o->sdx[i - 1], sy = o->sdy[j - 1], sz = o->sdz[k - 1];
                TN(o->T1, i, j, k) = CC(i, j, k, 0, 0) * dxx / sx + CC(i, j, k, 0, 1) * dyy / sy + CC(i, j, k, 0, 2) * dzz / sz;
                TN(o->T2, i, j, k) = CC(i, j, k, 1, 0) * dxx / sx + CC(i, j, k, 1, 1) * dyy / sy + CC(i, j, k, 1, 2) * dzz / sz;
                TN(o->T3, i, j, k) = CC(i, j, k, 2, 0) * dxx / sx + CC(i, j, k, 2, 1) * dyy / sy + CC(i, j, k, 2, 2) * dzz / sz;
            }
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx - 2; ++i)
        for (int j = 0; j <= ny - 2; ++j)
            for (int k = 0; k <= nz - 2; ++k)
                T4_(i, j, k) = CC(i, j + 1, k + 1, 3, 3) * ((UY(o->uy, i, j, k + 1) - UY(o->uy, i, j, k)) / o->fdz[k] +
                                                            (UZ(o->uz, i, j + 1, k) - UZ(o->uz, i, j, k)) / o->fdy[j]);
#pragma omp parallel for schedule(static)
    for (int i = 0; i <= nx - 2; ++i)
        for (int j = 1; j <= ny - 2; ++j)
            for (int k = 0; k <= nz - 2; ++k)
                T5_(i, j, k) = CC(i + 1, j, k + 1, 4, 4) * ((UX(o->ux, i, j, k + 1) - UX(o->ux, i, j, k)) / o->fdz[k] +
                                                            (UZ(o->uz, i + 1, j, k) - UZ(o->uz, i, j, k)) / o->fdx[i]);
#pragma omp parallel for schedule(static)
    for (int i = 0; i <= nx - 2; ++i)
        for (int j = 0; j <= ny - 2; ++j)
            for (int k = 1; k <= nz - 2; ++k)
                T6_(i, j, k) = CC(i + 1, j + 1, k, 5, 5) * ((UX(o->ux, i, j + 1, k) - UX(o->ux, i, j, k)) / o->fdy[j] +
                                                            (UY(o->uy, i + 1, j, k) - UY(o->uy, i, j, k)) / o->fdx[i]);
}

static void apply_T_tfbc(oc_t *o) { /* base_solver.py:402-433: first-element spacings, "wrong" axes kept */
    const int nx = o->nx, ny = o->ny, nz = o->nz;
    const double sdx0 = o->sdx[0], sdy0 = o->sdy[0], sdz0 = o->sdz[0], fdx0 = o->fdx[0], fdy0 = o->fdy[0], fdz0 = o->fdz[0];
    for (int i = 1; i <= nx - 2; ++i)
        for (int j = 1; j <= ny - 2; ++j) {
            const double dxx = UX(o->ux, i, j, 0) - UX(o->ux, i - 1, j, 0);
            const double dyy = UY(o->uy, i, j, 0) - UY(o->uy, i, j - 1, 0);
            const double dzz = UZ(o->uz, i, j, 0) - 0;
            TN(o->T1, i, j, 0) = CC(i, j, 0, 0, 0) * dxx / sdx0 + CC(i, j, 0, 0, 1) * dyy / sdy0 + CC(i, j, 0, 0, 2) * dzz / sdz0;
            TN(o->T2, i, j, 0) = CC(i, j, 0, 1, 0) * dxx / sdx0 + CC(i, j, 0, 1, 1) * dyy / sdy0 + CC(i, j, 0, 1, 2) * dzz / sdz0;
            TN(o->T3, i, j, 0) = 0;
        }
    for (int i = 1; i <= nx - 2; ++i)
        for (int j = 0; j <= ny - 2; ++j)
            T4_(i, j, 0) = CC(i, j + 1, 0, 3, 3) * ((UY(o->uy, i, j, 1) - UY(o->uy, i, j, 0)) / fdy0 +
                                                    (UZ(o->uz, i, j + 1, 0) - UZ(o->uz, i, j, 0)) / fdz0);
    for (int i = 0; i <= nx - 2; ++i)
        for (int j = 1; j <= ny - 2; ++j)
            T5_(i, j, 0) = CC(i + 1, j, 0, 4, 4) * ((UX(o->ux, i, j, 1) - UX(o->ux, i, j, 0)) / fdx0 +
                                                    (UZ(o->uz, i + 1, j, 0) - UZ(o->uz, i, j, 0)) / fdz0);
    for (int i = 0; i <= nx - 2; ++i)
        for (int j = 0; j <= ny - 2; ++j)
            T6_(i, j, 0) = CC(i + 1, j + 1, 0, 5, 5) * ((UX(o->ux, i, j + 1, 0) - UX(o->ux, i, j, 0)) / fdx0 +
                                                        (UY(o->uy, i + 1, j, 0) - UY(o->uy, i, j, 0)) / fdz0);
}

static void update_u(oc_t *o) { /* base_solver.py:435-463 */
    const int nx = o->nx, ny = o->ny, nz = o->nz;
    const double d2 = o->d2;
#pragma omp parallel for schedule(static)
    for (int i = 0; i <= nx - 2; ++i)
        for (int j = 1; j <= ny - 2; ++j)
            for (int k = 1; k <= nz - 2; ++k)
                UX(o->uxn, i, j, k) = 2 * UX(o->ux, i, j, k) - UX(o->uxo, i, j, k) +
                    (d2 / PP(i + 1, j, k)) * ((TN(o->T1, i + 1, j, k) - TN(o->T1, i, j, k)) / o->fdx[i] +
                                              (T6_(i, j, k) - T6_(i, j - 1, k)) / o->sdy[j - 1] +
                                              (T5_(i, j, k) - T5_(i, j, k - 1)) / o->sdz[k - 1]);
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx - 2; ++i)
        for (int j = 0; j <= ny - 2; ++j)
            for (int k = 1; k <= nz - 2; ++k)
                UY(o->uyn, i, j, k) = 2 * UY(o->uy, i, j, k) - UY(o->uyo, i, j, k) +
                    (d2 / PP(i, j + 1, k)) * ((T6_(i, j, k) - T6_(i - 1, j, k)) / o->sdx[i - 1] +
                                              (TN(o->T2, i, j + 1, k) - TN(o->T2, i, j, k)) / o->fdy[j] +
                                              (T4_(i, j, k) - T4_(i, j, k - 1)) / o->sdz[k - 1]);
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx - 2; ++i)
        for (int j = 1; j <= ny - 2; ++j)
            for (int k = 0; k <= nz - 2; ++k)
                UZ(o->uzn, i, j, k) = 2 * UZ(o->uz, i, j, k) - UZ(o->uzo, i, j, k) +
                    (d2 / PP(i, j, k + 1)) * ((T5_(i, j, k) - T5_(i - 1, j, k)) / o->sdx[i - 1] +
                                              (T4_(i, j, k) - T4_(i, j - 1, k)) / o->sdy[j - 1] +
                                              (TN(o->T3, i, j, k + 1) - TN(o->T3, i, j, k)) / o->fdz[k]);
}

static void apply_u_tfbc(oc_t *o) { /* base_solver.py:488-517 */
    const int nx = o->nx, ny = o->ny, nz = o->nz;
    const double d2 = o->d2;
    const double sdx0 = o->sdx[0], sdy0 = o->sdy[0], sdz0 = o->sdz[0], fdx0 = o->fdx[0], fdy0 = o->fdy[0], fdz0 = o->fdz[0];
    for (int i = 0; i <= nx - 2; ++i)
        for (int j = 1; j <= ny - 2; ++j)
            UX(o->uxn, i, j, 0) = 2 * UX(o->ux, i, j, 0) - UX(o->uxo, i, j, 0) +
                (d2 / PP(i + 1, j, 0)) * ((TN(o->T1, i + 1, j, 0) - TN(o->T1, i, j, 0)) / fdx0 +
                                          (T6_(i, j, 0) - T6_(i, j - 1, 0)) / sdy0 + (T5_(i, j, 0) - 0) / sdz0);
    for (int i = 1; i <= nx - 2; ++i)
        for (int j = 0; j <= ny - 2; ++j)
            UY(o->uyn, i, j, 0) = 2 * UY(o->uy, i, j, 0) - UY(o->uyo, i, j, 0) +
                (d2 / PP(i, j + 1, 0)) * ((T6_(i, j, 0) - T6_(i - 1, j, 0)) / sdx0 +
                                          (TN(o->T2, i, j + 1, 0) - TN(o->T2, i, j, 0)) / fdy0 + (T4_(i, j, 0) - 0) / sdz0);
    for (int i = 1; i <= nx - 2; ++i)
        for (int j = 1; j <= ny - 2; ++j) /* precedence slip kept: T3[..,1] is not divided; rho at k = 0 */
            UZ(o->uzn, i, j, 0) = 2 * UZ(o->uz, i, j, 0) - UZ(o->uzo, i, j, 0) +
                (d2 / PP(i, j, 0)) * ((T5_(i, j, 0) - T5_(i - 1, j, 0)) / sdx0 + (T4_(i, j, 0) - T4_(i, j - 1, 0)) / sdy0 +
                                      TN(o->T3, i, j, 1) - TN(o->T3, i, j, 0) / fdz0);
}

static void apply_u_abc(oc_t *o) { /* base_solver.py:519-554; face order x, y0, y1, z */
    const int nx = o->nx, ny = o->ny, nz = o->nz;
    const double dt = o->dt;
    const double vl = sqrt(CC(0, 0, 0, 0, 0) / PP(0, 0, 0)), vt = sqrt(CC(0, 0, 0, 3, 3) / PP(0, 0, 0));
#define KAP(v, d) (((v) * dt - (d)) / ((v) * dt + (d)))
    const double ctx = KAP(vt, o->fdx[nx - 2]), clx = KAP(vl, o->sdx[nx - 3]);
    const double cty0 = KAP(vt, o->fdy[0]), cly0 = KAP(vl, o->sdy[0]), cty1 = KAP(vt, o->fdy[ny - 2]), cly1 = KAP(vl, o->sdy[ny - 3]);
    const double ctz = KAP(vt, o->fdz[nz - 2]), clz = KAP(vl, o->sdz[nz - 3]);
    for (int j = 0; j < ny; ++j)
        for (int k = 0; k < nz; ++k)
            UX(o->uxn, nx - 2, j, k) = UX(o->ux, nx - 3, j, k) + clx * (UX(o->uxn, nx - 3, j, k) - UX(o->ux, nx - 2, j, k));
    for (int j = 0; j < ny - 1; ++j)
        for (int k = 0; k < nz; ++k)
            UY(o->uyn, nx - 1, j, k) = UY(o->uy, nx - 2, j, k) + ctx * (UY(o->uyn, nx - 2, j, k) - UY(o->uy, nx - 1, j, k));
    for (int j = 0; j < ny; ++j)
        for (int k = 0; k < nz - 1; ++k)
            UZ(o->uzn, nx - 1, j, k) = UZ(o->uz, nx - 2, j, k) + ctx * (UZ(o->uzn, nx - 2, j, k) - UZ(o->uz, nx - 1, j, k));
    for (int f = 0; f < 2; ++f) { /* y = 0, then y = -1 */
        const double ct = f ? cty1 : cty0, cl = f ? cly1 : cly0;
        for (int i = 0; i < nx - 1; ++i)
            for (int k = 0; k < nz; ++k) {
                const int a = f ? ny - 1 : 0, b = f ? ny - 2 : 1;
                UX(o->uxn, i, a, k) = UX(o->ux, i, b, k) + ct * (UX(o->uxn, i, b, k) - UX(o->ux, i, a, k));
            }
        for (int i = 0; i < nx; ++i)
            for (int k = 0; k < nz; ++k) {
                const int a = f ? ny - 2 : 0, b = f ? ny - 3 : 1;
                UY(o->uyn, i, a, k) = UY(o->uy, i, b, k) + cl * (UY(o->uyn, i, b, k) - UY(o->uy, i, a, k));
            }
        for (int i = 0; i < nx; ++i)
            for (int k = 0; k < nz - 1; ++k) {
                const int a = f ? ny - 1 : 0, b = f ? ny - 2 : 1;
                UZ(o->uzn, i, a, k) = UZ(o->uz, i, b, k) + ct * (UZ(o->uzn, i, b, k) - UZ(o->uz, i, a, k));
            }
    }
    for (int i = 0; i < nx - 1; ++i)
        for (int j = 0; j < ny; ++j)
            UX(o->uxn, i, j, nz - 1) = UX(o->ux, i, j, nz - 2) + ctz * (UX(o->uxn, i, j, nz - 2) - UX(o->ux, i, j, nz - 1));
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny - 1; ++j)
            UY(o->uyn, i, j, nz - 1) = UY(o->uy, i, j, nz - 2) + ctz * (UY(o->uyn, i, j, nz - 2) - UY(o->uy, i, j, nz - 1));
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            UZ(o->uzn, i, j, nz - 2) = UZ(o->uz, i, j, nz - 3) + clz * (UZ(o->uzn, i, j, nz - 3) - UZ(o->uz, i, j, nz - 2));
}

static void time_step(oc_t *o) { /* base_solver.py:556-571: copies, afterwards u_new == u */
    const int nx = o->nx, ny = o->ny, nz = o->nz;
    const size_t sx = (size_t)(nx - 1) * ny * nz, sy = (size_t)nx * (ny - 1) * nz, sz = (size_t)nx * ny * (nz - 1);
    memcpy(o->uxo, o->ux, sx * 8); memcpy(o->uyo, o->uy, sy * 8); memcpy(o->uzo, o->uz, sz * 8);
    memcpy(o->ux, o->uxn, sx * 8); memcpy(o->uy, o->uyn, sy * 8); memcpy(o->uz, o->uzn, sz * 8);
}

void oc_step(void *h, double w) { /* base_solver.py:251-256 */
    oc_t *o = (oc_t *)h;
    const int ny = o->ny, nz = o->nz;
    for (int j = 0; j < ny; ++j) UZ(o->uz, 0, j, 0) = w;
    update_T(o);
    apply_T_tfbc(o);
    update_u(o);
    apply_u_tfbc(o);
    apply_u_abc(o);
    time_step(o);
}

void oc_run(void *h, const double *w, int nsteps) {
    for (int s = 0; s < nsteps; ++s) oc_step(h, w[s]);
}

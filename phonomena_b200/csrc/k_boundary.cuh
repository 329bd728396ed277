// k_boundary.cuh -- source injection, absorbing faces, layout conversion, material codes,
// device-side inclusion fill, surface recorder.
#pragma once
#include "fd_common.cuh"

namespace phb {

// ---------------------------------------------------------------------------------------
// Source: uz[0, :, 0] = w(tt) on the CURRENT field before the stress update
// (base_solver.py:251).  The overwritten line is saved first: the reference's u_new keeps
// the pre-source value there (App. B #9) and the step kernel re-emits it.
// ---------------------------------------------------------------------------------------
template <class T>
__global__ void k_source(Geo<T> g, T *uz_cur, T *line_save, const double *w, long long tt) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= g.ny) return;
    const long long c = g.idx(0, j, 0);
    line_save[j] = uz_cur[c];
    uz_cur[c] = (T)w[tt];
}

// ---------------------------------------------------------------------------------------
// First-order Mur faces (base_solver.py:539-554), App. A.5:
//   q_new[face] = q[inner] + c * (q_new[inner] - q[face])
// Order matters on shared edges: x = -1, then y = 0 and y = -1, then z = -1; the host
// launches them in that order on one stream.
// Per component: extent (ex, ey, ez) = the reference array shape; coefficient.
// ---------------------------------------------------------------------------------------
template <class A>
__device__ __forceinline__ typename A::T mur(typename A::T q_inner, typename A::T qn_inner, typename A::T q_face,
                                             typename A::T c) {
    return A::add(q_inner, A::mul(c, A::sub(qn_inner, q_face)));
}

template <class T>
struct AbcArgs {
    Geo<T> g;
    Fld<T> cur, nw;
    T clx, ctx, cly0, cty0, cly1, cty1, clz, ctz;
    int i_begin, i_end;   // owned planes the y/z faces cover
};

// x = -1 face (only the rank that owns the last planes).  threads over (j, k).
template <class A>
__global__ void k_abc_x(AbcArgs<typename A::T> p) {
    using T = typename A::T;
    const Geo<T> &g = p.g;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (k >= g.nz || j >= g.ny) return;
    {   // ux: planes nx-2 (face), nx-3 (inner)
        const long long f = g.idx(g.nx - 2, j, k), n = g.idx(g.nx - 3, j, k);
        p.nw.ux[f] = mur<A>(p.cur.ux[n], p.nw.ux[n], p.cur.ux[f], p.clx);
    }
    const long long f = g.idx(g.nx - 1, j, k), n = g.idx(g.nx - 2, j, k);
    if (j < g.ny - 1) p.nw.uy[f] = mur<A>(p.cur.uy[n], p.nw.uy[n], p.cur.uy[f], p.ctx);
    if (k < g.nz - 1) p.nw.uz[f] = mur<A>(p.cur.uz[n], p.nw.uz[n], p.cur.uz[f], p.ctx);
}

// y = 0 (blockIdx.z == 0) and y = -1 (blockIdx.z == 1) faces.  threads over (i, k).
template <class A>
__global__ void k_abc_y(AbcArgs<typename A::T> p) {
    using T = typename A::T;
    const Geo<T> &g = p.g;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = p.i_begin + blockIdx.y * blockDim.y + threadIdx.y;
    if (k >= g.nz || i >= p.i_end) return;
    const bool hi = (blockIdx.z == 1);
    const T ct = hi ? p.cty1 : p.cty0, cl = hi ? p.cly1 : p.cly0;
    if (i < g.nx - 1) {                       // ux: rows 0|ny-1, inner 1|ny-2
        const int jf = hi ? g.ny - 1 : 0, jn = hi ? g.ny - 2 : 1;
        const long long f = g.idx(i, jf, k), n = g.idx(i, jn, k);
        p.nw.ux[f] = mur<A>(p.cur.ux[n], p.nw.ux[n], p.cur.ux[f], ct);
    }
    {                                         // uy: rows 0|ny-2, inner 1|ny-3
        const int jf = hi ? g.ny - 2 : 0, jn = hi ? g.ny - 3 : 1;
        const long long f = g.idx(i, jf, k), n = g.idx(i, jn, k);
        p.nw.uy[f] = mur<A>(p.cur.uy[n], p.nw.uy[n], p.cur.uy[f], cl);
    }
    if (k < g.nz - 1) {                       // uz: rows 0|ny-1
        const int jf = hi ? g.ny - 1 : 0, jn = hi ? g.ny - 2 : 1;
        const long long f = g.idx(i, jf, k), n = g.idx(i, jn, k);
        p.nw.uz[f] = mur<A>(p.cur.uz[n], p.nw.uz[n], p.cur.uz[f], ct);
    }
}

// z = -1 face.  threads over (i, j); j on threadIdx.x (stride nzp: 2 adjacent elements each).
template <class A>
__global__ void k_abc_z(AbcArgs<typename A::T> p) {
    using T = typename A::T;
    const Geo<T> &g = p.g;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = p.i_begin + blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= g.ny || i >= p.i_end) return;
    if (i < g.nx - 1) {
        const long long f = g.idx(i, j, g.nz - 1), n = f - 1;
        p.nw.ux[f] = mur<A>(p.cur.ux[n], p.nw.ux[n], p.cur.ux[f], p.ctz);
    }
    if (j < g.ny - 1) {
        const long long f = g.idx(i, j, g.nz - 1), n = f - 1;
        p.nw.uy[f] = mur<A>(p.cur.uy[n], p.nw.uy[n], p.cur.uy[f], p.ctz);
    }
    {
        const long long f = g.idx(i, j, g.nz - 2), n = f - 1;
        p.nw.uz[f] = mur<A>(p.cur.uz[n], p.nw.uz[n], p.cur.uz[f], p.clz);
    }
}

// ---------------------------------------------------------------------------------------
// Layout conversion between host arrays (float64, reference shapes, plane-major) and the
// padded device box.  src/dst host-shaped array: (np, ey, ez); device plane l0 + p.
// ---------------------------------------------------------------------------------------
template <class T>
__global__ void k_scatter(const double *src, T *dst, int np, int ey, int ez, int l0, int ny, int nzp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int p = blockIdx.z;
    if (k >= ez || j >= ey || p >= np) return;
    dst[((long long)(l0 + p) * ny + j) * nzp + k] = (T)src[((long long)p * ey + j) * ez + k];
}
template <class T>
__global__ void k_gather(const T *src, double *dst, int np, int ey, int ez, int l0, int ny, int nzp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int p = blockIdx.z;
    if (k >= ez || j >= ey || p >= np) return;
    dst[((long long)p * ey + j) * ez + k] = (double)src[((long long)(l0 + p) * ny + j) * nzp + k];
}

// ---------------------------------------------------------------------------------------
// Material stencil codes from the raw id box.  ids: planes [ib, ie) of the global grid,
// unpadded (ny, nz).  One code per cell of every local plane (ghosts included); indices are
// clamped -- a clamped lookup only ever feeds a stress the range masks zero out.
// ---------------------------------------------------------------------------------------
template <class CodeT, int B>
__global__ void k_build_codes(const uint8_t *ids, int ib, int ie, CodeT *code, int nx, int ny, int nz, int nzp,
                              int x0, int nxl) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int l = blockIdx.z;                    // local plane 0 .. nxl+1
    if (k >= nzp || j >= ny || l >= nxl + 2) return;
    const int i = x0 - 1 + l;
    auto id = [&](int ii, int jj, int kk) -> unsigned {
        ii = min(max(ii, ib), ie - 1);
        jj = min(max(jj, 0), ny - 1);
        kk = min(max(kk, 0), nz - 1);
        return ids[((long long)(ii - ib) * ny + jj) * nz + kk];
    };
    const int kz = (k == 0) ? 0 : k + 1;         // the k = 0 plane reads its own z level (App. A.2/A.4)
    unsigned c = 0;
    c |= id(i, j, k) << (F_NODE * B);
    c |= id(i, j + 1, kz) << (F_T4 * B);
    c |= id(i + 1, j, kz) << (F_T5 * B);
    c |= id(i + 1, j + 1, k) << (F_T6 * B);
    c |= id(i + 1, j, k) << (F_RX * B);
    c |= id(i, j + 1, k) << (F_RY * B);
    c |= id(i, j, kz) << (F_RZ * B);
    code[((long long)l * ny + j) * nzp + k] = (CodeT)c;
}

// ---------------------------------------------------------------------------------------
// Device-side inclusion fill (grid.py:158-175 + material.py:55-63), App. A.7.
// Pass 1: per (i, j) column, a bit per z-profile group: is the column inside any cylinder of
// the group?  R = sqrt((x-tx)^2 + (y-ty)^2) < r in float64 with the float32 target fields
// widened exactly, separately rounded ops as NumPy evaluates them.
// Pass 2: id(i,j,k) = (colbits & zbits[k]) != 0.
// ---------------------------------------------------------------------------------------
__global__ void k_incl_columns(const double *x, const double *y, int ib, int ie, int ny, const float *tg,
                               const int *group, int n, unsigned *colbits) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = ib + blockIdx.y;
    if (j >= ny || i >= ie) return;
    const double xi = x[i], yj = y[j];
    unsigned bits = 0;
    for (int t = 0; t < n; ++t) {
        const double dx = __dsub_rn(xi, (double)tg[4 * t + 0]);
        const double dy = __dsub_rn(yj, (double)tg[4 * t + 1]);
        const double R = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        if (R < (double)tg[4 * t + 3]) bits |= 1u << group[t];
    }
    colbits[(long long)(i - ib) * ny + j] = bits;
}
__global__ void k_incl_fill(const unsigned *colbits, const unsigned *zbits, uint8_t *ids, int np, int ny, int nz) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int p = blockIdx.z;
    if (k >= nz || j >= ny || p >= np) return;
    ids[((long long)p * ny + j) * nz + k] = (colbits[(long long)p * ny + j] & zbits[k]) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------
// Surface recorder: k = 0 plane of the selected components, converted to float64, written
// straight into a pinned host ring slot (zero-copy store over PCIe).  Slot layout: for each
// recorded component in order ux, uy, uz a (planes, ey) array.
// ---------------------------------------------------------------------------------------
template <class T>
__global__ void k_record(Geo<T> g, Fld<T> u, int mask, double *slot, int npx, int npyz) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;            // owned plane index
    if (j >= g.ny) return;
    const int i = g.x0 + p;
    double *o = slot;
    if (mask & 1) {
        if (p < npx) o[(long long)p * g.ny + j] = (double)u.ux[g.idx(i, j, 0)];
        o += (long long)npx * g.ny;
    }
    if (mask & 2) {
        if (j < g.ny - 1) o[(long long)p * (g.ny - 1) + j] = (double)u.uy[g.idx(i, j, 0)];
        o += (long long)npyz * (g.ny - 1);
    }
    if (mask & 4) o[(long long)p * g.ny + j] = (double)u.uz[g.idx(i, j, 0)];
}

}  // namespace phb

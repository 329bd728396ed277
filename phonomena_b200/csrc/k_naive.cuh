// k_naive.cuh -- straightforward fused step kernel: one thread per cell, every stress the
// cell needs is re-evaluated from global memory (L1/L2 absorb the re-reads).  It is the
// on-device specification: small, obviously equal to SURVEY App. A, validated bit-for-bit
// against the oracle; the marching kernel (k_march.cuh) is validated against both.
#pragma once
#include "fd_common.cuh"

namespace phb {

template <class T>
struct StepArgs {
    Geo<T> g;
    Fld<T> cur, old, nw;
    const T *line_save;   // pre-source uz(0, j, 0) of `cur` (App. B #9), or nullptr
    int i_begin, i_end;   // global planes to update: [i_begin, i_end)
    // fused halo push (k_march only): where the first / last owned plane of u_new also goes -- the matching ghost
    // plane of the left / right neighbour's buffer, mapped through CUDA IPC (NVLink peer stores); null = none
    T *push_lo[3], *push_hi[3];
    int edge_b;           // >= 0: 'edge launch' of two chunks of (i_end - i_begin) planes, starting at i_begin and at edge_b
    // fused z = -1 absorbing face (k_march only, not in COMP mode): the block that owns k = nz-1 applies the Mur
    // formula to its own results before storing them (the face points and their inner neighbours sit in one
    // lane); the host then re-applies the face only where the x / y faces change its inputs (k_abc_z, edges only)
    int zface;
    T zf_ct, zf_cl;
    int ztile0;           // k_march only: first z-tile of this launch (the step may be split into a launch for the z-tiles
                          // without the face and one for the tile that owns it, see phb200.cu physics())
};

// u_new for one cell from generic stress evaluations.  Writes only entries the reference's
// physics writes (App. A.3/A.4 ranges) plus the i = 0 copy that keeps `u_new == u` where
// nothing is ever written (App. B #9).
// EV: the stress evaluator (Eval, or EvalBloch for a Bloch-periodic pair of field sets)
template <class A, class EV>
__device__ __forceinline__ void naive_cell_ev(const StepArgs<typename A::T> &p, const EV &ev, int i, int j, int k) {
    using T = typename A::T;
    const Geo<T> &g = p.g;
    const bool k0 = (k == 0);
    const long long c = g.idx(i, j, k);
    if (k > g.nz - 2) return;   // k = nz-1: ABC face (ux, uy) / non-existent (uz)

    // ---- ux: 0<=i<=nx-2, 1<=j<=ny-2 -------------------------------------------------------
    if (i <= g.nx - 2 && j >= 1 && j <= g.ny - 2) {
        T a1, a2, a3, b1, b2, b3;
        ev.normal(i + 1, j, k, a1, a2, a3);
        ev.normal(i, j, k, b1, b2, b3);
        const T dA = A::sub(a1, b1);
        const T dB = A::sub(ev.t6(i, j, k), ev.t6(i, j - 1, k));
        const T dC = k0 ? ev.t5(i, j, 0) : A::sub(ev.t5(i, j, k), ev.t5(i, j, k - 1));
        const T acc = A::add(A::add(A::scl(dA, k0 ? g.fdx0 : g.fdx[i]), A::scl(dB, k0 ? g.sdy0 : g.sdy[j - 1])),
                             A::scl(dC, k0 ? g.sdz0 : g.sdz[k - 1]));
        const T rinv = ev.coef(i, j, k, CLS_RX);
        if constexpr (A::COMP) {
            const T dn = p.old.ux[c] + rinv * acc;      // p.old holds delta
            p.old.ux[c] = dn;
            p.nw.ux[c] = p.cur.ux[c] + dn;
        } else {
            p.nw.ux[c] = advance<A>(p.cur.ux[c], p.old.ux[c], rinv, acc);
        }
    }
    // ---- uy: 1<=i<=nx-2, 0<=j<=ny-2 -------------------------------------------------------
    if (i >= 1 && i <= g.nx - 2 && j <= g.ny - 2) {
        T a1, a2, a3, b1, b2, b3;
        ev.normal(i, j + 1, k, a1, a2, a3);
        ev.normal(i, j, k, b1, b2, b3);
        const T dA = A::sub(ev.t6(i, j, k), ev.t6(i - 1, j, k));
        const T dB = A::sub(a2, b2);
        const T dC = k0 ? ev.t4(i, j, 0) : A::sub(ev.t4(i, j, k), ev.t4(i, j, k - 1));
        const T acc = A::add(A::add(A::scl(dA, k0 ? g.sdx0 : g.sdx[i - 1]), A::scl(dB, k0 ? g.fdy0 : g.fdy[j])),
                             A::scl(dC, k0 ? g.sdz0 : g.sdz[k - 1]));
        const T rinv = ev.coef(i, j, k, CLS_RY);
        if constexpr (A::COMP) {
            const T dn = p.old.uy[c] + rinv * acc;      // p.old holds delta
            p.old.uy[c] = dn;
            p.nw.uy[c] = p.cur.uy[c] + dn;
        } else {
            p.nw.uy[c] = advance<A>(p.cur.uy[c], p.old.uy[c], rinv, acc);
        }
    }
    // ---- uz: 1<=i<=nx-2, 1<=j<=ny-2 -------------------------------------------------------
    if (i >= 1 && i <= g.nx - 2 && j >= 1 && j <= g.ny - 2) {
        T a1, a2, a3, b1, b2, b3;
        ev.normal(i, j, k + 1, a1, a2, a3);
        ev.normal(i, j, k, b1, b2, b3);
        const T dA = A::sub(ev.t5(i, j, k), ev.t5(i - 1, j, k));
        const T dB = A::sub(ev.t4(i, j, k), ev.t4(i, j - 1, k));
        const T sA = A::scl(dA, k0 ? g.sdx0 : g.sdx[i - 1]);
        const T sB = A::scl(dB, k0 ? g.sdy0 : g.sdy[j - 1]);
        T acc;
        if (k0) {
            // "+ T3[..,1] - T3[..,0]/fdz[0]" with T3[..,0] == 0 (base_solver.py:516, App. B #3)
            acc = A::sub(A::add(A::add(sA, sB), a3), A::scl(b3, g.fdz0));
        } else {
            acc = A::add(A::add(sA, sB), A::scl(A::sub(a3, b3), g.fdz[k]));
        }
        const T rinv = ev.coef(i, j, k, CLS_RZ);
        if constexpr (A::COMP) {
            const T dn = p.old.uz[c] + rinv * acc;      // p.old holds delta
            p.old.uz[c] = dn;
            p.nw.uz[c] = p.cur.uz[c] + dn;
        } else {
            p.nw.uz[c] = advance<A>(p.cur.uz[c], p.old.uz[c], rinv, acc);
        }
    }
    // ---- i = 0: uy, uz are never written by the physics; keep u_new == u (App. B #9) ------
    if (i == 0) {
        if (j <= g.ny - 2) p.nw.uy[c] = p.cur.uy[c];
        p.nw.uz[c] = (k0 && p.line_save) ? p.line_save[j] : p.cur.uz[c];
        if constexpr (A::COMP) {      // keep delta = u_new - u here too (only matters for reading u_old back)
            if (j <= g.ny - 2) p.old.uy[c] = (T)0;
            p.old.uz[c] = p.nw.uz[c] - p.cur.uz[c];
        }
    }
}

template <class A, class M>
__device__ __forceinline__ void naive_cell(const StepArgs<typename A::T> &p, const M &m, int i, int j, int k, bool pbc = false) {
    Eval<A, M> ev(p.g, p.cur, m, pbc);
    naive_cell_ev<A>(p, ev, i, j, k);
}

// grid: x -> k tiles, y -> j tiles, z -> plane (i_begin + blockIdx.z)
template <class A, class M>
__global__ void __launch_bounds__(256) k_step_naive(StepArgs<typename A::T> p, M m) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = p.i_begin + blockIdx.z;
    if (k >= p.g.nz || j >= p.g.ny || i >= p.i_end) return;
    naive_cell<A, M>(p, m, i, j, k);
}

// Periodic y boundaries (zero Bloch phase): the reference's archived stubs apply_T_pbc / apply_u_pbc
// (base_solver.py:383-400, 475-486), each applied after the corresponding traction-free update, in place of the
// y faces of the Mur ABC.  The stress copies only reach three rows of u_new -- ux and uz on row ny-2 (through T6 / T4
// of that row = those of row 1) and uy on row 0 (through T2 of row 0 = that of row ny-2) -- so the step kernels run
// unchanged and this kernel recomputes those rows from the current field with the wrapped stresses (same formula
// functions, so EXACT arithmetic stays bit-identical), then makes the displacement copies
//   ux_new[:,0,:] = ux_new[:,-2,:];  uz_new[:,0,:] = uz_new[:,-2,:];  uy_new[:,-1,:] = uy_new[:,1,:]
// and keeps u_new == u on the rows nothing writes in this mode (ux, uz on row ny-1).
// threads over (k, i); planes [i_begin, i_end).
template <class A, class M>
__global__ void __launch_bounds__(256) k_pbc_y(StepArgs<typename A::T> p, M m) {
    const Geo<typename A::T> &g = p.g;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = p.i_begin + blockIdx.y * blockDim.y + threadIdx.y;
    if (k >= g.nz || i >= p.i_end) return;
    naive_cell<A, M>(p, m, i, g.ny - 2, k, true);
    naive_cell<A, M>(p, m, i, 0, k, true);
    const long long r0 = g.idx(i, 0, k), r1 = g.idx(i, 1, k), rm = g.idx(i, g.ny - 2, k), rl = g.idx(i, g.ny - 1, k);
    if (i <= g.nx - 2) {
        p.nw.ux[r0] = p.nw.ux[rm];
        p.nw.ux[rl] = p.cur.ux[rl];
    }
    if (k <= g.nz - 2) {
        p.nw.uz[r0] = p.nw.uz[rm];
        // (the source line of the current field is not part of u_new: the pre-source value is, App. B #9)
        p.nw.uz[rl] = (i == 0 && k == 0 && p.line_save) ? p.line_save[g.ny - 1] : p.cur.uz[rl];
    }
    p.nw.uy[rm] = p.nw.uy[r1];
}

// ---------------------------------------------------------------------------------------
// Bloch-periodic y boundaries with a phase: u(y + L) = u(y) e^{i phi}, L = ny - 2 rows.  The field is complex = two real
// field sets (a real stencil advances them independently); they couple only where the periodic stubs copy: a copy from
// row ny-2 to row 0 (node rows: T1, T2, T3, T5, ux, uz) multiplies by e^{-i phi}, a copy from row 1 to the last row
// (staggered rows: T4, T6, uy) by e^{+i phi}.  phi = 0 is k_pbc_y, i.e. the reference's archived stubs.
// EvalBloch: evaluator of the set `a` whose wrapped stresses mix in the other set `b`:
//   node rows:      cph * T_a + sph * T_b        staggered rows:  cph * T_a - sph * T_b
// (real part: cph = cos phi, sph = sin phi, b = imaginary set; imaginary part: sph = -sin phi, b = real set).
// ---------------------------------------------------------------------------------------
template <class A, class M>
struct EvalBloch {
    using T = typename A::T;
    const Geo<T> &g;
    Eval<A, M> a, b;
    T cph, sph;
    __device__ EvalBloch(const Geo<T> &g_, const Fld<T> &ua, const Fld<T> &ub, const M &m, T c_, T s_)
        : g(g_), a(g_, ua, m, false), b(g_, ub, m, false), cph(c_), sph(s_) {}
    __device__ __forceinline__ T mix(T x, T y, T s) const { return A::add(A::mul(cph, x), A::mul(s, y)); }
    __device__ __forceinline__ T coef(int i, int j, int k, int e) const { return a.coef(i, j, k, e); }
    __device__ __forceinline__ void normal(int i, int j, int k, T &t1, T &t2, T &t3) const {
        if (j != 0) { a.normal(i, j, k, t1, t2, t3); return; }
        T x1, x2, x3, y1, y2, y3;
        a.normal(i, g.ny - 2, k, x1, x2, x3);
        b.normal(i, g.ny - 2, k, y1, y2, y3);
        t1 = mix(x1, y1, sph); t2 = mix(x2, y2, sph); t3 = mix(x3, y3, sph);
    }
    __device__ __forceinline__ T t5(int i, int j, int k) const {
        return j != 0 ? a.t5(i, j, k) : mix(a.t5(i, g.ny - 2, k), b.t5(i, g.ny - 2, k), sph);
    }
    __device__ __forceinline__ T t4(int i, int j, int k) const {
        return j != g.ny - 2 ? a.t4(i, j, k) : mix(a.t4(i, 1, k), b.t4(i, 1, k), -sph);
    }
    __device__ __forceinline__ T t6(int i, int j, int k) const {
        return j != g.ny - 2 ? a.t6(i, j, k) : mix(a.t6(i, 1, k), b.t6(i, 1, k), -sph);
    }
};

// pa: the real field set (carries the source line), pb: the imaginary one.  threads over (k, i).
template <class A, class M>
__global__ void __launch_bounds__(256) k_pbc_y_bloch(StepArgs<typename A::T> pa, StepArgs<typename A::T> pb, M m,
                                                      typename A::T cph, typename A::T sph) {
    using T = typename A::T;
    const Geo<T> &g = pa.g;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = pa.i_begin + blockIdx.y * blockDim.y + threadIdx.y;
    if (k >= g.nz || i >= pa.i_end) return;
    EvalBloch<A, M> ea(g, pa.cur, pb.cur, m, cph, sph), eb(g, pb.cur, pa.cur, m, cph, -sph);
    // the rows the wrapped stresses reach, for both parts, from the CURRENT fields ...
    naive_cell_ev<A>(pa, ea, i, g.ny - 2, k);
    naive_cell_ev<A>(pa, ea, i, 0, k);
    naive_cell_ev<A>(pb, eb, i, g.ny - 2, k);
    naive_cell_ev<A>(pb, eb, i, 0, k);
    // ... then the displacement copies with the phase (this thread wrote every value it reads here, or the step kernel did)
    const long long r0 = g.idx(i, 0, k), r1 = g.idx(i, 1, k), rm = g.idx(i, g.ny - 2, k), rl = g.idx(i, g.ny - 1, k);
    auto mix = [&](T x, T y, T s) { return A::add(A::mul(cph, x), A::mul(s, y)); };
    if (i <= g.nx - 2) {
        const T xa = pa.nw.ux[rm], xb = pb.nw.ux[rm];
        pa.nw.ux[r0] = mix(xa, xb, sph);
        pb.nw.ux[r0] = mix(xb, xa, -sph);
        pa.nw.ux[rl] = pa.cur.ux[rl];
        pb.nw.ux[rl] = pb.cur.ux[rl];
    }
    if (k <= g.nz - 2) {
        const T xa = pa.nw.uz[rm], xb = pb.nw.uz[rm];
        pa.nw.uz[r0] = mix(xa, xb, sph);
        pb.nw.uz[r0] = mix(xb, xa, -sph);
        pa.nw.uz[rl] = (i == 0 && k == 0 && pa.line_save) ? pa.line_save[g.ny - 1] : pa.cur.uz[rl];
        pb.nw.uz[rl] = (i == 0 && k == 0 && pb.line_save) ? pb.line_save[g.ny - 1] : pb.cur.uz[rl];
    }
    {
        const T ya = pa.nw.uy[r1], yb = pb.nw.uy[r1];
        pa.nw.uy[rm] = mix(ya, yb, -sph);
        pb.nw.uy[rm] = mix(yb, ya, sph);
    }
}

// Stress dump in the reference's array shapes (double), for phb_get_stress.
// nT1 planes etc. are the owned plane counts; out arrays are plane-major like the host arrays.
template <class A, class M>
__global__ void k_stress_dump(Geo<typename A::T> g, Fld<typename A::T> u, M m, int i_begin, int i_end,
                              double *T1, double *T2, double *T3, double *T4, double *T5, double *T6, bool pbc) {
    using T = typename A::T;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = i_begin + blockIdx.z;
    if (k >= g.nz || j >= g.ny || i >= i_end) return;
    Eval<A, M> ev(g, u, m, pbc);
    const long long li = i - i_begin;
    T t1, t2, t3;
    ev.normal(i, j, k, t1, t2, t3);
    const long long n = (li * g.ny + j) * g.nz + k;
    T1[n] = (double)t1; T2[n] = (double)t2; T3[n] = (double)t3;
    if (j < g.ny - 1 && k < g.nz - 1) T4[(li * (g.ny - 1) + j) * (g.nz - 1) + k] = (double)ev.t4(i, j, k);
    if (i < g.nx - 1 && k < g.nz - 1) T5[(li * g.ny + j) * (g.nz - 1) + k] = (double)ev.t5(i, j, k);
    if (i < g.nx - 1 && j < g.ny - 1) T6[(li * (g.ny - 1) + j) * g.nz + k] = (double)ev.t6(i, j, k);
}

}  // namespace phb

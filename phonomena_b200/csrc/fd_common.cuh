// fd_common.cuh -- shared device-side definitions for the sm_100a FDTD kernels.
//
// What is computed is the time step of phonomena/simulation/base_solver.py:245-260:
//   update_T (:323-372), apply_T_tfbc (:402-433), update_u (:435-463), apply_u_tfbc (:488-517)
// fused into one pass: the six stresses are on-chip intermediates, never stored.
// Index-level semantics: SURVEY.md App. A.  Quirks that are replicated on purpose: App. B.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace phb {

// ---------------------------------------------------------------------------------------
// Arithmetic policy.
//  EXACT: every operation is a separately rounded IEEE op in the reference's order
//         (`C*diff/sd` left to right, true division), so fp64 results are bit-identical to
//         NumPy's.  The *_rn intrinsics are never contracted into FMAs by nvcc.
//  FAST : spacing arrays hold reciprocals, `scl` multiplies, the compiler may contract.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double rn_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double rn_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double rn_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double rn_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float rn_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float rn_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float rn_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float rn_div(float a, float b) { return __fdiv_rn(a, b); }

// COMP (FAST arithmetic only): the state is (u, delta = u - u_old) instead of (u, u_old);
//   delta_new = delta + (d2/rho) * acc,  u_new = u + delta_new
// -- algebraically the same step, but the rounding of u no longer feeds back through the two-step
// recurrence, which is what makes plain fp32 drift past 1e-5 after ~10^4 steps.
template <class T_, bool EXACT_, bool COMP_ = false>
struct Ar {
    using T = T_;
    static constexpr bool EXACT = EXACT_;
    static constexpr bool COMP = COMP_;
    static_assert(!(EXACT_ && COMP_), "the compensated state is a FAST-arithmetic mode");
    static __device__ __forceinline__ T add(T a, T b) {
        if constexpr (EXACT) return rn_add(a, b); else return a + b;
    }
    static __device__ __forceinline__ T sub(T a, T b) {
        if constexpr (EXACT) return rn_sub(a, b); else return a - b;
    }
    static __device__ __forceinline__ T mul(T a, T b) {
        if constexpr (EXACT) return rn_mul(a, b); else return a * b;
    }
    // "divide by a spacing": s holds the spacing (EXACT) or its reciprocal (FAST)
    static __device__ __forceinline__ T scl(T a, T s) {
        if constexpr (EXACT) return rn_div(a, s); else return a * s;
    }
};

// ---------------------------------------------------------------------------------------
// Geometry / layout.
// Device layout of every displacement component: [local plane l][j][k] with z fastest,
// z pitch nzp (>= nz, multiple of the vector width), local plane l = i - x0 + 1 so that
// l = 0 and l = nxl + 1 are the ghost planes of an x-slab.  All three components use the
// same (nx, ny, nzp) box; the entries the reference's staggered shapes do not have
// (ux at i = nx-1, uy at j = ny-1, uz at k = nz-1, the z padding) stay zero forever.
// ---------------------------------------------------------------------------------------
template <class T>
struct Geo {
    int nx, ny, nz;       // global grid points
    int x0, nxl;          // owned planes [x0, x0+nxl)
    int nzp;              // z pitch in elements
    long long ps;         // plane stride = ny * nzp
    // per-axis spacing tables (spacing in EXACT mode, reciprocal in FAST mode), each
    // addressable on [-1, n]: fd*[m] = x[m+1]-x[m], sd*[m] = (fd[m+1]+fd[m])/2  (grid.py:118-127)
    const T *fdx, *fdy, *fdz, *sdx, *sdy, *sdz;
    // first elements, used on the k = 0 plane (App. B #1)
    T fdx0, fdy0, fdz0, sdx0, sdy0, sdz0;

    __device__ __forceinline__ long long idx(int i, int j, int k) const {
        return ((long long)(i - x0 + 1) * ny + j) * nzp + k;
    }
};

template <class T>
struct Fld {            // one displacement buffer (three components)
    T *ux, *uy, *uz;
};

// Material: one byte per cell = index of a "stencil class".  A class is the 7-tuple of
// material ids the stencil centred on (i,j,k) reads, already shifted to where each update
// needs them (k = 0 variants included) and already VOIDed where the reference never writes
// the corresponding stress / displacement (App. A.1 "never written => identically 0"), so the
// range masks of the reference's slices are data, not code:
//   node : C rows 0..2 at (i,j,k)            -> T1,T2,T3(i,j,k)   (row 3 zeroed on k = 0: T3 = 0, :418)
//   t4   : C44 at (i, j+1, k+1 | k=0: 0)     -> T4(i,j,k)         (base_solver.py:353,421)
//   t5   : C55 at (i+1, j, k+1 | k=0: 0)     -> T5(i,j,k)         (:360,426)
//   t6   : C66 at (i+1, j+1, k)              -> T6(i,j,k)         (:367,431)
//   rx   : rho at (i+1, j, k)                -> ux_new(i,j,k)     (:443,497)
//   ry   : rho at (i, j+1, k)                -> uy_new(i,j,k)     (:451,505)
//   rz   : rho at (i, j, k+1 | k=0: 0)       -> uz_new(i,j,k)     (:459,513)
// Class-table row (CLS_W values): c11 c12 c13 c21 c22 c23 c31 c32 c33 c44 c55 c66 rx ry rz pad,
// with rX = dt^2/rho; a VOID entry is 0.  Classes are the distinct tuples present in the
// grid, sorted by key (deterministic numbering), at most 256.
enum { CLS_W = 16, CLS_C44 = 9, CLS_C55 = 10, CLS_C66 = 11, CLS_RX = 12, CLS_RY = 13, CLS_RZ = 14, MAX_MAT = 15,
       MAT_VOID = 15, MAX_CLS = 256 };

template <class T>
struct MatCls {
    const uint8_t *code;   // same indexing as the fields
    const T *tab;          // ncls x CLS_W
    int ncls;
};

// 4 bits per id, 7 ids, bit 28 = "k == 0" (T3 forced to zero)
__host__ __device__ inline uint32_t cls_key(const int id[7], bool k0) {
    uint32_t key = k0 ? (1u << 28) : 0u;
    for (int f = 0; f < 7; ++f) key |= (uint32_t)(id[f] & 15) << (4 * f);
    return key;
}

// ---------------------------------------------------------------------------------------
// Formulas on values (shared by every kernel so that all variants round identically).
// ---------------------------------------------------------------------------------------
// T_r = C[r][0]*dxx/sx + C[r][1]*dyy/sy + C[r][2]*dzz/sz      (base_solver.py:329-351,408-416)
template <class A>
__device__ __forceinline__ typename A::T normal_row(const typename A::T *c3, typename A::T dxx, typename A::T dyy,
                                                    typename A::T dzz, typename A::T sx, typename A::T sy,
                                                    typename A::T sz) {
    return A::add(A::add(A::scl(A::mul(c3[0], dxx), sx), A::scl(A::mul(c3[1], dyy), sy)),
                  A::scl(A::mul(c3[2], dzz), sz));
}
// T = C * (a/sa + b/sb)                                        (base_solver.py:353-372,420-433)
template <class A>
__device__ __forceinline__ typename A::T shear(typename A::T c, typename A::T a, typename A::T sa,
                                               typename A::T b, typename A::T sb) {
    return A::mul(c, A::add(A::scl(a, sa), A::scl(b, sb)));
}
// first-order Mur face: q_new[face] = q[inner] + c * (q_new[inner] - q[face])   (base_solver.py:539-554, App. A.5)
template <class A>
__device__ __forceinline__ typename A::T mur(typename A::T q_inner, typename A::T qn_inner, typename A::T q_face,
                                             typename A::T c) {
    return A::add(q_inner, A::mul(c, A::sub(qn_inner, q_face)));
}
// u_new = 2u - u_old + (d2/rho) * acc                          (base_solver.py:441-463,495-517)
template <class A>
__device__ __forceinline__ typename A::T advance(typename A::T u, typename A::T uo, typename A::T rinv,
                                                 typename A::T acc) {
    return A::add(A::sub(A::mul((typename A::T)2, u), uo), A::mul(rinv, acc));
}

// ---------------------------------------------------------------------------------------
// Generic (global-memory) stress evaluation with the reference's write ranges: anything the
// reference never writes is identically zero (App. A.1).  Used by the naive kernel, by the
// stress dump (phb_get_stress) and as the specification the marching kernel is tested against.
// ---------------------------------------------------------------------------------------
template <class A, class M>
struct Eval {
    using T = typename A::T;
    const Geo<T> &g;
    const Fld<T> &u;
    const M &m;
    // Periodic y boundaries, zero Bloch phase -- the reference's archived apply_T_pbc (base_solver.py:383-400), applied
    // after the traction-free stresses: T1, T2, T3, T5 on row 0 are those of row ny-2; T4, T6 on their last row (ny-2)
    // are those of row 1.
    bool pbc;
    __device__ Eval(const Geo<T> &g_, const Fld<T> &u_, const M &m_, bool pbc_ = false) : g(g_), u(u_), m(m_), pbc(pbc_) {}

    // coefficient e of the class of cell (i,j,k)
    __device__ __forceinline__ T coef(int i, int j, int k, int e) const {
        return __ldg(m.tab + (int)__ldg(m.code + g.idx(i, j, k)) * CLS_W + e);
    }

    // T1, T2, T3 at node (i,j,k)
    __device__ __forceinline__ void normal(int i, int j, int k, T &t1, T &t2, T &t3) const {
        t1 = t2 = t3 = (T)0;
        if (pbc && j == 0) j = g.ny - 2;
        if (i < 1 || i > g.nx - 2 || j < 1 || j > g.ny - 2 || k < 0 || k > g.nz - 2) return;
        const bool k0 = (k == 0);
        const T dxx = A::sub(u.ux[g.idx(i, j, k)], u.ux[g.idx(i - 1, j, k)]);
        const T dyy = A::sub(u.uy[g.idx(i, j, k)], u.uy[g.idx(i, j - 1, k)]);
        const T dzz = k0 ? u.uz[g.idx(i, j, 0)] : A::sub(u.uz[g.idx(i, j, k)], u.uz[g.idx(i, j, k - 1)]);
        const T sx = k0 ? g.sdx0 : g.sdx[i - 1];
        const T sy = k0 ? g.sdy0 : g.sdy[j - 1];
        const T sz = k0 ? g.sdz0 : g.sdz[k - 1];
        T c[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) c[e] = coef(i, j, k, e);
        t1 = normal_row<A>(c + 0, dxx, dyy, dzz, sx, sy, sz);
        t2 = normal_row<A>(c + 3, dxx, dyy, dzz, sx, sy, sz);
        t3 = k0 ? (T)0 : normal_row<A>(c + 6, dxx, dyy, dzz, sx, sy, sz);
    }
    __device__ __forceinline__ T t4(int i, int j, int k) const {
        if (pbc && j == g.ny - 2) j = 1;
        if (i < 1 || i > g.nx - 2 || j < 0 || j > g.ny - 2 || k < 0 || k > g.nz - 2) return (T)0;
        const bool k0 = (k == 0);
        const T a = A::sub(u.uy[g.idx(i, j, k + 1)], u.uy[g.idx(i, j, k)]);
        const T b = A::sub(u.uz[g.idx(i, j + 1, k)], u.uz[g.idx(i, j, k)]);
        return shear<A>(coef(i, j, k, CLS_C44), a, k0 ? g.fdy0 : g.fdz[k], b, k0 ? g.fdz0 : g.fdy[j]);
    }
    __device__ __forceinline__ T t5(int i, int j, int k) const {
        if (pbc && j == 0) j = g.ny - 2;
        if (i < 0 || i > g.nx - 2 || j < 1 || j > g.ny - 2 || k < 0 || k > g.nz - 2) return (T)0;
        const bool k0 = (k == 0);
        const T a = A::sub(u.ux[g.idx(i, j, k + 1)], u.ux[g.idx(i, j, k)]);
        const T b = A::sub(u.uz[g.idx(i + 1, j, k)], u.uz[g.idx(i, j, k)]);
        return shear<A>(coef(i, j, k, CLS_C55), a, k0 ? g.fdx0 : g.fdz[k], b, k0 ? g.fdz0 : g.fdx[i]);
    }
    __device__ __forceinline__ T t6(int i, int j, int k) const {
        if (pbc && j == g.ny - 2) j = 1;
        if (i < 0 || i > g.nx - 2 || j < 0 || j > g.ny - 2 || k < 0 || k > g.nz - 2) return (T)0;
        const bool k0 = (k == 0);
        const T a = A::sub(u.ux[g.idx(i, j + 1, k)], u.ux[g.idx(i, j, k)]);
        const T b = A::sub(u.uy[g.idx(i + 1, j, k)], u.uy[g.idx(i, j, k)]);
        return shear<A>(coef(i, j, k, CLS_C66), a, k0 ? g.fdx0 : g.fdy[j], b, k0 ? g.fdz0 : g.fdx[i]);
    }
};

}  // namespace phb

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
// One block; the sample index lives in device memory and is advanced here, after every thread has read it, so
// the launch has no per-step argument and a captured CUDA graph of the step can be replayed (phb200.cu step()).
template <class T>
__global__ void k_source(Geo<T> g, T *uz_cur, T *line_save, const double *w, long long *idx) {
    const T v = (T)w[*idx];
    for (int j = threadIdx.x; j < g.ny; j += blockDim.x) {
        const long long c = g.idx(0, j, 0);
        line_save[j] = uz_cur[c];
        uz_cur[c] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) ++*idx;
}

// ---------------------------------------------------------------------------------------
// First-order Mur faces (base_solver.py:539-554), App. A.5:
//   q_new[face] = q[inner] + c * (q_new[inner] - q[face])
// Order matters on shared edges: x = -1, then y = 0 and y = -1, then z = -1; the host
// launches them in that order on one stream.
// Per component: extent (ex, ey, ez) = the reference array shape; coefficient.
// ---------------------------------------------------------------------------------------

#define PHB_FACE(Q, F, N, C)                                          \
    {                                                                 \
        const T v_ = mur<A>(p.cur.Q[N], p.nw.Q[N], p.cur.Q[F], C);    \
        p.nw.Q[F] = v_;                                               \
        if constexpr (A::COMP) p.dl.Q[F] = v_ - p.cur.Q[F];           \
    }

template <class T>
struct AbcArgs {
    Geo<T> g;
    Fld<T> cur, nw;
    Fld<T> dl;            // COMP state: delta arrays (a face value replaces u_new, so delta = u_new - u is refreshed)
    T clx, ctx, cly0, cty0, cly1, cty1, clz, ctz;
    int i_begin, i_end;   // owned planes the y/z faces cover
    int z_edges_only;     // the marching kernel applied the z face itself (StepArgs::zface): redo only the columns whose
                          // inputs the x / y faces have changed since (rows 0, ny-2, ny-1 and planes nx-2, nx-1)
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
        PHB_FACE(ux, f, n, p.clx)
    }
    const long long f = g.idx(g.nx - 1, j, k), n = g.idx(g.nx - 2, j, k);
    if (j < g.ny - 1) PHB_FACE(uy, f, n, p.ctx)
    if (k < g.nz - 1) PHB_FACE(uz, f, n, p.ctx)
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
        PHB_FACE(ux, f, n, ct)
    }
    {                                         // uy: rows 0|ny-2, inner 1|ny-3
        const int jf = hi ? g.ny - 2 : 0, jn = hi ? g.ny - 3 : 1;
        const long long f = g.idx(i, jf, k), n = g.idx(i, jn, k);
        PHB_FACE(uy, f, n, cl)
    }
    if (k < g.nz - 1) {                       // uz: rows 0|ny-1
        const int jf = hi ? g.ny - 1 : 0, jn = hi ? g.ny - 2 : 1;
        const long long f = g.idx(i, jf, k), n = g.idx(i, jn, k);
        PHB_FACE(uz, f, n, ct)
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
    if (p.z_edges_only && !(j == 0 || j >= g.ny - 2 || i >= g.nx - 2)) return;
    if (i < g.nx - 1) {
        const long long f = g.idx(i, j, g.nz - 1), n = f - 1;
        PHB_FACE(ux, f, n, p.ctz)
    }
    if (j < g.ny - 1) {
        const long long f = g.idx(i, j, g.nz - 1), n = f - 1;
        PHB_FACE(uy, f, n, p.ctz)
    }
    {
        const long long f = g.idx(i, j, g.nz - 2), n = f - 1;
        PHB_FACE(uz, f, n, p.clz)
    }
}

// ---------------------------------------------------------------------------------------
// All faces in ONE launch, for steps whose stencil kernel has applied the z = -1 face itself (StepArgs::zface):
// what is left is the x face, the two y faces and the z face on the columns whose inputs the x / y faces change
// (rows 0, ny-2, ny-1 and planes nx-2, nx-1).  The reference's order x -> y -> z (later faces read what earlier
// ones wrote on shared edges) is kept WITHOUT ordering the threads: every thread evaluates the earlier faces of
// the inner points it needs itself, from values no face writes (`nw` off the faces, `cur`), with the same mur<A>
// -- bit-identical to the three ordered launches -- and every face entry has exactly one writer: x threads leave
// the y-face rows and the z-face points to the later faces, y threads leave the z-face points to the z threads.
// 1-D grid over three segments: [x: ny*nz][y: 2*(ie-ib)*nz][z: 3*(ie-ib) + 2*(ny-3)].
// ---------------------------------------------------------------------------------------
template <class A>
struct FaceEval {
    using T = typename A::T;
    const AbcArgs<T> &p;
    const bool has_x;
    __device__ FaceEval(const AbcArgs<T> &p_, bool hx) : p(p_), has_x(hx) {}
    // comp: 0 ux, 1 uy, 2 uz
    __device__ __forceinline__ const T *cur(int c) const { return c == 0 ? p.cur.ux : c == 1 ? p.cur.uy : p.cur.uz; }
    __device__ __forceinline__ T *nw(int c) const { return c == 0 ? p.nw.ux : c == 1 ? p.nw.uy : p.nw.uz; }
    __device__ __forceinline__ int xplane(int c) const { return c == 0 ? p.g.nx - 2 : p.g.nx - 1; }
    __device__ __forceinline__ T cx(int c) const { return c == 0 ? p.clx : p.ctx; }
    // u_new after the x face
    __device__ __forceinline__ T after_x(int c, int i, int j, int k) const {
        const Geo<T> &g = p.g;
        if (has_x && i == xplane(c))
            return mur<A>(cur(c)[g.idx(i - 1, j, k)], nw(c)[g.idx(i - 1, j, k)], cur(c)[g.idx(i, j, k)], cx(c));
        return nw(c)[g.idx(i, j, k)];
    }
    // y-face rows of component c: side 0 -> (face 0, inner 1), side 1 -> (last row, the one before)
    __device__ __forceinline__ int yface(int c, int side) const { return side == 0 ? 0 : (c == 1 ? p.g.ny - 2 : p.g.ny - 1); }
    __device__ __forceinline__ int yinner(int c, int side) const { return side == 0 ? 1 : (c == 1 ? p.g.ny - 3 : p.g.ny - 2); }
    __device__ __forceinline__ T cy(int c, int side) const { return side == 0 ? (c == 1 ? p.cly0 : p.cty0) : (c == 1 ? p.cly1 : p.cty1); }
    __device__ __forceinline__ T y_value(int c, int side, int i, int k) const {
        const Geo<T> &g = p.g;
        const int jf = yface(c, side), jn = yinner(c, side);
        return mur<A>(cur(c)[g.idx(i, jn, k)], after_x(c, i, jn, k), cur(c)[g.idx(i, jf, k)], cy(c, side));
    }
    // u_new after the x and y faces
    __device__ __forceinline__ T after_xy(int c, int i, int j, int k) const {
        if (j == yface(c, 0)) return y_value(c, 0, i, k);
        if (j == yface(c, 1)) return y_value(c, 1, i, k);
        return after_x(c, i, j, k);
    }
    __device__ __forceinline__ bool exists(int c, int i, int j, int k) const {
        return i < p.g.nx - (c == 0) && j < p.g.ny - (c == 1) && k < p.g.nz - (c == 2);
    }
    __device__ __forceinline__ int zface(int c) const { return c == 2 ? p.g.nz - 2 : p.g.nz - 1; }
};

template <class A>
__global__ void __launch_bounds__(256) k_faces_fused(AbcArgs<typename A::T> p, int has_x) {
    using T = typename A::T;
    const Geo<T> &g = p.g;
    FaceEval<A> f(p, has_x != 0);
    const int np = p.i_end - p.i_begin;
    const long long nX = has_x ? (long long)g.ny * g.nz : 0, nY = 2LL * np * g.nz;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nX) {                                   // ---- x face: thread (j, k), k fastest
        const int k = (int)(t % g.nz), j = (int)(t / g.nz);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int i = f.xplane(c);
            if (!f.exists(c, i, j, k) || k == f.zface(c) || j == f.yface(c, 0) || j == f.yface(c, 1)) continue;
            f.nw(c)[g.idx(i, j, k)] = f.after_x(c, i, j, k);
        }
        return;
    }
    t -= nX;
    if (t < nY) {                                   // ---- y faces: thread (side, i, k), k fastest
        const int k = (int)(t % g.nz);
        const int q = (int)(t / g.nz), i = p.i_begin + q % np, side = q / np;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int jf = f.yface(c, side);
            if (!f.exists(c, i, jf, k) || k == f.zface(c)) continue;
            f.nw(c)[g.idx(i, jf, k)] = f.y_value(c, side, i, k);
        }
        return;
    }
    t -= nY;
    {                                               // ---- z face on the columns the x / y faces touched
        int i, j;
        if (t < 3LL * np) {
            i = p.i_begin + (int)(t % np);
            const int r = (int)(t / np);
            j = r == 0 ? 0 : g.ny - 3 + r;          // rows 0, ny-2, ny-1
        } else {
            t -= 3LL * np;
            if (!has_x || t >= 2LL * (g.ny - 3)) return;
            i = g.nx - 2 + (int)(t / (g.ny - 3));
            j = 1 + (int)(t % (g.ny - 3));          // rows 1 .. ny-3 (the others are in the first set)
            if (i < p.i_begin || i >= p.i_end) return;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int kf = f.zface(c), kn = kf - 1;
            if (!f.exists(c, i, j, kf)) continue;
            const T cz = c == 2 ? p.clz : p.ctz;
            f.nw(c)[g.idx(i, j, kf)] = mur<A>(f.cur(c)[g.idx(i, j, kn)], f.after_xy(c, i, j, kn), f.cur(c)[g.idx(i, j, kf)], cz);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Fused halo push: after the stencil kernel has stored the edge planes into the neighbours' ghost
// planes, tell them which step is complete (flags live in the NEIGHBOUR's memory, CUDA IPC).
// ---------------------------------------------------------------------------------------
// Halo push after the step (default peer-memory mode): copy the slab's first / last owned plane of u_new (all faces
// applied: the values are final) into the matching ghost plane of the left / right neighbour -- 16-byte vectors over
// NVLink through the CUDA-IPC mapping -- then the last block to finish publishes the step number in the neighbours'
// flag words.  src / dst: [edge 0 = low, 1 = high][component]; a null dst row = no neighbour on that side.
// 12 MB per step at 512^2 fp64: ~7 us of NVLink time; storing the same planes from inside the stencil kernel
// (template PUSH of k_step_march, kept as halo mode "fused") costs that memory-bound kernel 38 us through its
// instruction footprint alone (measured with device-local targets).
struct PushArgs {
    const void *src[2][3];
    void *dst[2][3];
    long long vecs;              // 16-byte vectors per plane
    volatile int *flag[2];       // the neighbours' flag words (theirs to read): [0] left neighbour's "from right", [1] right neighbour's "from left"
    int *steps_done;             // device count of completed steps: this kernel publishes steps_done + 1 and advances it
    unsigned *arrive;            // block-arrival counter (zero between launches)
};
__global__ void __launch_bounds__(256) k_push_signal(PushArgs a) {
    const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        if (!a.dst[e][0]) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int4 *s = static_cast<const int4 *>(a.src[e][c]);
            int4 *d = static_cast<int4 *>(a.dst[e][c]);
#pragma unroll 4
            for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < a.vecs; v += stride) d[v] = s[v];
        }
    }
    __threadfence_system();          // this thread's peer stores are visible system-wide
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(a.arrive, 1u);
        if (ticket == gridDim.x - 1) {       // every block's stores are fenced: publish
            *a.arrive = 0;
            const int step = *a.steps_done + 1;
            __threadfence_system();
            if (a.flag[0]) *a.flag[0] = step;
            if (a.flag[1]) *a.flag[1] = step;
            __threadfence_system();
            *a.steps_done = step;
        }
    }
}

// The step number is not a launch argument: both kernels read the count of completed steps from device memory and the
// wait kernel advances it, so the whole slab step can be replayed as a CUDA graph (phb200.cu step()).
__global__ void k_signal(volatile int *left_flag, volatile int *right_flag, const int *steps_done) {
    const int step = *steps_done + 1;
    __threadfence_system();
    if (left_flag) *left_flag = step;
    if (right_flag) *right_flag = step;
    __threadfence_system();
}
// Wait until both neighbours' flags have reached steps_done + plus, then (inc) count the step.  One thread.
//  * push-after-faces mode: plus = 0, inc = 0, at the START of a step -- "the edge planes of every step I have completed
//    have arrived in my ghost planes"; it gates only the launches that read ghost planes (phb200.cu launch_step).
//  * in-kernel push mode: plus = 1, inc = 1, at the end of the step.
// timeout_ns > 0 (phb_sync only): give up after that long -- a neighbour that is gone must not hang the caller.
__global__ void k_wait_flags(const volatile int *from_left, const volatile int *from_right, int *steps_done, int plus, int inc,
                             unsigned long long timeout_ns) {
    const int step = *steps_done + plus;
    unsigned long long t0 = 0;
    if (timeout_ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int side = 0; side < 2; ++side) {
        const volatile int *f = side ? from_right : from_left;
        if (!f) continue;
        while (*f < step) {
            __nanosleep(40);
            if (timeout_ns) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t - t0 > timeout_ns) return;
            }
        }
    }
    __threadfence_system();
    if (inc) *steps_done = step;
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

// decimated gather (PHB_REC_FULL with record_stride): dst (npd, eyd, ezd) <- every (sx, sy, sz)-th entry
template <class T>
__global__ void k_gather_strided(const T *src, double *dst, int npd, int eyd, int ezd, int l0, int ny, int nzp, int sx, int sy, int sz) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int p = blockIdx.z;
    if (k >= ezd || j >= eyd || p >= npd) return;
    dst[((long long)p * eyd + j) * ezd + k] = (double)src[((long long)(l0 + p * sx) * ny + (long long)j * sy) * nzp + (long long)k * sz];
}

// COMP state: the `old` slot holds delta = u - u_old.  Host-side "u_old" <-> device delta.
template <class T>
__global__ void k_scatter_delta(const double *src_old, const T *cur, T *delta, int np, int ey, int ez, int l0, int ny, int nzp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int p = blockIdx.z;
    if (k >= ez || j >= ey || p >= np) return;
    const long long d = ((long long)(l0 + p) * ny + j) * nzp + k;
    delta[d] = (T)((double)cur[d] - src_old[((long long)p * ey + j) * ez + k]);
}
template <class T>
__global__ void k_gather_delta(const T *cur, const T *delta, double *dst, int np, int ey, int ez, int l0, int ny, int nzp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int p = blockIdx.z;
    if (k >= ez || j >= ey || p >= np) return;
    const long long d = ((long long)(l0 + p) * ny + j) * nzp + k;
    dst[((long long)p * ey + j) * ez + k] = (double)cur[d] - (double)delta[d];
}

// ---------------------------------------------------------------------------------------
// Stencil classes from the raw id box (see fd_common.cuh "Material").  ids: planes [ib, ie) of
// the global grid, unpadded (ny, nz).  A field whose stress / displacement the reference never
// writes at this cell gets MAT_VOID (App. A.1 ranges); clamped lookups only ever feed VOIDed
// fields.  Pass 1 inserts every cell's key into a small open-addressing set; the host sorts the
// distinct keys; pass 2 writes each cell's index in the sorted list.
// ---------------------------------------------------------------------------------------
struct ClsGeo {
    const uint8_t *ids;
    int ib, ie;          // id planes available
    int nx, ny, nz, nzp, x0, nxl;
};
__device__ __forceinline__ uint32_t cell_key(const ClsGeo &q, int i, int j, int k) {
    int f[7];
    if (i < 0 || i >= q.nx || k >= q.nz) {
        for (int e = 0; e < 7; ++e) f[e] = MAT_VOID;
        return cls_key(f, false);
    }
    auto id = [&](int ii, int jj, int kk) -> int {
        ii = min(max(ii, q.ib), q.ie - 1);
        jj = min(max(jj, 0), q.ny - 1);
        kk = min(max(kk, 0), q.nz - 1);
        return q.ids[((long long)(ii - q.ib) * q.ny + jj) * q.nz + kk];
    };
    const bool k0 = (k == 0);
    const int kz = k0 ? 0 : k + 1;                     // the k = 0 plane reads its own z level (App. A.2/A.4)
    const bool vk = (k <= q.nz - 2);
    const bool xi = (i >= 1 && i <= q.nx - 2), xs = (i <= q.nx - 2);      // i >= 0 here
    const bool yi = (j >= 1 && j <= q.ny - 2), ys = (j <= q.ny - 2);      // j >= 0 always
    f[0] = (xi && yi && vk) ? id(i, j, k) : MAT_VOID;              // T1..T3
    f[1] = (xi && ys && vk) ? id(i, j + 1, kz) : MAT_VOID;         // T4
    f[2] = (xs && yi && vk) ? id(i + 1, j, kz) : MAT_VOID;         // T5
    f[3] = (xs && ys && vk) ? id(i + 1, j + 1, k) : MAT_VOID;      // T6
    f[4] = (xs && yi && vk) ? id(i + 1, j, k) : MAT_VOID;          // ux_new
    f[5] = (xi && ys && vk) ? id(i, j + 1, k) : MAT_VOID;          // uy_new
    f[6] = (xi && yi && vk) ? id(i, j, kz) : MAT_VOID;             // uz_new
    return cls_key(f, k0 && f[0] != MAT_VOID);
}
enum { CLS_SLOTS = 4096, CLS_EMPTY = 0xFFFFFFFFu };
__global__ void k_cls_collect(ClsGeo q, uint32_t *slots, int *overflow) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int l = blockIdx.z;                    // local plane 0 .. nxl+1
    if (k >= q.nzp || j >= q.ny || l >= q.nxl + 2) return;
    const uint32_t key = cell_key(q, q.x0 - 1 + l, j, k);
    uint32_t h = (key * 2654435761u) >> 20;      // 12 bits
    for (int probe = 0; probe < CLS_SLOTS; ++probe, h = (h + 1) & (CLS_SLOTS - 1)) {
        uint32_t cur = slots[h];
        if (cur == key) return;
        if (cur == CLS_EMPTY) {
            cur = atomicCAS(&slots[h], CLS_EMPTY, key);
            if (cur == CLS_EMPTY || cur == key) return;
        }
    }
    *overflow = 1;
}
__global__ void k_cls_assign(ClsGeo q, const uint32_t *keys, int nkeys, uint8_t *code) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int l = blockIdx.z;
    if (k >= q.nzp || j >= q.ny || l >= q.nxl + 2) return;
    const uint32_t key = cell_key(q, q.x0 - 1 + l, j, k);
    // keys[0] is the all-VOID class; keys[1..] are sorted
    int lo = 1, hi = nkeys - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    if (nkeys < 2 || keys[lo] != key) lo = 0;
    code[((long long)l * q.ny + j) * q.nzp + k] = (uint8_t)lo;
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

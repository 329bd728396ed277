// k_march.cuh -- the production step kernel for sm_100a.
//
// One fused pass per time step: u_new = f(u_cur stencil, u_old, material); the six stresses
// never touch HBM.  Structure (DESIGN.md "K1"):
//
//  * The reference stores z fastest, so lanes run along z (V cells per lane, 16-byte
//    vectors) and the block MARCHES along x, the slowest axis -- the role "z-marching"
//    plays in the usual 2.5-D blocking.  A block owns an (R rows x 32*V cells) tile of the
//    (y, z) plane and a chunk of x-planes; the first and last tile rows are the y-halo rows
//    (stresses only), the TY = R - 2 rows between them produce u_new.
//  * RW rows per warp (template; production: R = 16, RW = 2 -> 8 warps x 255 registers).  A thread
//    owns RW x V cells with independent dependency chains; every phase of the plane loop is one
//    basic block over all rows, so the chains interleave, and the y-exchange, the barrier traffic
//    and the address arithmetic are paid once per RW rows.  RW = 1 is the 16-warp variant.
//  * EVERY input arrives by TMA (cp.async.bulk.tensor.4d / .3d -> UTMALDG): per x-plane one stage of
//    an NST-deep shared ring gets the three u_cur tiles in ONE 4-D transfer (z, y, plane,
//    component; 16-byte halo vector on each side in z, halo row on each side in y) plus the
//    1-byte stencil-class tile; the three u_old tiles (no halo, needed only late in the
//    iteration) use their own ring, again one 4-D transfer.  Out-of-range rows / columns /
//    planes are zero-filled by the TMA unit, so there is no load-side boundary code and no global
//    load instruction in the plane loop.  `full[s]` mbarriers signal arrival; `done[s]` (bar_empty,
//    one arrival per warp) releases a stage.
//  * Ring refill is round robin: in iteration `it` warp it % W reloads the stage plane it - 1
//    occupied, between its publish and its neighbour wait (PHB_REFILL == 1).  No producer warp, no
//    barrier test or counter at the end of a plane, and the extra work never lands on the warp that
//    is already last.  Every issued TMA is waited for by some warp before the block exits.
//  * There is NO block-wide barrier in the plane loop.  Per plane a warp publishes T2 of its first
//    row and T4/T6 of its last row (the stresses its y-neighbour warps need) into a double-buffered
//    exchange tile and arrives on its two neighbours' `pub[w][parity]` mbarriers; it then waits on
//    its own, once.  Rows inside a warp exchange through registers.
//  * Multi-GPU (template PUSH): the first / last owned plane of u_new is also stored into the
//    neighbours' ghost planes through CUDA-IPC peer pointers (NVLink), see phb200.cu step().
//  * Per plane a thread keeps in registers, per cell: u(n) (own position), the normal stresses
//    T1..T3(n) and the shear stresses T5(n-1), T6(n-1) carried from the previous plane.
//    z-neighbours come from warp shuffles; the two edge lanes rebuild the three halo stresses
//    they need from the ring's halo vectors.
//  * The reference's slice ranges ("never written => 0") are not code here: the per-cell
//    stencil class (1 byte, fd_common.cuh) selects a 16-entry coefficient row in shared memory
//    that is already zero wherever a stress or an update does not exist.
//  * u_new leaves as 16-byte vector stores; with template ZF the block that owns k = nz-1 first applies
//    the z = -1 absorbing face to its own results (branch-free selects in the face lane).
//  * Split step (phb200.cu physics()): a step is launched as up to three instantiations side by side, one per class
//    of z-tiles -- the tile with k = 0 (K0 = true, EDGE = 1), the tiles between (K0 = false: no first-element
//    selects), the tile that owns the face (ZF, K0 = false, EDGE = 2) -- so that no block carries code only another
//    tile can need; blockIdx.x is offset by StepArgs::ztile0.
//
// Arithmetic goes through the same formula functions as the naive kernel, so in EXACT mode
// both are bit-identical to the reference.
#pragma once
#include <cuda.h>

#include <cstdlib>

#include <type_traits>

#include "k_naive.cuh"

#ifndef PHB_UNROLL
#define PHB_UNROLL 2
#endif
#ifndef PHB_UNROLL_RW2
#define PHB_UNROLL_RW2 2
#endif
#ifndef PHB_NSO_RW2
#define PHB_NSO_RW2 3
#endif
// Development-only timing variants (never built into libphb200.so; tools/diag_build.sh):
//   PHB_DIAG == 1  "memory only": the TMA ring, the u_old ring and the vector stores run, the arithmetic and the
//                  warp-to-warp exchange do not -> the rate the load/store pipeline allows
//   PHB_DIAG == 2  "compute only": nothing is loaded (the tiles hold whatever shared memory holds), arithmetic,
//                  exchange and stores run -> the rate the per-warp dependency chains allow
#ifndef PHB_DIAG
#define PHB_DIAG 0
#endif
// who refills the TMA rings: 1 = round robin over the warps, mid-iteration (see the plane loop); 0 = whichever warp
// finds a `done` phase complete after its own arrival (barrier test + CAS in every warp at the end of every plane)
#ifndef PHB_REFILL
#define PHB_REFILL 1
#endif

namespace phb {

template <class T> struct VecOf;
template <> struct VecOf<float> { static constexpr int V = 4; };
template <> struct VecOf<double> { static constexpr int V = 2; };

// R = tile rows (TY = R - 2 of them produce output), RW = rows per warp (R / RW warps), NST = depth of the u_cur ring.
template <class T, int R, int NST, int RW = 1>
struct MarchCfg {
    static_assert(R % RW == 0, "rows per warp must divide the tile rows");
    static constexpr int V = VecOf<T>::V;
    static constexpr int SZ = (int)sizeof(T);
    static constexpr int W = R / RW;                  // warps per block
    // depth of the u_old ring: 4 + 2 stages fit 227 KB with one row per warp; two rows per warp halve the
    // exchange tiles, which pays for a third u_old stage
    static constexpr int NSO = (NST >= 4) ? (RW >= 2 ? PHB_NSO_RW2 : 2) : NST;
    static constexpr int TZ = 32 * V;                 // cells per tile row
    static constexpr int TY = R - 2;                  // output rows per tile
    static constexpr int ROWB = 34 * 16;              // u_cur ring row: 34 vectors of 16 B (halo vector each side)
    static constexpr int UCOMP = R * ROWB;            // one u_cur component tile
    static constexpr int OROWB = 32 * 16;             // u_old row (no halo)
    static constexpr int OCOMP = TY * OROWB;
    static constexpr int CB = TZ + 32;                // class row: 16 halo bytes each side
    static constexpr int CTILE = R * CB;
    static constexpr int OFF_C = 3 * UCOMP;           // u_cur stage layout: [U x3][C]
    static constexpr int STAGE = (OFF_C + CTILE + 127) / 128 * 128;
    static constexpr int OSTAGE = 3 * OCOMP;          // u_old stage layout: [O x3]
    static constexpr int OFF_O = NST * STAGE;
    static constexpr int XROWB = 32 * 16;
    static constexpr int XCOMP = W * XROWB;           // one vector per lane and warp
    static constexpr int XBUF = 3 * XCOMP;            // T2 (first row of the warp), T4, T6 (its last row)
    static constexpr int OFF_X = OFF_O + NSO * OSTAGE;
    static constexpr int OFF_BAR = OFF_X + 2 * XBUF;
    static constexpr int NBAR = 2 * NST + NSO + 2 * W;   // full[NST], done[NST], fullO[NSO], pub[W][2]
    static constexpr int OFF_SX = OFF_BAR + (NBAR * 8 + 8 + 127) / 128 * 128;   // + the `issued` counter
    static constexpr uint32_t TX_U = 3u * UCOMP + CTILE;   // bytes per u_cur stage / u_old stage
    static constexpr uint32_t TX_O = 3u * OCOMP;
    // x-spacing table: 2 values per plane for planes [ia-2, ib+1]; z table [2][TZ]; y table [R][2]; the two z
    // spacings of the tile's halo columns; class table
    __host__ __device__ static size_t off_ztab(int chunk) { return (size_t)OFF_SX + ((size_t)2 * (chunk + 4) * SZ + 127) / 128 * 128; }
    __host__ __device__ static size_t off_tab(int chunk) { return off_ztab(chunk) + ((size_t)(2 * TZ + 2 * R + 2) * SZ + 127) / 128 * 128; }
    __host__ __device__ static size_t smem_bytes(int chunk, int ncls) { return off_tab(chunk) + (size_t)ncls * CLS_W * SZ; }
};

// ---- PTX helpers (32-bit shared-window addresses throughout) ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: sleep in HW, do not spin
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <class T, int V> struct alignas(sizeof(T) * V) Pack { T v[V]; };

__device__ __forceinline__ Pack<float, 4> lds_vec(uint32_t a, float) {
    Pack<float, 4> r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "r"(a));
    return r;
}
__device__ __forceinline__ Pack<double, 2> lds_vec(uint32_t a, double) {
    Pack<double, 2> r;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts_vec(uint32_t a, const Pack<float, 4> &x) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x.v[0]), "f"(x.v[1]), "f"(x.v[2]), "f"(x.v[3]) : "memory");
}
__device__ __forceinline__ void sts_vec(uint32_t a, const Pack<double, 2> &x) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(x.v[0]), "d"(x.v[1]) : "memory");
}
__device__ __forceinline__ float lds1(uint32_t a, float) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ double lds1(uint32_t a, double) {
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}
// Class-row coefficient loads (rows are CLS_W * sizeof(T) bytes, 64/128-byte aligned):
// entries 0..8 (normal stresses), entries 9..11 (c44 c55 c66), entries 12..14 (rx ry rz).
__device__ __forceinline__ void lds_c9(uint32_t row, float (&c)[9]) {
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3]) : "r"(row));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(c[4]), "=f"(c[5]), "=f"(c[6]), "=f"(c[7]) : "r"(row));
    asm volatile("ld.shared.f32 %0, [%1+32];" : "=f"(c[8]) : "r"(row));
}
__device__ __forceinline__ void lds_c9(uint32_t row, double (&c)[9]) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(c[0]), "=d"(c[1]) : "r"(row));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+16];" : "=d"(c[2]), "=d"(c[3]) : "r"(row));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+32];" : "=d"(c[4]), "=d"(c[5]) : "r"(row));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+48];" : "=d"(c[6]), "=d"(c[7]) : "r"(row));
    asm volatile("ld.shared.f64 %0, [%1+64];" : "=d"(c[8]) : "r"(row));
}
// four entries starting at a multiple of 4 (9..11 live in entries 8..11, 12..14 in 12..15)
__device__ __forceinline__ void lds_c4(uint32_t row, int e0, float (&c)[4]) {
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3]) : "r"(row + e0 * 4));
}
__device__ __forceinline__ void lds_c4(uint32_t row, int e0, double (&c)[4]) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(c[0]), "=d"(c[1]) : "r"(row + e0 * 8));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+16];" : "=d"(c[2]), "=d"(c[3]) : "r"(row + e0 * 8));
}

template <class T> __device__ __forceinline__ T shfl_up1(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <class T> __device__ __forceinline__ T shfl_dn1(T v) { return __shfl_down_sync(0xffffffffu, v, 1); }

struct MarchMaps {
    CUtensorMap u;      // u_cur, all three components (4-D: z, y, plane, component), box (34 V, R, 1, 3)
    CUtensorMap o;      // u_old, all three components,                               box (32 V, TY, 1, 3)
    CUtensorMap c;      // class bytes (3-D),                                         box (32 V + 32, R, 1)
};

// ---- the kernel ---------------------------------------------------------------------------------
// K0 = false: the launch covers no block with k = 0 (StepArgs::ztile0 >= 1), so the first-element spacing selects of the
// free-surface plane (App. B #1-#3) -- 9 % of the fp64 instruction mix, needed by one thread of one z-tile -- fold away.
// EDGE: the launch covers only the first (1) or only the last (2) z-tile, which has no left / right halo column: the
// single-lane sections that rebuild the halo stresses there (5-6 % of a plane's issue slots) fold away; 0 = any tiles.
template <class A, int R, int NST, bool PUSH, int RW, bool ZF = false, bool K0 = true, int EDGE = 0>
__global__ void __launch_bounds__(R / RW * 32, ((R <= 8 && RW == 1) ? 2 : 1))
k_step_march(const __grid_constant__ MarchMaps tm, StepArgs<typename A::T> p, MatCls<typename A::T> m, int chunk) {
    using T = typename A::T;
    using C_ = MarchCfg<T, R, NST, RW>;
    constexpr int V = C_::V, SZ = C_::SZ, TZ = C_::TZ, W = C_::W, NT = W * 32;
    constexpr int UCE = C_::UCOMP / SZ, OCE = C_::OCOMP / SZ, XCE = C_::XCOMP / SZ;   // component strides in elements
    static_assert(NST >= C_::NSO, "u_cur ring must be at least as deep as the u_old ring");
    static_assert((W & (W - 1)) == 0, "the round-robin ring refill takes the warp index modulo a power of two");
    constexpr int ROWE = C_::ROWB / SZ;
    constexpr uint32_t kClsMask = (PHB_DIAG == 2) ? 3u : 255u;   // compute-only timing variant: tiles are never loaded
    using PV = Pack<T, V>;
    using CW = typename std::conditional<V == 4, uint32_t, uint16_t>::type;            // V class bytes
    const Geo<T> &g = p.g;

    extern __shared__ __align__(1024) unsigned char sm[];
    const uint32_t sb = smem_u32(sm);

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r0 = w * RW;                           // first tile row of this warp (rows r0 .. r0 + RW - 1)
    const int k0t = (blockIdx.x + p.ztile0) * TZ;    // first cell of the tile row
    const int j0 = blockIdx.y * C_::TY;              // first output row
    // planes [ia, ib): x-chunk blockIdx.z of [i_begin, i_end), or (edge launch) the two chunks that start at i_begin / edge_b
    const int ia = (p.edge_b >= 0) ? (blockIdx.z == 0 ? p.i_begin : p.edge_b) : p.i_begin + blockIdx.z * chunk;
    const int ib = (p.edge_b >= 0) ? ia + chunk : min(ia + chunk, p.i_end);
    if (ia >= ib) return;
    const int j = j0 - 1 + r0;                       // y index of the warp's first row
    const int kb = k0t + lane * V;                   // first cell of this lane
    const bool k0c = K0 && (kb == 0);                // element 0 of this thread is the k = 0 plane
    const int nplanes = (ib - ia) + 2;               // planes ia-1 .. ib, consumed in order q = 0, 1, ...
    const int lbase = (ia - 1) - g.x0 + 1;           // local plane index of q = 0

    constexpr int NSO = C_::NSO;
    const uint32_t bar_full = sb + C_::OFF_BAR, bar_empty = bar_full + NST * 8, bar_fullO = bar_empty + NST * 8,
                   bar_pub = bar_fullO + NSO * 8;
    int *const issued = reinterpret_cast<int *>(sm + C_::OFF_BAR + C_::NBAR * 8);
    // small tables: x spacings per plane, z spacings per cell of the tile row, y spacings per row
    T *const sxp = reinterpret_cast<T *>(sm + C_::OFF_SX);                        // [plane-(ia-2)][2] = fdx, sdx
    T *const ztab = reinterpret_cast<T *>(sm + C_::off_ztab(chunk));              // [2][TZ] = fdz[k], sdz[k-1]
    T *const ytab = ztab + 2 * TZ;                                                // [R][2]  = fdy[j], sdy[j-1]; then fdz[kL], sdz[kR-1]
    const T *const stab = reinterpret_cast<const T *>(__builtin_assume_aligned(sm + C_::off_tab(chunk), 128));

    // ---- one-time setup ----
    {
        T *tabw = reinterpret_cast<T *>(sm + C_::off_tab(chunk));
        for (int q = threadIdx.x; q < m.ncls * CLS_W; q += NT) tabw[q] = m.tab[q];
        for (int q = threadIdx.x; q < nplanes + 2; q += NT) {
            const int n = min(max(ia - 2 + q, -1), g.nx);          // tables are addressable on [-1, n]
            sxp[2 * q + 0] = g.fdx[n];
            sxp[2 * q + 1] = g.sdx[n];
        }
        for (int q = threadIdx.x; q < TZ; q += NT) {
            const int k = min(k0t + q, g.nz);
            ztab[q] = g.fdz[k];
            ztab[TZ + q] = (k == 0) ? g.sdz0 : g.sdz[k - 1];       // k = 0 uses sdz[0] (App. B #1)
        }
        if (threadIdx.x < R) {
            const int jj = min(max(j0 - 1 + (int)threadIdx.x, 0), g.ny);
            ytab[2 * threadIdx.x + 0] = g.fdy[jj];
            ytab[2 * threadIdx.x + 1] = g.sdy[jj - 1];
        }
        if (threadIdx.x == R) {   // halo columns kL = k0t - 1 (shear) and kR = k0t + TZ (its T3 needs sdz[kR - 1])
            ytab[2 * R + 0] = g.fdz[k0t - 1];
            ytab[2 * R + 1] = g.sdz[min(k0t + TZ - 1, g.nz)];
        }
        if (threadIdx.x == 0) {
            for (int s = 0; s < NST; ++s) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, W); }
            for (int s = 0; s < NSO; ++s) mbar_init(bar_fullO + s * 8, 1);
            for (int q = 0; q < 2 * W; ++q) mbar_init(bar_pub + q * 8, (q / 2 == 0 || q / 2 == W - 1) ? 1 : 2);   // arrivals = neighbours
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();

    // producer: u_cur/class plane q -> stage q % NST; u_old plane q -> stage q % NSO, issued NST-NSO planes
    // later (it is only needed in the second half of iteration q and must not hold a deep ring).
    auto issue_u = [&](int q) {
        const int s = q % NST;
        const uint32_t dst = sb + s * C_::STAGE, bar = bar_full + s * 8;
        mbar_expect_tx(bar, C_::TX_U);
        tma_load_4d(dst, &tm.u, bar, k0t - V, j0 - 1, lbase + q, 0);
        tma_load_3d(dst + C_::OFF_C, &tm.c, bar, k0t - 16, j0 - 1, lbase + q);
    };
    auto issue_o = [&](int q) {
        const int s = q % NSO;
        const uint32_t dst = sb + C_::OFF_O + s * C_::OSTAGE, bar = bar_fullO + s * 8;
        mbar_expect_tx(bar, C_::TX_O);
        tma_load_4d(dst, &tm.o, bar, k0t, j0, lbase + q, 0);
    };
    // claim number c: u_cur plane c (if c < nplanes) and u_old plane c - (NST - NSO) if that plane is one
    // whose u_new is produced (1 .. nplanes-2).  Every issued load is waited for by some warp before the
    // block exits -- a TMA still in flight at exit would land in the next block's shared memory.
    auto issue = [&](int c) {
        if (c < nplanes) issue_u(c);
        const int qo = c - (NST - NSO);
        if (qo >= 1 && qo <= nplanes - 2) issue_o(qo);
    };
    const int nclaims = max(nplanes, nplanes - 1 + (NST - NSO));   // last u_cur plane / last u_old plane
    // `issued` = next claim.  Nobody blocks to produce: the lane that sees a `done[s]` phase complete
    // after its own arrival claims with a CAS and issues.
    if (threadIdx.x == 0) {
        if (PHB_DIAG != 2) for (int c = 0; c < NST; ++c) issue(c);
        *issued = (PHB_DIAG == 2) ? nclaims : NST;
    }
    __syncthreads();

    // ---- loop-invariant per-thread quantities (kept few: everything else is re-read from smem) ----
    // Row q of the warp is tile row r0 + q, y index j + q.  The first and last TILE rows are the y-halo rows:
    // they only produce what their one neighbour reads (bottom row: T4 and T6, top row: T2); everything else
    // they would compute is dead (warp-uniform branches).
    bool row_out[RW], in_box[RW], need_shear[RW], need_normal[RW];
    bool any_out = false;
#pragma unroll
    for (int q = 0; q < RW; ++q) {
        const int rr = r0 + q;
        row_out[q] = (rr >= 1 && rr <= R - 2) && (j + q < g.ny);           // this row produces u_new (warp-uniform)
        in_box[q] = (j + q >= 0 && j + q < g.ny && kb < g.nzp);            // global vector stores allowed
        need_shear[q] = (rr != R - 1);
        need_normal[q] = (rr != 0);
        any_out = any_out || row_out[q];
    }
    const bool haveL = (EDGE != 1) && (lane == 0) && (k0t >= 1);                  // halo column kL = k0t-1 exists
    const bool haveR = (EDGE != 2) && (lane == 31) && (k0t + TZ <= g.nz - 1);     // halo column kR = k0t+TZ exists
    const int eo = r0 * ROWE + (lane + 1) * V;         // own vector (row 0 of the warp) inside a u_cur component tile (elements)
    const int eS = (r0 >= 1) ? -ROWE : 0;              // row below the warp's first row (clamped)
    const int eN = (r0 + RW - 1 <= R - 2) ? ROWE : 0;  // row above the warp's last row (clamped), relative to that row
    const int xo = w * TZ + lane * V;                  // own vector inside an exchange component tile
    const int xS = (w >= 1) ? -TZ : 0, xN = (w <= W - 2) ? TZ : 0;
    const uint32_t pubO = bar_pub + (uint32_t)w * 16;
    // lane that holds k = nz-2, nz-1 if this block applies the z face itself, else -1
    const int zlane = (ZF && p.zface == 1 && k0t <= g.nz - 1 && g.nz - 1 < k0t + TZ) ? (g.nz - 1 - k0t) / V : -1;   // zface == 2: timing aid (face code resident, never run)
    const T *const zt = ztab + lane * V;               // this lane's z spacings
    const T *const yt = ytab + 2 * r0;

    // ---- registers carried across planes ----
    T t1c[RW][V], t2c[RW][V], t3c[RW][V];    // T1..T3(n)
    T t5m[RW][V], t6m[RW][V];                // T5(n-1), T6(n-1)
    T t3R[RW];                               // T3(n, j, kR) for lane 31
    const T *rowc[RW][V];                    // class rows (coefficients) of this thread's cells at plane n
#pragma unroll
    for (int q = 0; q < RW; ++q) {
        t3R[q] = (T)0;
#pragma unroll
        for (int e = 0; e < V; ++e) { t1c[q][e] = t2c[q][e] = t3c[q][e] = t5m[q][e] = t6m[q][e] = (T)0; rowc[q][e] = stab; }
    }

    // element offset of (plane n, j, kb) in a displacement array (the host guarantees it fits 31 bits)
    int off = (int)((long long)lbase * g.ps + (long long)j * g.nzp + kb);
    const int dps = (int)g.ps;

    if (PHB_DIAG != 2) mbar_wait(bar_full, 0);        // plane q = 0 (n = ia - 1)
    int sCi = 0, sNi = 1 % NST;                       // stage indices of plane n and n + 1
    uint32_t phN = 0;                                 // phase parity of full[sNi] for plane n + 1

    bool pushed = false;                              // this thread stored into a neighbour's ghost plane (PUSH)
    constexpr int kUnroll = (RW == 1) ? PHB_UNROLL : PHB_UNROLL_RW2;
#pragma unroll kUnroll
    for (int it = 0; it + 1 < nplanes; ++it) {
        const int n = ia - 1 + it;                    // plane being completed; n + 1 is the newest
        const bool emit = (it >= 1);                  // it = 0 only primes the carried stresses
        const T *const uC = reinterpret_cast<const T *>(sm + sCi * C_::STAGE);     // plane n tiles
        const T *const uN = reinterpret_cast<const T *>(sm + sNi * C_::STAGE);     // plane n+1 tiles
        auto vec = [](const T *q) -> PV { return *reinterpret_cast<const PV *>(q); };

        if (PHB_DIAG != 2) mbar_wait(bar_full + sNi * 8, phN);

        // per-plane x spacings (uniform); element 0 of the k0c thread uses the first elements (App. B #1, #2)
        const T sfx_n = sxp[2 * (it + 1) + 0];        // fdx[n]
        const T ssx_n = sxp[2 * (it + 1) + 1];        // sdx[n]   = sdx[(n+1)-1]
        const T ssx_m = sxp[2 * it + 1];              // sdx[n-1]
        const PV zf = vec(zt), zs = vec(zt + TZ);     // fdz[k], sdz[k-1] (k = 0: sdz[0])

        // With several rows per warp every phase below is ONE basic block over all rows (no per-row branches), so
        // the rows' independent dependency chains interleave; halo rows then compute values nobody reads and only
        // the stores are predicated.  With one row per warp the halo-row warps skip the dead phases instead.
        // (1) shear stresses at plane n
        T t4[RW][V], t5[RW][V], t6[RW][V];
        T t4L[RW], t5L[RW];
        PV uxc[RW], uyc[RW], uzc[RW];
#pragma unroll
        for (int q = 0; q < RW; ++q) {
            uxc[q] = vec(uC + eo + q * ROWE); uyc[q] = vec(uC + UCE + eo + q * ROWE); uzc[q] = vec(uC + 2 * UCE + eo + q * ROWE);
            t4L[q] = t5L[q] = (T)0;
#pragma unroll
            for (int e = 0; e < V; ++e) t4[q][e] = t5[q][e] = t6[q][e] = (T)0;
        }
        if (it == 0 || (RW == 1 && !need_normal[0])) {   // (with one row per warp the bottom halo row never runs the normal-stress phase that carries the rows)
#pragma unroll
            for (int q = 0; q < RW; ++q) {
                const CW cwc = *reinterpret_cast<const CW *>(sm + sCi * C_::STAGE + C_::OFF_C + (r0 + q) * C_::CB + 16 + lane * V);
#pragma unroll
                for (int e = 0; e < V; ++e) rowc[q][e] = stab + (int)((cwc >> (8 * e)) & kClsMask) * CLS_W;
            }
        }
        if ((RW > 1 || need_shear[0]) && PHB_DIAG != 1) {
            const T shb0 = k0c ? g.fdz0 : sfx_n;       // T5, T6 second term on k = 0: fdz[0]
            PV uyn[RW], uzn[RW], uzN[RW], uxN[RW];
            T uyE[RW], uxE[RW];
#pragma unroll
            for (int q = 0; q < RW; ++q) {
                uyn[q] = vec(uN + UCE + eo + q * ROWE); uzn[q] = vec(uN + 2 * UCE + eo + q * ROWE);
                // uz(n, j+1), ux(n, j+1): the warp's own next row, or the ring row above its last row
                uzN[q] = (q < RW - 1) ? uzc[q < RW - 1 ? q + 1 : q] : vec(uC + 2 * UCE + eo + q * ROWE + eN);
                uxN[q] = (q < RW - 1) ? uxc[q < RW - 1 ? q + 1 : q] : vec(uC + eo + q * ROWE + eN);
                const T hy = uC[UCE + (r0 + q) * ROWE + 33 * V], hx = uC[(r0 + q) * ROWE + 33 * V];   // ring halo column kR (broadcast)
                const T sy_ = shfl_dn1(uyc[q].v[0]), sx_ = shfl_dn1(uxc[q].v[0]);                        // uy(n, k+1), ux(n, k+1)
                uyE[q] = (lane == 31) ? hy : sy_;
                uxE[q] = (lane == 31) ? hx : sx_;
            }
#pragma unroll
            for (int q = 0; q < RW; ++q) {
                const T sfy = yt[2 * q];
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const T uye = (e == V - 1) ? uyE[q] : uyc[q].v[e < V - 1 ? e + 1 : e];
                    const T uxe = (e == V - 1) ? uxE[q] : uxc[q].v[e < V - 1 ? e + 1 : e];
                    const T *c = rowc[q][e];
                    const bool k0 = (e == 0) && k0c;
                    const T shb = (e == 0) ? shb0 : sfx_n;
                    t4[q][e] = shear<A>(c[CLS_C44], A::sub(uye, uyc[q].v[e]), k0 ? g.fdy0 : zf.v[e],
                                        A::sub(uzN[q].v[e], uzc[q].v[e]), k0 ? g.fdz0 : sfy);
                    t5[q][e] = shear<A>(c[CLS_C55], A::sub(uxe, uxc[q].v[e]), k0 ? g.fdx0 : zf.v[e],
                                        A::sub(uzn[q].v[e], uzc[q].v[e]), shb);
                    t6[q][e] = shear<A>(c[CLS_C66], A::sub(uxN[q].v[e], uxc[q].v[e]), k0 ? g.fdx0 : sfy,
                                        A::sub(uyn[q].v[e], uyc[q].v[e]), shb);
                }
            }
            if (haveL && (RW > 1 || row_out[0])) {   // T4, T5 at (n, j, kL) for the first element's uy / ux update (never k = 0)
                const T sfzL = ytab[2 * R];
#pragma unroll
                for (int q = 0; q < RW; ++q) {
                    const T *c = stab + (int)(sm[sCi * C_::STAGE + C_::OFF_C + (r0 + q) * C_::CB + 15] & kClsMask) * CLS_W;
                    const T *h = uC + (r0 + q) * ROWE + V - 1;
                    const int hN = (q < RW - 1) ? ROWE : eN;
                    const T uxL = h[0], uyL = h[UCE], uzL = h[2 * UCE], uzLN = h[2 * UCE + hN];
                    const T uzLn = uN[2 * UCE + (r0 + q) * ROWE + V - 1];
                    t4L[q] = shear<A>(c[CLS_C44], A::sub(uyc[q].v[0], uyL), sfzL, A::sub(uzLN, uzL), yt[2 * q]);
                    t5L[q] = shear<A>(c[CLS_C55], A::sub(uxc[q].v[0], uxL), sfzL, A::sub(uzLn, uzL), sfx_n);
                }
            }
        }

        // (2) publish T2(n) of the warp's first row and T4(n), T6(n) of its last row for the y-neighbours and signal them
        T *const xb = reinterpret_cast<T *>(sm + C_::OFF_X + (it & 1) * C_::XBUF) + xo;
        if (PHB_DIAG != 1) {
            PV a, b, c;
#pragma unroll
            for (int e = 0; e < V; ++e) { a.v[e] = t2c[0][e]; b.v[e] = t4[RW - 1][e]; c.v[e] = t6[RW - 1][e]; }
            *reinterpret_cast<PV *>(xb) = a;
            *reinterpret_cast<PV *>(xb + XCE) = b;
            *reinterpret_cast<PV *>(xb + 2 * XCE) = c;
        }
        __syncwarp();
        // tell both neighbours (arrive on THEIR barrier): each warp then waits on one barrier only
        if (lane == 0 && w >= 1 && PHB_DIAG != 1) mbar_arrive(pubO - 16 + (it & 1) * 8);
        if (lane == 1 && w <= W - 2 && PHB_DIAG != 1) mbar_arrive(pubO + 16 + (it & 1) * 8);
#if PHB_REFILL == 1
        // Ring refill, round robin: in iteration `it` warp it % W reloads the stage that plane it - 1 occupied (claim
        // it - 1 + NST).  It does so here, between its publish and its neighbour wait, where a warp has slack, and it
        // waits for a `done` phase that normally completed long ago -- so no warp tests a barrier or touches a counter
        // at the end of its iteration, and the extra work never lands on the warp that is already last.
        if (PHB_DIAG != 2 && it >= 1 && w == (it & (W - 1)) && it - 1 + NST < nclaims) {
            if (lane == 0) {
                mbar_wait(bar_empty + ((it - 1) % NST) * 8, (uint32_t)(((it - 1) / NST) & 1));
                issue(it - 1 + NST);
            }
            __syncwarp();
        }
#endif

        // (3) normal stresses at plane n + 1 (overlaps the neighbours' publishing)
        T t1n[RW][V], t2n[RW][V], t3n[RW][V];
        T t3Rn[RW];
        const T *rown[RW][V];
#pragma unroll
        for (int q = 0; q < RW; ++q) {
            t3Rn[q] = (T)0;
#pragma unroll
            for (int e = 0; e < V; ++e) { t1n[q][e] = t2n[q][e] = t3n[q][e] = (T)0; rown[q][e] = stab; }
        }
        if ((RW > 1 || need_normal[0]) && PHB_DIAG != 1) {
            PV uxn[RW], uyn[RW], uzn[RW], uyS[RW];
            T uzW0[RW];
            CW cwn[RW];
#pragma unroll
            for (int q = 0; q < RW; ++q) {
                uxn[q] = vec(uN + eo + q * ROWE); uyn[q] = vec(uN + UCE + eo + q * ROWE); uzn[q] = vec(uN + 2 * UCE + eo + q * ROWE);
                cwn[q] = *reinterpret_cast<const CW *>(sm + sNi * C_::STAGE + C_::OFF_C + (r0 + q) * C_::CB + 16 + lane * V);
            }
#pragma unroll
            for (int q = 0; q < RW; ++q) {
                uyS[q] = (q >= 1) ? uyn[q >= 1 ? q - 1 : 0] : vec(uN + UCE + eo + eS);      // uy(n+1, j-1)
                const T hz = uN[2 * UCE + (r0 + q) * ROWE + V - 1];                  // ring halo column kL (0 below k = 0)
                const T sz_ = shfl_up1(uzn[q].v[V - 1]);                             // uz(n+1, k-1) for element 0
                uzW0[q] = (lane == 0) ? hz : sz_;
            }
#pragma unroll
            for (int q = 0; q < RW; ++q) {
                const T ssy = yt[2 * q + 1];
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const bool k0 = (e == 0) && k0c;
                    const T dxx = A::sub(uxn[q].v[e], uxc[q].v[e]);
                    const T dyy = A::sub(uyn[q].v[e], uyS[q].v[e]);
                    const T dzz = A::sub(uzn[q].v[e], (e == 0) ? uzW0[q] : uzn[q].v[e > 0 ? e - 1 : 0]);
                    const T sx = k0 ? g.sdx0 : ssx_n, sy = k0 ? g.sdy0 : ssy;
                    const T *c = rown[q][e] = stab + (int)((cwn[q] >> (8 * e)) & kClsMask) * CLS_W;
                    if constexpr (A::EXACT) {
                        t1n[q][e] = normal_row<A>(c + 0, dxx, dyy, dzz, sx, sy, zs.v[e]);
                        t2n[q][e] = normal_row<A>(c + 3, dxx, dyy, dzz, sx, sy, zs.v[e]);
                        t3n[q][e] = normal_row<A>(c + 6, dxx, dyy, dzz, sx, sy, zs.v[e]);
                    } else {   // FAST: the three rows share the scaled strains (spacing tables hold reciprocals)
                        const T ex = dxx * sx, ey = dyy * sy, ez = dzz * zs.v[e];
                        t1n[q][e] = c[0] * ex + c[1] * ey + c[2] * ez;
                        t2n[q][e] = c[3] * ex + c[4] * ey + c[5] * ez;
                        t3n[q][e] = c[6] * ex + c[7] * ey + c[8] * ez;
                    }
                }
            }
            if (haveR && (RW > 1 || row_out[0])) {   // T3(n+1, j, kR) for the next plane's uz update of the last element (never k = 0)
                const T szR = ytab[2 * R + 1];
#pragma unroll
                for (int q = 0; q < RW; ++q) {
                    const T *c = stab + (int)(sm[sNi * C_::STAGE + C_::OFF_C + (r0 + q) * C_::CB + 16 + TZ] & kClsMask) * CLS_W;
                    const T *hC = uC + (r0 + q) * ROWE + 33 * V, *hN = uN + (r0 + q) * ROWE + 33 * V;
                    const T dxx = A::sub(hN[0], hC[0]);
                    const T dyy = A::sub(hN[UCE], hN[UCE + (q >= 1 ? -ROWE : eS)]);
                    const T dzz = A::sub(hN[2 * UCE], uzn[q].v[V - 1]);
                    t3Rn[q] = normal_row<A>(c + 6, dxx, dyy, dzz, ssx_n, yt[2 * q + 1], szR);
                }
            }
        }

        // wait for both neighbours' plane-n stresses (also bounds the drift between warps)
        if (PHB_DIAG != 1) mbar_wait(pubO + (it & 1) * 8, (it >> 1) & 1);

        if (emit && any_out) {
            if (PHB_DIAG != 2) mbar_wait(bar_fullO + (it % NSO) * 8, (uint32_t)(((it - 1) / NSO) & 1));   // planes 1, 2, ... use the ring
            PV ox[RW], oy[RW], oz[RW];
            PV dx[RW], dy[RW], dz[RW];                // COMP: new increments (stored in place of u_old)
            // (4) z-neighbour stresses
            T t3U[RW], t4W[RW], t5W[RW];
#pragma unroll
            for (int q = 0; q < RW; ++q) {
                const T a = shfl_dn1(t3c[q][0]);          // T3(n, k+1) for element V-1
                const T b = shfl_up1(t4[q][V - 1]);       // T4(n, k-1) for element 0
                const T c = shfl_up1(t5[q][V - 1]);       // T5(n, k-1) for element 0
                t3U[q] = (lane == 31) ? t3R[q] : a;
                t4W[q] = (lane == 0) ? t4L[q] : b;        // zero below the k = 0 plane
                t5W[q] = (lane == 0) ? t5L[q] : c;
            }
#pragma unroll
            for (int q = 0; q < RW; ++q) {
                const T sfy = yt[2 * q], ssy = yt[2 * q + 1];                         // fdy[j], sdy[j-1]
                const T *const oC = reinterpret_cast<const T *>(sm + C_::OFF_O + (it % NSO) * C_::OSTAGE) + (r0 + q - 1) * TZ + lane * V;   // u_old(n)
                {   // (5) ux
                    PV t6S;                               // T6(n, j-1): the warp's own previous row, or the neighbour warp's last row
                    if (q >= 1) {
#pragma unroll
                        for (int e = 0; e < V; ++e) t6S.v[e] = t6[q >= 1 ? q - 1 : 0][e];
                    } else t6S = vec(xb + 2 * XCE + xS);
                    const PV uo = vec(oC);
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const bool k0 = (e == 0) && k0c;
                        const T *c = rowc[q][e];
                        const T t5w = (e == 0) ? t5W[q] : t5[q][e > 0 ? e - 1 : 0];
                        const T acc = A::add(A::add(A::scl(A::sub(t1n[q][e], t1c[q][e]), k0 ? g.fdx0 : sfx_n),
                                                    A::scl(A::sub(t6[q][e], t6S.v[e]), k0 ? g.sdy0 : ssy)),
                                             A::scl(A::sub(t5[q][e], t5w), zs.v[e]));
                        if constexpr (A::COMP) { dx[q].v[e] = uo.v[e] + c[CLS_RX] * acc; ox[q].v[e] = uxc[q].v[e] + dx[q].v[e]; }
                        else ox[q].v[e] = advance<A>(uxc[q].v[e], uo.v[e], c[CLS_RX], acc);
                    }
                }
                {   // (6) uy
                    PV t2N;                               // T2(n, j+1): the warp's own next row, or the neighbour warp's first row
                    if (q < RW - 1) {
#pragma unroll
                        for (int e = 0; e < V; ++e) t2N.v[e] = t2c[q < RW - 1 ? q + 1 : q][e];
                    } else t2N = vec(xb + xN);
                    const PV uo = vec(oC + OCE);
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const bool k0 = (e == 0) && k0c;
                        const T *c = rowc[q][e];
                        const T t4w = (e == 0) ? t4W[q] : t4[q][e > 0 ? e - 1 : 0];
                        const T acc = A::add(A::add(A::scl(A::sub(t6[q][e], t6m[q][e]), k0 ? g.sdx0 : ssx_m),
                                                    A::scl(A::sub(t2N.v[e], t2c[q][e]), k0 ? g.fdy0 : sfy)),
                                             A::scl(A::sub(t4[q][e], t4w), zs.v[e]));
                        if constexpr (A::COMP) { dy[q].v[e] = uo.v[e] + c[CLS_RY] * acc; oy[q].v[e] = uyc[q].v[e] + dy[q].v[e]; }
                        else oy[q].v[e] = advance<A>(uyc[q].v[e], uo.v[e], c[CLS_RY], acc);
                    }
                }
                {   // (7) uz   (k = 0: "+ T3[..,1] - T3[..,0]/fdz[0]", T3[..,1] NOT divided, App. B #3)
                    PV t4S;
                    if (q >= 1) {
#pragma unroll
                        for (int e = 0; e < V; ++e) t4S.v[e] = t4[q >= 1 ? q - 1 : 0][e];
                    } else t4S = vec(xb + XCE + xS);
                    const PV uo = vec(oC + 2 * OCE);
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const bool k0 = (e == 0) && k0c;
                        const T *c = rowc[q][e];
                        const T t3u = (e == V - 1) ? t3U[q] : t3c[q][e < V - 1 ? e + 1 : e];
                        const T acc = A::add(A::add(A::scl(A::sub(t5[q][e], t5m[q][e]), k0 ? g.sdx0 : ssx_m),
                                                    A::scl(A::sub(t4[q][e], t4S.v[e]), k0 ? g.sdy0 : ssy)),
                                             A::scl(A::sub(t3u, t3c[q][e]), k0 ? (T)1 : zf.v[e]));
                        if constexpr (A::COMP) { dz[q].v[e] = uo.v[e] + c[CLS_RZ] * acc; oz[q].v[e] = uzc[q].v[e] + dz[q].v[e]; }
                        else oz[q].v[e] = advance<A>(uzc[q].v[e], uo.v[e], c[CLS_RZ], acc);
                    }
                }
            }
            // i = 0: uy, uz keep u_new == u (App. B #9); uz(0, j, 0) is the pre-source value
            if (n == 0) {
#pragma unroll
                for (int q = 0; q < RW; ++q) {
                    oy[q] = uyc[q];
                    oz[q] = uzc[q];
                    if (k0c && p.line_save && row_out[q]) oz[q].v[0] = p.line_save[j + q];
                    if constexpr (A::COMP) {      // keep delta = u_new - u here too (only matters for reading u_old back)
#pragma unroll
                        for (int e = 0; e < V; ++e) { dy[q].v[e] = (T)0; dz[q].v[e] = oz[q].v[e] - uzc[q].v[e]; }
                    }
                }
            }
            // z = -1 absorbing face (base_solver.py:551-554) on the block's own results: ux, uy at k = nz-1 from
            // k = nz-2, uz at k = nz-2 from k = nz-3.  nz is a multiple of V here (host check), so k = nz-2 and nz-1
            // are the last two elements of lane zlane; for V = 2 the uz inner point is the previous lane's last element.
            // Branch-free: every lane of the face-owning block evaluates the Mur formula on its own last elements, lane
            // zlane keeps the results (selects) and the ordinary vector stores carry them (a divergent single-lane
            // section with scalar re-stores measured 1.5 % slower in fp64).
            if constexpr (ZF && !A::COMP) {
                if (zlane >= 0) {     // block-uniform
                    const bool zl = (lane == zlane);
#pragma unroll
                    for (int q = 0; q < RW; ++q) {
                        const T un_ = shfl_up1(uzc[q].v[V - 1]), on_ = shfl_up1(oz[q].v[V - 1]);
                        const T fx = mur<A>(uxc[q].v[V - 2], ox[q].v[V - 2], uxc[q].v[V - 1], p.zf_ct);
                        const T fy = mur<A>(uyc[q].v[V - 2], oy[q].v[V - 2], uyc[q].v[V - 1], p.zf_ct);
                        const T un = (V >= 3) ? uzc[q].v[V >= 3 ? V - 3 : 0] : un_;
                        const T on = (V >= 3) ? oz[q].v[V >= 3 ? V - 3 : 0] : on_;
                        const T fz = mur<A>(un, on, uzc[q].v[V - 2], p.zf_cl);
                        ox[q].v[V - 1] = (zl && n < g.nx - 1) ? fx : ox[q].v[V - 1];
                        oy[q].v[V - 1] = (zl && j + q < g.ny - 1) ? fy : oy[q].v[V - 1];
                        oz[q].v[V - 2] = zl ? fz : oz[q].v[V - 2];
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < RW; ++q) {
                const int offq = off + q * g.nzp;
                const bool st = in_box[q] && row_out[q];
                if constexpr (A::COMP) {
                    if (st) {     // the u_old tile of this plane was read by TMA before: in-place update of delta
                        *reinterpret_cast<PV *>(p.old.ux + offq) = dx[q];
                        *reinterpret_cast<PV *>(p.old.uy + offq) = dy[q];
                        *reinterpret_cast<PV *>(p.old.uz + offq) = dz[q];
                    }
                }
                if (st) {
                    *reinterpret_cast<PV *>(p.nw.ux + offq) = ox[q];
                    *reinterpret_cast<PV *>(p.nw.uy + offq) = oy[q];
                    *reinterpret_cast<PV *>(p.nw.uz + offq) = oz[q];
                }
            }
            // fused halo exchange: the slab's edge planes go straight into the neighbours' ghost planes over NVLink
            if constexpr (PUSH) {
                const bool lo = (n == g.x0 && p.push_lo[0]), hi = (n == g.x0 + g.nxl - 1 && p.push_hi[0]);
                if (lo || hi) {
#pragma unroll
                    for (int q = 0; q < RW; ++q) {
                        if (!(in_box[q] && row_out[q])) continue;
                        const int po = (j + q) * g.nzp + kb;
                        if (lo) {
                            *reinterpret_cast<PV *>(p.push_lo[0] + po) = ox[q];
                            *reinterpret_cast<PV *>(p.push_lo[1] + po) = oy[q];
                            *reinterpret_cast<PV *>(p.push_lo[2] + po) = oz[q];
                        }
                        if (hi) {
                            *reinterpret_cast<PV *>(p.push_hi[0] + po) = ox[q];
                            *reinterpret_cast<PV *>(p.push_hi[1] + po) = oy[q];
                            *reinterpret_cast<PV *>(p.push_hi[2] + po) = oz[q];
                        }
                    }
                    pushed = true;
                }
            }
        }
        // every read this warp makes of the stage holding plane n is done
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(bar_empty + sCi * 8);
#if PHB_REFILL == 0
            // refill every stage whose readers are all done (usually the one just released by the
            // slowest warp of plane `it`)
            for (;;) {
                const int q = *reinterpret_cast<volatile int *>(issued);
                if (q >= nclaims) break;
                if (!mbar_test(bar_empty + (q % NST) * 8, (uint32_t)((q / NST - 1) & 1))) break;
                if (atomicCAS(issued, q, q + 1) == q) issue(q);
            }
#endif
        }

        // (8) rotate
        off += dps;
        sCi = sNi;
        if (++sNi == NST) { sNi = 0; phN ^= 1u; }
#pragma unroll
        for (int q = 0; q < RW; ++q) {
            t3R[q] = t3Rn[q];
#pragma unroll
            for (int e = 0; e < V; ++e) {
                t1c[q][e] = t1n[q][e]; t2c[q][e] = t2n[q][e]; t3c[q][e] = t3n[q][e];
                t5m[q][e] = t5[q][e]; t6m[q][e] = t6[q][e]; rowc[q][e] = rown[q][e];
            }
        }
    }
    // Peer stores are posted writes over NVLink: one system-scope fence per pushing thread at the END of the block
    // (a fence inside the plane loop stalls the warp for an NVLink round trip per row: 40 us per neighbour and step).
    // The flag the neighbour waits on is written by k_signal, stream-ordered after this kernel has completed.
    if constexpr (PUSH) {
        if (pushed) __threadfence_system();
    }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// L2 promotion of the TMA loads (PHB_TMA_PROMO = 0 | 64 | 128 | 256).  Default 64 B: the step is split into launches per
// z-tile class, and where the neighbour tile runs at another time a tile's 16-byte z-halo vector costs a DRAM fetch of
// the promotion size -- 512^3 split step, 256 -> 64 B: fp64 1.627 -> 1.607 ms, fp32 0.883 -> 0.829 ms (one launch for all
// tiles is indifferent to it: 0.842 ms at any size).
inline CUtensorMapL2promotion tma_promotion(int esz) {
    static const int env = getenv("PHB_TMA_PROMO") ? atoi(getenv("PHB_TMA_PROMO")) : -1;
    const int v = env >= 0 ? env : 64;
    (void)esz;
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
         : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}
// 3-D tensor map over a (planes, ny, nzp) box of `esz`-byte elements with box (bx, by, 1).
inline bool make_map3(CUtensorMap *tm, CUtensorMapDataType dt, int esz, void *base, int nzp, int ny, int planes, int bx,
                      int by, int field_esz) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)nzp, (cuuint64_t)ny, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)nzp * esz, (cuuint64_t)nzp * ny * esz};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(tm, dt, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               tma_promotion(field_esz), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// 4-D tensor map over the three components of one displacement buffer (allocated back to back):
// dims (nzp, ny, planes, 3), box (bx, by, 1, 3).
inline bool make_map4(CUtensorMap *tm, CUtensorMapDataType dt, int esz, void *base, int nzp, int ny, int planes, int bx,
                      int by) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)nzp, (cuuint64_t)ny, (cuuint64_t)planes, 3};
    const cuuint64_t strides[3] = {(cuuint64_t)nzp * esz, (cuuint64_t)nzp * ny * esz, (cuuint64_t)nzp * ny * esz * planes};
    const cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, 1u, 3u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return enc(tm, dt, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               tma_promotion(esz), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
template <class T>
inline bool make_field_maps(CUtensorMap *cur_box, CUtensorMap *old_box, void *base, int nzp, int ny, int planes, int R) {
    constexpr int V = VecOf<T>::V;
    const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    return make_map4(cur_box, dt, sizeof(T), base, nzp, ny, planes, 34 * V, R) &&
           make_map4(old_box, dt, sizeof(T), base, nzp, ny, planes, 32 * V, R - 2);
}
template <class T>
inline bool make_class_map(CUtensorMap *tm, void *base, int nzp, int ny, int planes, int R) {
    return make_map3(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, base, nzp, ny, planes, 32 * VecOf<T>::V + 32, R, (int)sizeof(T));
}

template <class T> inline const char *march_name() { return "march_tma"; }

// returns launches made (1), 0 for an empty range, -1 if the shared-memory request is refused
// part: 0 = any z-tiles (with the k = 0 selects); split steps: 1 = z-tile 0 alone, 2 = the last z-tile alone, 3 = tiles in
// between (ztile0 >= 1, not the last) -- each picks the instantiation without what it cannot need
template <class A, int R, int NST, int RW = 1, bool PUSH = false>
inline int launch_march_cfg(const StepArgs<typename A::T> &p, const MatCls<typename A::T> &m, const MarchMaps &maps,
                            int chunks, cudaStream_t st, int ntiles = 0, int part = 0) {
    using T = typename A::T;
    using C_ = MarchCfg<T, R, NST, RW>;
    if constexpr (!PUSH) {
        if (p.push_lo[0] || p.push_hi[0]) return launch_march_cfg<A, R, NST, RW, true>(p, m, maps, chunks, st, ntiles, 0);
    }
    constexpr bool kHasZF = (RW == 2) && !A::COMP;     // instantiations with the fused z face (StepArgs::zface)
    constexpr bool kParts = (RW == 2) && !PUSH && !A::COMP;      // the specialised instantiations of a split step exist
    if (p.zface && !kHasZF) return -3;
    if (!kParts) part = 0;
    if ((part == 1 && (p.ztile0 != 0 || ntiles != 1 || p.zface)) || (part == 2 && ntiles != 1) || (part == 3 && (p.ztile0 < 1 || p.zface))) return -3;
    constexpr int E1 = kParts ? 1 : 0, E2 = kParts ? 2 : 0;
    auto kern = k_step_march<A, R, NST, PUSH, RW, false, true, 0>;
    if (part == 1) kern = k_step_march<A, R, NST, PUSH, RW, false, true, E1>;
    else if (part == 3) kern = k_step_march<A, R, NST, PUSH, RW, false, !kParts, 0>;
    else if (part == 2) kern = p.zface ? k_step_march<A, R, NST, PUSH, RW, kHasZF, !kParts, E2> : k_step_march<A, R, NST, PUSH, RW, false, !kParts, E2>;
    else if (p.zface) kern = k_step_march<A, R, NST, PUSH, RW, kHasZF, true, 0>;
    static size_t attr_by_dev2[8][64] = {};   // per template instantiation and device (the attribute is per device)
    size_t (&attr_by_dev)[64] = attr_by_dev2[(p.zface ? 1 : 0) + 2 * part];
    int dev = 0;
    cudaGetDevice(&dev);
    size_t &attr_bytes = attr_by_dev[dev & 63];
    const int np = p.i_end - p.i_begin;
    if (np <= 0) return 0;
    if (chunks < 1) chunks = 1;
    if (chunks > np) chunks = np;
    const int chunk = (np + chunks - 1) / chunks;
    const size_t smem = C_::smem_bytes(chunk, m.ncls);
    if (smem > attr_bytes) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
        attr_bytes = smem;
    }
    // ntiles > 0: z-tiles [p.ztile0, p.ztile0 + ntiles) only
    dim3 grid(ntiles > 0 ? ntiles : (p.g.nzp + C_::TZ - 1) / C_::TZ, (p.g.ny + C_::TY - 1) / C_::TY, p.edge_b >= 0 ? 2 : (np + chunk - 1) / chunk);
    kern<<<grid, C_::W * 32, smem, st>>>(maps, p, m, chunk);
    return 1;
}

}  // namespace phb

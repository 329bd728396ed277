// k_march.cuh -- the production step kernel for sm_100a.
//
// One fused pass per time step: u_new = f(u_cur stencil, u_old, material); the six stresses
// never touch HBM.  Structure (DESIGN.md "K1"):
//
//  * The reference stores z fastest, so lanes run along z (V cells per lane, 16-byte
//    vectors) and the block MARCHES along x, the slowest axis -- the role "z-marching"
//    plays in the usual 2.5-D blocking.  A block owns a (TY rows x 32*V cells) tile of the
//    (y, z) plane and a chunk of x-planes; warp r of the block is row j0-1+r, so the first
//    and last warps are the y-halo rows (they only compute stresses).
//  * u_cur planes arrive by TMA (cp.async.bulk.tensor.3d, UTMALDG in SASS) into an NST-stage
//    shared-memory ring, box = (34*V, R, 1): one extra 16-byte vector on each side in z gives
//    the z-halo, and out-of-range rows/columns/planes are zero-filled by the TMA unit, so the
//    kernel has no load-side boundary code.  One elected thread issues, everybody waits on
//    the stage's mbarrier.
//  * Per plane n a thread keeps in registers, per cell: u(n), u(n+1) (own position), the
//    normal stresses T1..T3(n) and the shear stresses T5(n-1), T6(n-1) carried from the
//    previous iteration.  Neighbours in y come from shared memory (u from the TMA ring, T2/T4/
//    T6 from a double-buffered exchange tile: ONE __syncthreads per plane); neighbours in z
//    come from warp shuffles, and the two edge lanes rebuild the three halo stresses they need
//    from the ring's halo vectors.
//  * The reference's slice ranges ("never written => 0") are not code here: the per-cell
//    stencil class (1 byte, fd_common.cuh) selects a 16-entry coefficient row in shared memory
//    that is already zero wherever a stress or an update does not exist.
//  * u_old, the class byte and u_new are pure streams (16-byte vector LDG/STG, no halo).
//
// Arithmetic goes through the same formula functions as the naive kernel, so in EXACT mode
// both are bit-identical to the reference.
#pragma once
#include <cuda.h>

#include "k_naive.cuh"

namespace phb {

template <class T> struct VecOf;
template <> struct VecOf<float> { static constexpr int V = 4; };
template <> struct VecOf<double> { static constexpr int V = 2; };

template <class T, int R, int NST>
struct MarchCfg {
    static constexpr int V = VecOf<T>::V;
    static constexpr int TZ = 32 * V;              // cells per tile row
    static constexpr int PITCH = 34 * V;           // ring row pitch in elements (halo vector each side)
    static constexpr int TY = R - 2;               // output rows per tile
    static constexpr int THREADS = R * 32;
    static constexpr int STAGE_ELEMS = 3 * R * PITCH;
    static constexpr size_t STAGE_BYTES = (size_t)STAGE_ELEMS * sizeof(T);
    static constexpr size_t RING_BYTES = NST * STAGE_BYTES;
    static constexpr size_t XCH_BYTES = (size_t)2 * 3 * R * TZ * sizeof(T);
    static constexpr size_t BAR_BYTES = 128;
    static constexpr size_t smem_bytes(int ncls) {
        return RING_BYTES + XCH_BYTES + BAR_BYTES + (size_t)ncls * CLS_W * sizeof(T);
    }
};

// ---- PTX helpers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <class T> __device__ __forceinline__ T shfl_up1(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <class T> __device__ __forceinline__ T shfl_dn1(T v) { return __shfl_down_sync(0xffffffffu, v, 1); }

template <class T, int V> struct alignas(sizeof(T) * V) Pack { T v[V]; };

// ---- the kernel ---------------------------------------------------------------------------------
template <class A, int R, int NST>
__global__ void __launch_bounds__(R * 32, (R <= 8 ? 2 : 1))
k_step_march(const __grid_constant__ CUtensorMap tm_ux, const __grid_constant__ CUtensorMap tm_uy,
             const __grid_constant__ CUtensorMap tm_uz, StepArgs<typename A::T> p, MatCls<typename A::T> m, int chunk) {
    using T = typename A::T;
    using C_ = MarchCfg<T, R, NST>;
    constexpr int V = C_::V, TZ = C_::TZ, PITCH = C_::PITCH, SE = C_::STAGE_ELEMS;
    using PV = Pack<T, V>;
    using PC = Pack<uint8_t, V>;
    const Geo<T> &g = p.g;

    extern __shared__ __align__(1024) unsigned char sm[];
    T *const ring = reinterpret_cast<T *>(sm);                                      // [NST][3][R][PITCH]
    T *const xch = reinterpret_cast<T *>(sm + C_::RING_BYTES);                      // [2][3][R][TZ]
    uint64_t *const bars = reinterpret_cast<uint64_t *>(sm + C_::RING_BYTES + C_::XCH_BYTES);
    T *const stab = reinterpret_cast<T *>(sm + C_::RING_BYTES + C_::XCH_BYTES + C_::BAR_BYTES);   // [ncls][CLS_W]

    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int k0t = blockIdx.x * TZ;                 // first cell of the tile row
    const int j0 = blockIdx.y * C_::TY;              // first output row
    const int ia = p.i_begin + blockIdx.z * chunk;   // planes [ia, ib)
    const int ib = min(ia + chunk, p.i_end);
    if (ia >= ib) return;
    const int j = j0 - 1 + r;
    const int kb = k0t + lane * V;                   // first cell of this lane
    const bool k0c = (kb == 0);                      // element 0 of this thread is the k = 0 plane
    const int rS = max(r - 1, 0), rN = min(r + 1, R - 1);

    // ---- one-time setup: class table to smem, barriers ----
    for (int q = threadIdx.x; q < m.ncls * CLS_W; q += R * 32) stab[q] = m.tab[q];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // planes are consumed in order q = 0, 1, ...; plane n = ia - 1 + q; local plane l = n - x0 + 1
    const int nplanes = (ib - ia) + 2;
    const CUtensorMap *pm0 = &tm_ux, *pm1 = &tm_uy, *pm2 = &tm_uz;
    const int lbase = (ia - 1) - g.x0 + 1;
    auto issue = [=](int q) {
        const int s = q % NST;
        T *dst = ring + s * SE;
        mbar_expect_tx(&bars[s], (uint32_t)C_::STAGE_BYTES);
        tma_load_3d(dst + 0 * R * PITCH, pm0, &bars[s], k0t - V, j0 - 1, lbase + q);
        tma_load_3d(dst + 1 * R * PITCH, pm1, &bars[s], k0t - V, j0 - 1, lbase + q);
        tma_load_3d(dst + 2 * R * PITCH, pm2, &bars[s], k0t - V, j0 - 1, lbase + q);
    };
    if (threadIdx.x == 0) {
        for (int q = 0; q < NST && q < nplanes; ++q) issue(q);
    }

    // ---- loop-invariant per-thread quantities ----
    const bool in_box = (j >= 0 && j < g.ny && kb < g.nzp);      // global vector accesses allowed
    const bool row_out = (r >= 1 && r <= R - 2) && (j < g.ny);    // this WARP produces u_new (warp-uniform:
                                                                  // the shuffles below need converged warps)
    // spacings (or reciprocals).  On the k = 0 plane the reference uses the FIRST element of each
    // spacing array, on "wrong" axes in the shear terms (App. B #1, #2): folded here into the
    // values element 0 of the k0c thread uses, so the plane loop has no k = 0 code path except
    // the per-plane x spacings below.
    T sfz[V], ssz[V];                                             // fdz[k], sdz[k-1]
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const int kc = min(kb + e, g.nz);                         // tables are addressable on [-1, n]
        sfz[e] = g.fdz[kc];
        ssz[e] = g.sdz[kc - 1];
    }
    const int jc = min(max(j, 0), g.ny);
    const T sfy = g.fdy[jc], ssy = g.sdy[jc - 1];
    if (k0c) ssz[0] = g.sdz0;                                     // normal z-term and ux/uy z-term: sdz[0]
    const T e0_ssy = k0c ? g.sdy0 : ssy;                          // sdy[0] instead of sdy[j-1]
    const T e0_t4a = k0c ? g.fdy0 : sfz[0];                       // T4: (uy[k+1]-uy[k]) / fdy[0]
    const T e0_t4b = k0c ? g.fdz0 : sfy;                          // T4: (uz[j+1]-uz[j]) / fdz[0]
    const T e0_t5a = k0c ? g.fdx0 : sfz[0];                       // T5: (ux[k+1]-ux[k]) / fdx[0]
    const T e0_t6a = k0c ? g.fdx0 : sfy;                          // T6: (ux[j+1]-ux[j]) / fdx[0]
    const T e0_uyb = k0c ? g.fdy0 : sfy;                          // uy: (T2[j+1]-T2[j]) / fdy[0]
    const T e0_uzc = k0c ? (T)1 : sfz[0];                         // uz: T3[..,1] is NOT divided (App. B #3)
    // halo columns of the edge lanes (never the k = 0 plane)
    const int kL = k0t - 1, kR = k0t + TZ;
    const bool haveL = (lane == 0) && (kL >= 0) && in_box;
    const bool haveR = (lane == 31) && (kR <= g.nz - 1) && in_box;
    const T sfzL = haveL ? g.fdz[kL] : (T)1, sszR = haveR ? g.sdz[kR - 1] : (T)1;

    const T *const own0 = ring + r * PITCH + (lane + 1) * V;      // own vector inside a component tile
    const T *const ownS = ring + rS * PITCH + (lane + 1) * V;
    const T *const ownN = ring + rN * PITCH + (lane + 1) * V;
    const T *const rowp = ring + r * PITCH, *const rowS = ring + rS * PITCH, *const rowN = ring + rN * PITCH;
    auto vec = [](const T *q) -> PV { return *reinterpret_cast<const PV *>(q); };

    // ---- registers carried across planes ----
    PV uxc, uyc, uzc;            // u(n) own
    T t1c[V], t2c[V], t3c[V];    // T1..T3(n)
    T t5m[V], t6m[V];            // T5(n-1), T6(n-1)
    T t3R = (T)0;                // T3(n, j, kR) for lane 31
    PC codec;                    // class(n)
#pragma unroll
    for (int e = 0; e < V; ++e) { t1c[e] = t2c[e] = t3c[e] = t5m[e] = t6m[e] = (T)0; codec.v[e] = 0; }

    // running element offset of (plane n, row j, cell kb)
    long long off = (long long)lbase * g.ps + (long long)j * g.nzp + kb;

    // plane q = 0 (n = ia - 1): own values
    mbar_wait(&bars[0], 0);
    uxc = vec(own0); uyc = vec(own0 + R * PITCH); uzc = vec(own0 + 2 * R * PITCH);
    if (in_box) codec = *reinterpret_cast<const PC *>(m.code + off);

    for (int it = 0; it + 1 < nplanes; ++it) {
        const int n = ia - 1 + it;                    // plane being completed; n + 1 is the newest
        const bool emit = (it >= 1);                  // it = 0 only primes the carried stresses
        const int sC = (it % NST) * SE, sN = ((it + 1) % NST) * SE;

        // (0) streaming operands: class bytes of plane n+1, u_old(n), halo class bytes
        PV uox, uoy, uoz;
        PC coden;
#pragma unroll
        for (int e = 0; e < V; ++e) { uox.v[e] = uoy.v[e] = uoz.v[e] = (T)0; coden.v[e] = 0; }
        uint8_t clsL = 0, clsR = 0;
        if (in_box) {
            coden = *reinterpret_cast<const PC *>(m.code + off + g.ps);
            if (emit && row_out) {   // in_box holds here
                uox = *reinterpret_cast<const PV *>(p.old.ux + off);
                uoy = *reinterpret_cast<const PV *>(p.old.uy + off);
                uoz = *reinterpret_cast<const PV *>(p.old.uz + off);
            }
            if (haveL) clsL = m.code[off - 1];
            if (haveR) clsR = m.code[off + g.ps + V];
        }

        // per-plane x spacings (uniform); element 0 of the k0c thread uses the first elements
        const T sfx_n = g.fdx[n], ssx_n = g.sdx[n], ssx_m = g.sdx[max(n - 1, -1)];
        const T e0_ssx_n = k0c ? g.sdx0 : ssx_n;     // normal(n+1): sdx[0]
        const T e0_ssx_m = k0c ? g.sdx0 : ssx_m;     // uy, uz: sdx[0]
        const T e0_sfx_n = k0c ? g.fdx0 : sfx_n;     // ux: fdx[0]
        const T e0_shb = k0c ? g.fdz0 : sfx_n;       // T5, T6 second term: fdz[0]

        // (1) newest plane
        mbar_wait(&bars[(it + 1) % NST], ((it + 1) / NST) & 1);
        const PV uxn = vec(own0 + sN), uyn = vec(own0 + sN + R * PITCH), uzn = vec(own0 + sN + 2 * R * PITCH);

        // (2) normal stresses at plane n + 1
        const PV uyS = vec(ownS + sN + R * PITCH);                // uy(n+1, j-1)
        T uzW0 = shfl_up1(uzn.v[V - 1]);                          // uz(n+1, k-1) for element 0
        if (lane == 0) uzW0 = rowp[sN + 2 * R * PITCH + V - 1];   // ring halo (zero-filled below k = 0)
        T t1n[V], t2n[V], t3n[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const T dxx = A::sub(uxn.v[e], uxc.v[e]);
            const T dyy = A::sub(uyn.v[e], uyS.v[e]);
            const T dzz = A::sub(uzn.v[e], (e == 0) ? uzW0 : uzn.v[e > 0 ? e - 1 : 0]);
            const T sx = (e == 0) ? e0_ssx_n : ssx_n, sy = (e == 0) ? e0_ssy : ssy;
            const T *c = stab + (int)coden.v[e] * CLS_W;
            t1n[e] = normal_row<A>(c + 0, dxx, dyy, dzz, sx, sy, ssz[e]);
            t2n[e] = normal_row<A>(c + 3, dxx, dyy, dzz, sx, sy, ssz[e]);
            t3n[e] = normal_row<A>(c + 6, dxx, dyy, dzz, sx, sy, ssz[e]);
        }

        // (3) shear stresses at plane n
        const PV uzN = vec(ownN + sC + 2 * R * PITCH), uxN = vec(ownN + sC);   // uz(n, j+1), ux(n, j+1)
        T uyE = shfl_dn1(uyc.v[0]), uxE = shfl_dn1(uxc.v[0]);                  // uy(n, k+1), ux(n, k+1) for element V-1
        if (lane == 31) { uyE = rowp[sC + R * PITCH + 33 * V]; uxE = rowp[sC + 33 * V]; }
        T t4[V], t5[V], t6[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const T uye = (e == V - 1) ? uyE : uyc.v[e < V - 1 ? e + 1 : e];
            const T uxe = (e == V - 1) ? uxE : uxc.v[e < V - 1 ? e + 1 : e];
            const T *c = stab + (int)codec.v[e] * CLS_W;
            const T shb = (e == 0) ? e0_shb : sfx_n;
            t4[e] = shear<A>(c[CLS_C44], A::sub(uye, uyc.v[e]), (e == 0) ? e0_t4a : sfz[e],
                             A::sub(uzN.v[e], uzc.v[e]), (e == 0) ? e0_t4b : sfy);
            t5[e] = shear<A>(c[CLS_C55], A::sub(uxe, uxc.v[e]), (e == 0) ? e0_t5a : sfz[e],
                             A::sub(uzn.v[e], uzc.v[e]), shb);
            t6[e] = shear<A>(c[CLS_C66], A::sub(uxN.v[e], uxc.v[e]), (e == 0) ? e0_t6a : sfy,
                             A::sub(uyn.v[e], uyc.v[e]), shb);
        }

        // (4) halo stresses of the edge lanes: columns kL = k0t-1 (T4, T5 at plane n) and
        //     kR = k0t+TZ (T3 at plane n+1)
        T t4L = (T)0, t5L = (T)0, t3Rn = (T)0;
        if (haveL) {
            const T *c = stab + (int)clsL * CLS_W;
            const T uyL = rowp[sC + R * PITCH + V - 1], uzL = rowp[sC + 2 * R * PITCH + V - 1], uxL = rowp[sC + V - 1];
            const T uzLN = rowN[sC + 2 * R * PITCH + V - 1], uzLn = rowp[sN + 2 * R * PITCH + V - 1];
            t4L = shear<A>(c[CLS_C44], A::sub(uyc.v[0], uyL), sfzL, A::sub(uzLN, uzL), sfy);
            t5L = shear<A>(c[CLS_C55], A::sub(uxc.v[0], uxL), sfzL, A::sub(uzLn, uzL), sfx_n);
        }
        if (haveR) {
            const T *c = stab + (int)clsR * CLS_W;
            const T dxx = A::sub(rowp[sN + 33 * V], rowp[sC + 33 * V]);
            const T dyy = A::sub(rowp[sN + R * PITCH + 33 * V], rowS[sN + R * PITCH + 33 * V]);
            const T dzz = A::sub(rowp[sN + 2 * R * PITCH + 33 * V], uzn.v[V - 1]);
            t3Rn = normal_row<A>(c + 6, dxx, dyy, dzz, ssx_n, ssy, sszR);
        }

        // (5) publish T2(n), T4(n), T6(n) for the y-neighbours; one barrier per plane
        T *const xb = xch + (it & 1) * (3 * R * TZ) + lane * V;
        {
            PV a, b, c;
#pragma unroll
            for (int e = 0; e < V; ++e) { a.v[e] = t2c[e]; b.v[e] = t4[e]; c.v[e] = t6[e]; }
            *reinterpret_cast<PV *>(xb + (0 * R + r) * TZ) = a;
            *reinterpret_cast<PV *>(xb + (1 * R + r) * TZ) = b;
            *reinterpret_cast<PV *>(xb + (2 * R + r) * TZ) = c;
        }
        __syncthreads();
        // every read of ring stage it % NST is done: refill it with plane it + NST
        if (threadIdx.x == 0 && it + NST < nplanes) issue(it + NST);

        if (emit && row_out) {
            // (6) y-neighbour stresses
            const PV t2N = vec(xb + (0 * R + rN) * TZ);
            const PV t4S = vec(xb + (1 * R + rS) * TZ);
            const PV t6S = vec(xb + (2 * R + rS) * TZ);
            // (7) z-neighbour stresses
            T t3U = shfl_dn1(t3c[0]);                 // T3(n, k+1) for element V-1
            T t4W = shfl_up1(t4[V - 1]);              // T4(n, k-1) for element 0
            T t5W = shfl_up1(t5[V - 1]);              // T5(n, k-1) for element 0
            if (lane == 31) t3U = t3R;
            if (lane == 0) { t4W = t4L; t5W = t5L; }  // zero below the k = 0 plane
            // (8) displacement update of plane n
            PV ox, oy, oz;
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const T *c = stab + (int)codec.v[e] * CLS_W;
                const T t3u = (e == V - 1) ? t3U : t3c[e < V - 1 ? e + 1 : e];
                const T t4w = (e == 0) ? t4W : t4[e > 0 ? e - 1 : 0];
                const T t5w = (e == 0) ? t5W : t5[e > 0 ? e - 1 : 0];
                const T sy_s = (e == 0) ? e0_ssy : ssy;
                {   // ux
                    const T acc = A::add(A::add(A::scl(A::sub(t1n[e], t1c[e]), (e == 0) ? e0_sfx_n : sfx_n),
                                                A::scl(A::sub(t6[e], t6S.v[e]), sy_s)),
                                         A::scl(A::sub(t5[e], t5w), ssz[e]));
                    ox.v[e] = advance<A>(uxc.v[e], uox.v[e], c[CLS_RX], acc);
                }
                {   // uy
                    const T acc = A::add(A::add(A::scl(A::sub(t6[e], t6m[e]), (e == 0) ? e0_ssx_m : ssx_m),
                                                A::scl(A::sub(t2N.v[e], t2c[e]), (e == 0) ? e0_uyb : sfy)),
                                         A::scl(A::sub(t4[e], t4w), ssz[e]));
                    oy.v[e] = advance<A>(uyc.v[e], uoy.v[e], c[CLS_RY], acc);
                }
                {   // uz
                    const T acc = A::add(A::add(A::scl(A::sub(t5[e], t5m[e]), (e == 0) ? e0_ssx_m : ssx_m),
                                                A::scl(A::sub(t4[e], t4S.v[e]), sy_s)),
                                         A::scl(A::sub(t3u, t3c[e]), (e == 0) ? e0_uzc : sfz[e]));
                    oz.v[e] = advance<A>(uzc.v[e], uoz.v[e], c[CLS_RZ], acc);
                }
            }
            // i = 0: uy, uz keep u_new == u (App. B #9); uz(0, j, 0) is the pre-source value
            if (n == 0) {
                oy = uyc;
                oz = uzc;
                if (k0c && p.line_save) oz.v[0] = p.line_save[j];
            }
            if (in_box) {
                *reinterpret_cast<PV *>(p.nw.ux + off) = ox;
                *reinterpret_cast<PV *>(p.nw.uy + off) = oy;
                *reinterpret_cast<PV *>(p.nw.uz + off) = oz;
            }
        }

        // (9) rotate the carried registers
        uxc = uxn; uyc = uyn; uzc = uzn;
        codec = coden;
        t3R = t3Rn;
        off += g.ps;
#pragma unroll
        for (int e = 0; e < V; ++e) { t1c[e] = t1n[e]; t2c[e] = t2n[e]; t3c[e] = t3n[e]; t5m[e] = t5[e]; t6m[e] = t6[e]; }
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

// Tensor map over one displacement component: dims (nzp, ny, planes), box (34 V, R, 1).
template <class T>
inline bool make_field_map(CUtensorMap *tm, void *base, int nzp, int ny, int planes, int R) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    constexpr int V = VecOf<T>::V;
    const cuuint64_t dims[3] = {(cuuint64_t)nzp, (cuuint64_t)ny, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)nzp * sizeof(T), (cuuint64_t)nzp * ny * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)(34 * V), (cuuint32_t)R, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    return enc(tm, dt, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class T> inline const char *march_name() { return "march_tma"; }

template <class A, int R, int NST>
inline int launch_march_cfg(const StepArgs<typename A::T> &p, const MatCls<typename A::T> &m, const CUtensorMap *maps,
                            int chunks, cudaStream_t st) {
    using T = typename A::T;
    using C_ = MarchCfg<T, R, NST>;
    auto kern = k_step_march<A, R, NST>;
    static size_t attr_bytes = 0;   // per template instantiation
    const size_t smem = C_::smem_bytes(m.ncls);
    if (smem > attr_bytes) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        attr_bytes = smem;
    }
    const int np = p.i_end - p.i_begin;
    if (np <= 0) return 0;
    if (chunks < 1) chunks = 1;
    if (chunks > np) chunks = np;
    const int chunk = (np + chunks - 1) / chunks;
    dim3 grid((p.g.nzp + C_::TZ - 1) / C_::TZ, (p.g.ny + C_::TY - 1) / C_::TY, (np + chunk - 1) / chunk);
    kern<<<grid, R * 32, smem, st>>>(maps[0], maps[1], maps[2], p, m, chunk);
    return 1;
}

}  // namespace phb

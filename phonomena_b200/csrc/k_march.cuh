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
//  * u_cur planes arrive by TMA (cp.async.bulk.tensor.3d, UBLKCP/UTMALDG in SASS) into an
//    NST-stage shared-memory ring, box = (34*V, R, 1): one extra 16-byte vector on each side in
//    z gives the z-halo, out-of-range rows/columns/planes are zero-filled by the TMA unit, so
//    the kernel has no load-side boundary code.  One elected thread issues, everybody waits on
//    the stage's mbarrier.
//  * Per plane n a thread keeps in registers, per cell: u(n), u(n+1) (own position), the
//    normal stresses T1..T3(n) and the shear stresses T5(n-1), T6(n-1) carried from the
//    previous iteration.  Neighbours in y come from shared memory (u from the TMA ring, T2/T4/
//    T6 from a double-buffered exchange tile: ONE __syncthreads per plane); neighbours in z
//    come from warp shuffles, and the two edge lanes rebuild the three halo stresses they need
//    from the ring's halo vectors.
//  * u_old, the material code and u_new are pure streams (16-byte vector LDG/STG, no halo).
//
// Arithmetic goes through the same formula functions as the naive kernel, so in EXACT mode
// both are bit-identical to the reference.
#pragma once
#include <cuda.h>

#include "k_naive.cuh"

namespace phb {

template <class T> struct VecOf;
template <> struct VecOf<float> { static constexpr int V = 4; using type = float4; };
template <> struct VecOf<double> { static constexpr int V = 2; using type = double2; };

template <class T, int R, int NST>
struct MarchCfg {
    static constexpr int V = VecOf<T>::V;
    static constexpr int TZ = 32 * V;              // cells per tile row
    static constexpr int PITCH = 34 * V;           // ring row pitch in elements (halo vector each side)
    static constexpr int TY = R - 2;               // output rows per tile
    static constexpr int THREADS = R * 32;
    static constexpr size_t STAGE_COMP_BYTES = (size_t)R * PITCH * sizeof(T);
    static constexpr size_t STAGE_BYTES = 3 * STAGE_COMP_BYTES;
    static constexpr size_t RING_BYTES = NST * STAGE_BYTES;
    static constexpr size_t XCH_BYTES = (size_t)2 * 3 * R * TZ * sizeof(T);
    static constexpr size_t TAB_BYTES = (size_t)MAX_MAT * TAB_W * sizeof(T);
    static constexpr size_t SMEM_BYTES = RING_BYTES + XCH_BYTES + TAB_BYTES + 64 /*mbarriers*/ + 128 /*align*/;
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
template <class A, class M, int R, int NST>
__global__ void __launch_bounds__(R * 32, (R <= 8 ? 2 : 1))
k_step_march(const __grid_constant__ CUtensorMap tm_ux, const __grid_constant__ CUtensorMap tm_uy,
             const __grid_constant__ CUtensorMap tm_uz, StepArgs<typename A::T> p, M m, int chunk, int nmat) {
    using T = typename A::T;
    using CodeT = typename M::CodeT;
    using C_ = MarchCfg<T, R, NST>;
    constexpr int V = C_::V, TZ = C_::TZ, PITCH = C_::PITCH;
    using PV = Pack<T, V>;
    using PC = Pack<CodeT, V>;
    const Geo<T> &g = p.g;

    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    T *ring = (T *)sm;                                           // [NST][3][R][PITCH]
    T *xch = (T *)(sm + C_::RING_BYTES);                         // [2][3][R][TZ]
    T *stab = (T *)(sm + C_::RING_BYTES + C_::XCH_BYTES);        // [nmat][TAB_W]
    uint64_t *bars = (uint64_t *)(sm + C_::RING_BYTES + C_::XCH_BYTES + C_::TAB_BYTES);

    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int k0t = blockIdx.x * TZ;                 // first cell of the tile row
    const int j0 = blockIdx.y * C_::TY;              // first output row
    const int ia = p.i_begin + blockIdx.z * chunk;   // planes [ia, ib)
    const int ib = min(ia + chunk, p.i_end);
    if (ia >= ib) return;
    const int j = j0 - 1 + r;
    const int kb = k0t + lane * V;                   // first cell of this lane
    const bool first_cell = (kb == 0);               // element 0 is the k = 0 plane
    const int rS = max(r - 1, 0), rN = min(r + 1, R - 1);

    // ---- one-time setup: material table to smem, barriers ----
    for (int q = threadIdx.x; q < nmat * TAB_W; q += R * 32) stab[q] = m.tab[q];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // planes are consumed in order q = 0, 1, ...; plane index n = ia - 1 + q; local plane l = n - x0 + 1
    const int nplanes = (ib - ia) + 2;
    const CUtensorMap *pm0 = &tm_ux, *pm1 = &tm_uy, *pm2 = &tm_uz;
    const int lbase = (ia - 1) - g.x0 + 1;
    auto issue = [=](int q) {
        const int s = q % NST;
        T *dst = ring + (size_t)s * 3 * R * PITCH;
        mbar_expect_tx(&bars[s], (uint32_t)C_::STAGE_BYTES);
        const int l = lbase + q;
        tma_load_3d(dst + 0 * R * PITCH, pm0, &bars[s], k0t - V, j0 - 1, l);
        tma_load_3d(dst + 1 * R * PITCH, pm1, &bars[s], k0t - V, j0 - 1, l);
        tma_load_3d(dst + 2 * R * PITCH, pm2, &bars[s], k0t - V, j0 - 1, l);
    };
    if (threadIdx.x == 0) {
        for (int q = 0; q < NST && q < nplanes; ++q) issue(q);
    }

    // ---- loop-invariant per-thread quantities ----
    const bool row_out = (r >= 1 && r <= R - 2 && j < g.ny);     // this thread writes u_new
    const bool vj_n = (j >= 1 && j <= g.ny - 2);                  // T1..T3, T5 rows
    const bool vj_s = (j >= 0 && j <= g.ny - 2);                  // T4, T6 rows
    bool vk[V];
    T sfz[V], ssz[V];                                             // fdz[k], sdz[k-1]
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const int k = kb + e;
        vk[e] = (k <= g.nz - 2);
        const int kc = min(k, g.nz);                              // tables are addressable on [-1, n]
        sfz[e] = g.fdz[kc];
        ssz[e] = g.sdz[kc - 1];
    }
    const T sfy = g.fdy[min(max(j, -1), g.ny)], ssy = g.sdy[min(max(j - 1, -1), g.ny)];
    // halo columns of the edge lanes
    const int kL = k0t - 1, kR = k0t + TZ;
    const bool haveL = (lane == 0) && (kL >= 0);
    const bool haveR = (lane == 31) && (kR <= g.nz - 2);
    const T sfzL = haveL ? g.fdz[kL] : (T)1, sszR = haveR ? g.sdz[kR - 1] : (T)1;

    auto ld_own = [&](const T *stage, int comp, int row) -> PV {
        return *reinterpret_cast<const PV *>(stage + ((size_t)comp * R + row) * PITCH + (lane + 1) * V);
    };
    auto ld_halo = [&](const T *stage, int comp, int row, int col /*element index in the ring row*/) -> T {
        return stage[((size_t)comp * R + row) * PITCH + col];
    };

    // ---- registers carried across planes ----
    PV uxc, uyc, uzc;            // u(n) own
    T t1c[V], t2c[V], t3c[V];    // T1..T3(n)
    T t5m[V], t6m[V];            // T5(n-1), T6(n-1)
    T t3R = (T)0;                // T3(n, j, kR) for lane 31
    PC codec;                    // code(n)
#pragma unroll
    for (int e = 0; e < V; ++e) { t1c[e] = t2c[e] = t3c[e] = t5m[e] = t6m[e] = (T)0; codec.v[e] = 0; }

    const long long rowoff = (long long)j * g.nzp + kb;           // offset inside a plane
    const bool in_box = (j >= 0 && j < g.ny && kb < g.nzp);       // global vector accesses allowed

    // plane q = 0 (n = ia - 1): own values
    mbar_wait(&bars[0], 0);
    {
        const T *st = ring;
        uxc = ld_own(st, 0, r); uyc = ld_own(st, 1, r); uzc = ld_own(st, 2, r);
    }
    if (in_box) codec = *reinterpret_cast<const PC *>(m.code + (long long)lbase * g.ps + rowoff);

    for (int it = 0; it + 1 < nplanes; ++it) {
        const int n = ia - 1 + it;                    // plane being completed; n + 1 is the newest
        const bool emit = (it >= 1);                  // it = 0 only primes the carried stresses
        const T *stC = ring + (size_t)(it % NST) * 3 * R * PITCH;
        const T *stN = ring + (size_t)((it + 1) % NST) * 3 * R * PITCH;
        const long long pl_n = ((long long)n - g.x0 + 1) * g.ps + rowoff;

        // (0) streaming operands of the output stage
        PV uox, uoy, uoz;
        PC coden;
#pragma unroll
        for (int e = 0; e < V; ++e) { uox.v[e] = uoy.v[e] = uoz.v[e] = (T)0; coden.v[e] = 0; }
        if (in_box) {
            coden = *reinterpret_cast<const PC *>(m.code + pl_n + g.ps);
            if (emit && row_out) {
                uox = *reinterpret_cast<const PV *>(p.old.ux + pl_n);
                uoy = *reinterpret_cast<const PV *>(p.old.uy + pl_n);
                uoz = *reinterpret_cast<const PV *>(p.old.uz + pl_n);
            }
        }

        // (1) newest plane
        mbar_wait(&bars[(it + 1) % NST], ((it + 1) / NST) & 1);
        const PV uxn = ld_own(stN, 0, r), uyn = ld_own(stN, 1, r), uzn = ld_own(stN, 2, r);

        // per-plane spacings (uniform)
        const T sfx_n = g.fdx[n], ssx_n = g.sdx[n];          // fdx[n], sdx[(n+1)-1]
        const T ssx_m = g.sdx[max(n - 1, -1)];               // sdx[n-1]
        const bool vi_nn = (n + 1 >= 1 && n + 1 <= g.nx - 2);   // normal stresses at plane n+1
        const bool vi_n4 = (n >= 1 && n <= g.nx - 2);           // T4 at plane n
        const bool vi_s = (n >= 0 && n <= g.nx - 2);            // T5, T6 at plane n

        // (2) normal stresses at plane n + 1
        const PV uyS = ld_own(stN, 1, rS);                   // uy(n+1, j-1)
        T uzW0 = shfl_up1(uzn.v[V - 1]);                     // uz(n+1, k-1) for element 0
        if (lane == 0) uzW0 = ld_halo(stN, 2, r, V - 1);
        T t1n[V], t2n[V], t3n[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const bool k0 = (e == 0) && first_cell;
            const T dxx = A::sub(uxn.v[e], uxc.v[e]);
            const T dyy = A::sub(uyn.v[e], uyS.v[e]);
            const T uzw = (e == 0) ? uzW0 : uzn.v[e > 0 ? e - 1 : 0];
            const T dzz = k0 ? uzn.v[e] : A::sub(uzn.v[e], uzw);
            const T sx = k0 ? g.sdx0 : ssx_n, sy = k0 ? g.sdy0 : ssy, sz = k0 ? g.sdz0 : ssz[e];
            const T *c = stab + code_field<CodeT, M::B>(coden.v[e], F_NODE) * TAB_W;
            const bool ok = vi_nn && vj_n && vk[e];
            t1n[e] = ok ? normal_row<A>(c + 0, dxx, dyy, dzz, sx, sy, sz) : (T)0;
            t2n[e] = ok ? normal_row<A>(c + 3, dxx, dyy, dzz, sx, sy, sz) : (T)0;
            t3n[e] = (ok && !k0) ? normal_row<A>(c + 6, dxx, dyy, dzz, sx, sy, sz) : (T)0;
        }

        // (3) shear stresses at plane n
        const PV uzN = ld_own(stC, 2, rN), uxN = ld_own(stC, 0, rN);    // uz(n, j+1), ux(n, j+1)
        T uyE = shfl_dn1(uyc.v[0]), uxE = shfl_dn1(uxc.v[0]);           // uy(n, k+1), ux(n, k+1) for element V-1
        if (lane == 31) { uyE = ld_halo(stC, 1, r, 33 * V); uxE = ld_halo(stC, 0, r, 33 * V); }
        T t4[V], t5[V], t6[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const bool k0 = (e == 0) && first_cell;
            const T uye = (e == V - 1) ? uyE : uyc.v[e < V - 1 ? e + 1 : e];
            const T uxe = (e == V - 1) ? uxE : uxc.v[e < V - 1 ? e + 1 : e];
            const CodeT cd = codec.v[e];
            {
                const T a = A::sub(uye, uyc.v[e]), b = A::sub(uzN.v[e], uzc.v[e]);
                const T c44 = stab[code_field<CodeT, M::B>(cd, F_T4) * TAB_W + TAB_C44];
                const T v = shear<A>(c44, a, k0 ? g.fdy0 : sfz[e], b, k0 ? g.fdz0 : sfy);
                t4[e] = (vi_n4 && vj_s && vk[e]) ? v : (T)0;
            }
            {
                const T a = A::sub(uxe, uxc.v[e]), b = A::sub(uzn.v[e], uzc.v[e]);
                const T c55 = stab[code_field<CodeT, M::B>(cd, F_T5) * TAB_W + TAB_C55];
                const T v = shear<A>(c55, a, k0 ? g.fdx0 : sfz[e], b, k0 ? g.fdz0 : sfx_n);
                t5[e] = (vi_s && vj_n && vk[e]) ? v : (T)0;
            }
            {
                const T a = A::sub(uxN.v[e], uxc.v[e]), b = A::sub(uyn.v[e], uyc.v[e]);
                const T c66 = stab[code_field<CodeT, M::B>(cd, F_T6) * TAB_W + TAB_C66];
                const T v = shear<A>(c66, a, k0 ? g.fdx0 : sfy, b, k0 ? g.fdz0 : sfx_n);
                t6[e] = (vi_s && vj_s && vk[e]) ? v : (T)0;
            }
        }

        // (4) halo stresses of the edge lanes (columns kL = k0t-1 and kR = k0t+TZ; never k = 0)
        T t4L = (T)0, t5L = (T)0, t3Rn = (T)0;
        if (haveL && j >= 0 && j < g.ny) {
            const CodeT cd = m.code[pl_n - rowoff + (long long)j * g.nzp + kL];
            const T uyL = ld_halo(stC, 1, r, V - 1), uzL = ld_halo(stC, 2, r, V - 1), uxL = ld_halo(stC, 0, r, V - 1);
            const T uzLN = ld_halo(stC, 2, rN, V - 1), uzLn = ld_halo(stN, 2, r, V - 1);
            if (vi_n4 && vj_s) {
                const T c44 = stab[code_field<CodeT, M::B>(cd, F_T4) * TAB_W + TAB_C44];
                t4L = shear<A>(c44, A::sub(uyc.v[0], uyL), sfzL, A::sub(uzLN, uzL), sfy);
            }
            if (vi_s && vj_n) {
                const T c55 = stab[code_field<CodeT, M::B>(cd, F_T5) * TAB_W + TAB_C55];
                t5L = shear<A>(c55, A::sub(uxc.v[0], uxL), sfzL, A::sub(uzLn, uzL), sfx_n);
            }
        }
        if (haveR && vi_nn && vj_n) {
            const CodeT cd = m.code[pl_n + g.ps - rowoff + (long long)j * g.nzp + kR];
            const T *c = stab + code_field<CodeT, M::B>(cd, F_NODE) * TAB_W;
            const T dxx = A::sub(ld_halo(stN, 0, r, 33 * V), ld_halo(stC, 0, r, 33 * V));
            const T dyy = A::sub(ld_halo(stN, 1, r, 33 * V), ld_halo(stN, 1, rS, 33 * V));
            const T dzz = A::sub(ld_halo(stN, 2, r, 33 * V), uzn.v[V - 1]);
            t3Rn = normal_row<A>(c + 6, dxx, dyy, dzz, ssx_n, ssy, sszR);
        }

        // (5) publish T2(n), T4(n), T6(n) for the y-neighbours; one barrier per plane
        T *xb = xch + (size_t)(it & 1) * 3 * R * TZ;
        {
            PV a, b, c;
#pragma unroll
            for (int e = 0; e < V; ++e) { a.v[e] = t2c[e]; b.v[e] = t4[e]; c.v[e] = t6[e]; }
            *reinterpret_cast<PV *>(xb + ((size_t)0 * R + r) * TZ + lane * V) = a;
            *reinterpret_cast<PV *>(xb + ((size_t)1 * R + r) * TZ + lane * V) = b;
            *reinterpret_cast<PV *>(xb + ((size_t)2 * R + r) * TZ + lane * V) = c;
        }
        __syncthreads();
        // every read of ring stage it % NST is done: refill it with plane it + NST
        if (threadIdx.x == 0 && it + NST < nplanes) issue(it + NST);

        if (emit && row_out) {
            // (6) y-neighbour stresses
            const PV t2N = *reinterpret_cast<const PV *>(xb + ((size_t)0 * R + rN) * TZ + lane * V);
            const PV t4S = *reinterpret_cast<const PV *>(xb + ((size_t)1 * R + rS) * TZ + lane * V);
            const PV t6S = *reinterpret_cast<const PV *>(xb + ((size_t)2 * R + rS) * TZ + lane * V);
            // (7) z-neighbour stresses
            T t3U = shfl_dn1(t3c[0]);                 // T3(n, k+1) for element V-1
            T t4W = shfl_up1(t4[V - 1]);              // T4(n, k-1) for element 0
            T t5W = shfl_up1(t5[V - 1]);              // T5(n, k-1) for element 0
            if (lane == 31) t3U = t3R;
            if (lane == 0) { t4W = t4L; t5W = t5L; }
            // (8) displacement update of plane n
            PV ox, oy, oz;
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const bool k0 = (e == 0) && first_cell;
                const CodeT cd = codec.v[e];
                const T t3u = (e == V - 1) ? t3U : t3c[e < V - 1 ? e + 1 : e];
                const T t4w = (e == 0) ? t4W : t4[e > 0 ? e - 1 : 0];
                const T t5w = (e == 0) ? t5W : t5[e > 0 ? e - 1 : 0];
                const T sz_s = k0 ? g.sdz0 : ssz[e];
                {   // ux
                    const T dA = A::sub(t1n[e], t1c[e]);
                    const T dB = A::sub(t6[e], t6S.v[e]);
                    const T dC = k0 ? t5[e] : A::sub(t5[e], t5w);
                    const T acc = A::add(A::add(A::scl(dA, k0 ? g.fdx0 : sfx_n), A::scl(dB, k0 ? g.sdy0 : ssy)),
                                         A::scl(dC, sz_s));
                    const T rinv = stab[code_field<CodeT, M::B>(cd, F_RX) * TAB_W + TAB_RINV];
                    const bool ok = (n <= g.nx - 2) && vj_n && vk[e];
                    ox.v[e] = ok ? advance<A>(uxc.v[e], uox.v[e], rinv, acc) : (T)0;
                }
                {   // uy
                    const T dA = A::sub(t6[e], t6m[e]);
                    const T dB = A::sub(t2N.v[e], t2c[e]);
                    const T dC = k0 ? t4[e] : A::sub(t4[e], t4w);
                    const T acc = A::add(A::add(A::scl(dA, k0 ? g.sdx0 : ssx_m), A::scl(dB, k0 ? g.fdy0 : sfy)),
                                         A::scl(dC, sz_s));
                    const T rinv = stab[code_field<CodeT, M::B>(cd, F_RY) * TAB_W + TAB_RINV];
                    const bool ok = vi_n4 && vj_s && vk[e];
                    oy.v[e] = ok ? advance<A>(uyc.v[e], uoy.v[e], rinv, acc) : (T)0;
                }
                {   // uz
                    const T sA = A::scl(A::sub(t5[e], t5m[e]), k0 ? g.sdx0 : ssx_m);
                    const T sB = A::scl(A::sub(t4[e], t4S.v[e]), k0 ? g.sdy0 : ssy);
                    T acc;
                    if (k0) acc = A::sub(A::add(A::add(sA, sB), t3u), A::scl(t3c[e], g.fdz0));
                    else acc = A::add(A::add(sA, sB), A::scl(A::sub(t3u, t3c[e]), sfz[e]));
                    const T rinv = stab[code_field<CodeT, M::B>(cd, F_RZ) * TAB_W + TAB_RINV];
                    const bool ok = vi_n4 && vj_n && vk[e];
                    oz.v[e] = ok ? advance<A>(uzc.v[e], uoz.v[e], rinv, acc) : (T)0;
                }
            }
            // i = 0: uy, uz keep u_new == u (App. B #9)
            if (n == 0) {
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const bool k0 = (e == 0) && first_cell;
                    oy.v[e] = (vj_s && (kb + e) < g.nz) ? uyc.v[e] : (T)0;
                    oz.v[e] = vk[e] ? ((k0 && p.line_save) ? p.line_save[j] : uzc.v[e]) : (T)0;
                }
            }
            if (in_box) {
                if (n <= g.nx - 2) *reinterpret_cast<PV *>(p.nw.ux + pl_n) = ox;
                if (j <= g.ny - 2) *reinterpret_cast<PV *>(p.nw.uy + pl_n) = oy;
                *reinterpret_cast<PV *>(p.nw.uz + pl_n) = oz;
            }
        } else {
            // shuffles must be executed by full warps: rows are warp-uniform, so nothing to do here
        }

        // (9) rotate the carried registers
        uxc = uxn; uyc = uyn; uzc = uzn;
        codec = coden;
        t3R = t3Rn;
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
template <class T, int R>
inline bool make_field_map(CUtensorMap *tm, void *base, int nzp, int ny, int planes) {
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

struct MarchPlan {
    int R = 16, nst = 3;
    int chunks = 1;
};

template <class T> inline bool march_supported(int nx, int ny, int nz, int nzp) {
    (void)nx; (void)ny; (void)nz; (void)nzp;
    return get_encode_tiled() != nullptr;
}
template <class T> inline const char *march_name() { return "march_tma"; }

template <class A, class M, int R, int NST>
inline int launch_march_cfg(const StepArgs<typename A::T> &p, const M &m, const CUtensorMap *maps, int nmat, int chunks,
                            cudaStream_t st) {
    using T = typename A::T;
    using C_ = MarchCfg<T, R, NST>;
    auto kern = k_step_march<A, M, R, NST>;
    static bool attr_done = false;   // per template instantiation
    if (!attr_done) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_::SMEM_BYTES);
        attr_done = true;
    }
    const int np = p.i_end - p.i_begin;
    if (np <= 0) return 0;
    if (chunks < 1) chunks = 1;
    if (chunks > np) chunks = np;
    const int chunk = (np + chunks - 1) / chunks;
    dim3 grid((p.g.nzp + C_::TZ - 1) / C_::TZ, (p.g.ny + C_::TY - 1) / C_::TY, (np + chunk - 1) / chunk);
    kern<<<grid, R * 32, C_::SMEM_BYTES, st>>>(maps[0], maps[1], maps[2], p, m, chunk, nmat);
    return 1;
}

}  // namespace phb

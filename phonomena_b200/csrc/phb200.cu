// phb200.cu -- host side of libphb200.so: C ABI (include/phb200.h), device memory, step
// orchestration, NCCL halo exchange, pinned-ring surface recorder.  sm_100a only.
#include "../../include/phb200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#ifndef PHB_DEFAULT_RW
#define PHB_DEFAULT_RW 2   // rows per warp of the marching kernel (PHB_MARCH_RW overrides at run time)
#endif
#include "fd_common.cuh"
#include "k_boundary.cuh"
#include "k_march.cuh"
#include "k_probe.cuh"
#include "k_naive.cuh"
#include "rec_ring.h"

using namespace phb;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const char *fmt, ...) {
    char b[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(b, sizeof b, fmt, ap);
    va_end(ap);
    g_err = b;
    return 1;
}
#define CU(x)                                                                                       \
    do {                                                                                            \
        cudaError_t e_ = (x);                                                                       \
        if (e_ != cudaSuccess) return fail("%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
    } while (0)
#define OK(x)                 \
    do {                      \
        int r_ = (x);         \
        if (r_) return r_;    \
    } while (0)

// ------------------------------------------------------------------------------------------
// NCCL through dlopen: the library has no link-time NCCL dependency (single-GPU use needs
// none); in a torch process the already-loaded libnccl.so.2 is picked up by SONAME.
// ------------------------------------------------------------------------------------------
struct Nccl {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static Nccl g_nccl;
static std::mutex g_nccl_mu;
static int nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.h) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail("cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(f)                                                                  \
    *(void **)(&g_nccl.f) = dlsym(h, "nccl" #f);                                \
    if (!g_nccl.f) return fail("libnccl: missing symbol nccl" #f);
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(Send) SYM(Recv) SYM(GroupStart) SYM(GroupEnd)
    SYM(GetErrorString)
#undef SYM
    g_nccl.h = h;
    return 0;
}
#define NC(x)                                                                                          \
    do {                                                                                               \
        ncclResult_t r_ = (x);                                                                         \
        if (r_ != ncclSuccess) return fail("%s:%d %s -> %s", __FILE__, __LINE__, #x, g_nccl.GetErrorString(r_)); \
    } while (0)

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct IEngine {
    virtual ~IEngine() {}
    virtual int set_spacing(const double *, const double *, const double *, const double *, const double *,
                            const double *) = 0;
    virtual int set_table(int, const double *, const double *) = 0;
    virtual int build_codes() = 0;
    virtual int set_abc(const double *) = 0;
    virtual int xfer(bool to_dev, int which, double *ux, double *uy, double *uz) = 0;
    virtual int stress(int which, double *T[6]) = 0;
    virtual int step() = 0;
    virtual int make_maps() = 0;
    virtual const char *kernel_name() = 0;
    // Bloch pair: the follower's share of a step, launched by the primary on the primary's streams
    virtual int follower_physics() = 0;
    virtual int follower_faces() = 0;
    virtual void advance() = 0;
};

struct phb_ctx {
    phb_cfg cfg;
    int nzp = 0;
    long long ps = 0;          // plane stride (elements)
    size_t esz = 0;            // sizeof(T)
    int device = 0;
    cudaStream_t st = nullptr, cst = nullptr;
    cudaStream_t zst = nullptr;        // second launch stream of a split step (the z-tile that owns the z = -1 face)
    cudaStream_t kst = nullptr;        // third: the z-tile that owns k = 0
    cudaEvent_t ev_kjoin = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // slab mode, push-after-faces: the x-chunks that read ghost planes run on their own lane, gated by the neighbours'
    // flags, beside the chunks that do not (launch_step)
    cudaStream_t est = nullptr, ezst = nullptr;
    cudaEvent_t ev_efork = nullptr, ev_ejoin = nullptr, ev_lane = nullptr, ev_lane_done = nullptr;
    int overlap = 0;                   // PHB_OVERLAP=1: edge x-chunks on their own lane behind the flag wait, middle chunks ungated
                                       // (measured at 2 GPUs, 512 planes each: 1.706 vs 1.693 ms -- the two extra x-chunks cost more
                                       // than the hidden wait saves; kept as an opt-in, tested bit-identical)
    int zsplit = -1;                   // PHB_ZSPLIT: 0 one launch for all z-tiles, 1 three specialised launches, 2 face tile + rest, -1 auto
    int faces_fused = 1;               // PHB_FACES_FUSED=0: ordered x, y, z face launches even when the z face is fused into the stencil
    cudaEvent_t ev_edge = nullptr, ev_comm = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    void *buf[3][3] = {};      // [buffer][component], each (nxl+2) planes
    int cur = 0;               // buffer holding u; old = (cur+2)%3, new = (cur+1)%3
    void *sp[6] = {};          // fdx fdy fdz sdx sdy sdz (padded, +1 offset applied at use)
    double sp0[6] = {};        // first elements (spacing)
    bool have_spacing = false;
    std::vector<double> mat_c12, mat_rho;   // host copy of the material table
    int nmat = 0;
    uint8_t *ids = nullptr;    // raw ids, planes [ids_ib, ids_ie)
    int ids_ib = 0, ids_ie = 0;
    uint8_t *code = nullptr;   // stencil class per cell, (nxl+2) planes, padded layout
    void *tab = nullptr;       // class table, ncls x CLS_W
    int ncls = 0;
    void *line_save = nullptr;
    void *dbg_push = nullptr;  // PHB_DEBUG_SELF_PUSH scratch planes
    double *w = nullptr;       // source samples for steps w_base .. w_base + nw - 1 (capacity w_cap, pointer kept while it fits)
    long long nw = 0, w_base = 0, w_cap = 0;
    long long *src_idx = nullptr;   // device: index into w of the next step's sample (k_source advances it)
    // single-GPU steps replayed as CUDA graphs (one per buffer rotation phase): a step is 5 short launches, and on
    // small grids the host's launch rate, not the GPU, sets the pace
    cudaGraphExec_t gexec[3] = {};
    int gnodes[3] = {0, 0, 0};
    int graph_mode = 1;             // PHB_GRAPH=0 disables
    long long plain_steps = 0;      // steps launched kernel by kernel since the last graph invalidation
    double abc[8] = {};
    bool have_abc = false;
    long long tt = 0;
    std::atomic<long long> launches{0};
    long long dev_bytes = 0;
    // comm
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    // fused halo push over NVLink peer memory (CUDA IPC); halo = 0 none, 1 NCCL, 2 peer stores
    int push_fused = 0;          // halo == 2: 0 = push kernel after the faces (default), 1 = edge planes stored by the stencil kernel itself
    int halo = 0;
    void *peer_buf[2][3] = {};   // [left | right][buffer]: the neighbour's displacement buffers, mapped
    int peer_nxl[2] = {0, 0};
    int *flags = nullptr;        // [0] written by the left neighbour, [1] by the right one: last step pushed; [8] = steps done (k_wait_flags)
    int *peer_flags[2] = {};     // the neighbours' flag arrays, mapped
    // recorder: pinned host ring (rec_ring.h) filled through a device staging ring; drained by the native writer
    // threads (phb_writer_start) or by the caller (phb_record_next / phb_record_release)
    RecRing rec;
    NativeWriter wr;
    double *ring_dev = nullptr;   // device staging ring (same slots): the gather kernel runs at HBM speed,
                                  // the D2H copy to the pinned ring runs on its own stream behind it
    cudaStream_t rst = nullptr;
    std::vector<cudaEvent_t> stage_ev;
    std::atomic<int> cancel{0};   // phb_cancel: the running / next phb_run returns after the current step
    // marching kernel: TMA descriptors per [buffer][component], tile plan
    MarchMaps mm[3];           // indexed by the buffer that holds u_cur
    bool maps_ok = false;
    int zfuse = -1;          // z = -1 face inside the marching kernel: -1 auto, 0 never, 1 whenever possible (PHB_ZFUSE)
    int mR = 16, mNST = 4, mRW = PHB_DEFAULT_RW, mChunks = 0;   // marching kernel: tile rows, u_cur ring depth, rows per warp
    // per-kernel timing of the stencil launches (bench roofline): event pairs on the launch stream
    bool prof = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev;
    size_t prof_used = 0;
    double prof_ms = 0;
    long long prof_n = 0;
    // line probes (k_probe.cuh)
    struct Probe { int comp, j, k, rows; long long cap, frames; double *trace; };
    std::vector<Probe> probes;
    // Bloch-periodic pair (phb_bloch_pair): the real part (role 1) steps both contexts; the imaginary part (role 2) follows
    phb_ctx *partner = nullptr;
    int bloch_role = 0;
    double bloch_c = 1.0, bloch_s = 0.0;
    cudaEvent_t ev_pair = nullptr;
    IEngine *eng = nullptr;
    std::mutex mu;
};

static void graph_invalidate(phb_ctx *c) {
    for (int q = 0; q < 3; ++q) {
        if (c->gexec[q]) cudaGraphExecDestroy(c->gexec[q]);
        c->gexec[q] = nullptr;
    }
    c->plain_steps = 0;
}

static int prof_collect(phb_ctx *c) {
    for (size_t q = 0; q < c->prof_used; ++q) {
        float ms = 0;
        CU(cudaEventSynchronize(c->prof_ev[q].second));
        CU(cudaEventElapsedTime(&ms, c->prof_ev[q].first, c->prof_ev[q].second));
        c->prof_ms += ms;
        c->prof_n++;
    }
    c->prof_used = 0;
    return 0;
}

static int dmalloc(phb_ctx *c, void **p, size_t bytes, bool zero = true) {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
    if (e != cudaSuccess)
        return fail("cudaMalloc(%zu bytes) failed: %s (allocated so far: %lld bytes)", bytes, cudaGetErrorString(e),
                    c->dev_bytes);
    c->dev_bytes += (long long)bytes;
    if (zero) CU(cudaMemsetAsync(*p, 0, bytes ? bytes : 1, c->st));
    return 0;
}

static inline dim3 grid3(int nx, int ny, int nz, dim3 b) {
    return dim3((nx + b.x - 1) / b.x, (ny + b.y - 1) / b.y, (nz + b.z - 1) / b.z);
}
static inline dim3 block_for(int nfast) {
    int bx = 32;
    while (bx < nfast && bx < 128) bx <<= 1;
    return dim3(bx, 256 / bx, 1);
}

// ------------------------------------------------------------------------------------------
// typed engine
// ------------------------------------------------------------------------------------------
template <class T>
struct Engine : IEngine {
    phb_ctx *c;
    explicit Engine(phb_ctx *c_) : c(c_) {}

    Geo<T> geo() const {
        Geo<T> g;
        g.nx = c->cfg.nx; g.ny = c->cfg.ny; g.nz = c->cfg.nz;
        g.x0 = c->cfg.x0; g.nxl = c->cfg.nxl;
        g.nzp = c->nzp; g.ps = c->ps;
        g.fdx = (const T *)c->sp[0] + 1; g.fdy = (const T *)c->sp[1] + 1; g.fdz = (const T *)c->sp[2] + 1;
        g.sdx = (const T *)c->sp[3] + 1; g.sdy = (const T *)c->sp[4] + 1; g.sdz = (const T *)c->sp[5] + 1;
        const bool ex = c->cfg.arith == PHB_EXACT;
        auto f = [&](double s) { return (T)(ex ? s : 1.0 / s); };
        g.fdx0 = f(c->sp0[0]); g.fdy0 = f(c->sp0[1]); g.fdz0 = f(c->sp0[2]);
        g.sdx0 = f(c->sp0[3]); g.sdy0 = f(c->sp0[4]); g.sdz0 = f(c->sp0[5]);
        return g;
    }
    Fld<T> fld(int b) const { return Fld<T>{(T *)c->buf[b][0], (T *)c->buf[b][1], (T *)c->buf[b][2]}; }
    // plain state: three buffers rotate (old <- cur <- new).  Compensated state (PHB_COMP): buffers 0 / 1
    // ping-pong u, buffer 2 holds delta = u - u_old and is updated in place.
    bool comp() const { return c->cfg.arith == PHB_COMP; }
    int b_cur() const { return c->cur; }
    int b_old() const { return comp() ? 2 : (c->cur + 2) % 3; }
    int b_new() const { return comp() ? 1 - c->cur : (c->cur + 1) % 3; }

    int set_spacing(const double *fdx, const double *fdy, const double *fdz, const double *sdx, const double *sdy,
                    const double *sdz) override {
        const double *src[6] = {fdx, fdy, fdz, sdx, sdy, sdz};
        const int n[3] = {c->cfg.nx, c->cfg.ny, c->cfg.nz};
        const bool ex = c->cfg.arith == PHB_EXACT;
        for (int a = 0; a < 6; ++a) {
            const int ax = a % 3, len = (a < 3) ? n[ax] - 1 : n[ax] - 2;
            std::vector<T> h(n[ax] + 2, (T)1);
            for (int m = 0; m < len; ++m) {
                if (!(src[a][m] > 0)) return fail("spacing array %d has a non-positive entry at %d", a, m);
                h[m + 1] = (T)(ex ? src[a][m] : 1.0 / src[a][m]);
            }
            c->sp0[a] = src[a][0];
            if (!c->sp[a]) OK(dmalloc(c, &c->sp[a], h.size() * sizeof(T)));
            CU(cudaMemcpyAsync(c->sp[a], h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, c->st));
            CU(cudaStreamSynchronize(c->st));
        }
        c->have_spacing = true;
        return 0;
    }

    int set_table(int nmat, const double *c12, const double *rho) override {
        c->mat_c12.assign(c12, c12 + (size_t)nmat * 12);
        c->mat_rho.assign(rho, rho + nmat);
        c->nmat = nmat;
        return 0;
    }

    // stencil classes: distinct (shifted, range-masked) material tuples -> 1 byte per cell + table
    int build_codes() override {
        if (!c->ids || !c->nmat) return fail("material ids / table not set");
        ClsGeo q{c->ids, c->ids_ib, c->ids_ie, c->cfg.nx, c->cfg.ny, c->cfg.nz, c->nzp, c->cfg.x0, c->cfg.nxl};
        uint32_t *slots = nullptr, *dkeys = nullptr;
        int *ovf = nullptr;
        CU(cudaMalloc(&slots, CLS_SLOTS * sizeof(uint32_t)));
        CU(cudaMalloc(&ovf, sizeof(int)));
        CU(cudaMemsetAsync(slots, 0xFF, CLS_SLOTS * sizeof(uint32_t), c->st));
        CU(cudaMemsetAsync(ovf, 0, sizeof(int), c->st));
        dim3 b = block_for(c->nzp), g = grid3(c->nzp, c->cfg.ny, c->cfg.nxl + 2, b);
        k_cls_collect<<<g, b, 0, c->st>>>(q, slots, ovf);
        std::vector<uint32_t> hs(CLS_SLOTS);
        int hovf = 0;
        CU(cudaMemcpyAsync(hs.data(), slots, CLS_SLOTS * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->st));
        CU(cudaMemcpyAsync(&hovf, ovf, sizeof(int), cudaMemcpyDeviceToHost, c->st));
        cudaError_t e = cudaStreamSynchronize(c->st);
        cudaFree(ovf);
        if (e != cudaSuccess) { cudaFree(slots); CU(e); }
        std::vector<uint32_t> keys;
        for (uint32_t v : hs) if (v != CLS_EMPTY) keys.push_back(v);
        {   // class 0 is always the all-VOID class (what zero-filled out-of-range class bytes select)
            int fv[7] = {MAT_VOID, MAT_VOID, MAT_VOID, MAT_VOID, MAT_VOID, MAT_VOID, MAT_VOID};
            const uint32_t vk = cls_key(fv, false);
            keys.erase(std::remove(keys.begin(), keys.end(), vk), keys.end());
            std::sort(keys.begin(), keys.end());
            keys.insert(keys.begin(), vk);
        }
        cudaFree(slots);
        if (hovf || keys.size() > MAX_CLS)
            return fail("more than %d distinct material stencil classes (%zu found): medium too heterogeneous for the "
                        "1-byte class code", (int)MAX_CLS, keys.size());
        // table rows, evaluated in float64 exactly as the reference's operands
        std::vector<T> tab(keys.size() * CLS_W, (T)0);
        for (size_t n = 0; n < keys.size(); ++n) {
            int f[7];
            for (int m = 0; m < 7; ++m) f[m] = (keys[n] >> (4 * m)) & 15;
            for (int m = 0; m < 7; ++m)
                if (f[m] != MAT_VOID && f[m] >= c->nmat) return fail("material id %d used in the grid but the table has %d entries", f[m], c->nmat);
            const bool k0 = (keys[n] >> 28) & 1;
            T *row = &tab[n * CLS_W];
            if (f[0] != MAT_VOID)
                for (int m = 0; m < 9; ++m) row[m] = (T)c->mat_c12[(size_t)f[0] * 12 + m];
            if (k0) row[6] = row[7] = row[8] = (T)0;                       // T3 = 0 on the free surface (:418)
            if (f[1] != MAT_VOID) row[CLS_C44] = (T)c->mat_c12[(size_t)f[1] * 12 + 9];
            if (f[2] != MAT_VOID) row[CLS_C55] = (T)c->mat_c12[(size_t)f[2] * 12 + 10];
            if (f[3] != MAT_VOID) row[CLS_C66] = (T)c->mat_c12[(size_t)f[3] * 12 + 11];
            // (dt**2 / P) is evaluated first in the reference (base_solver.py:443), in float64
            if (f[4] != MAT_VOID) row[CLS_RX] = (T)(c->cfg.d2 / c->mat_rho[f[4]]);
            if (f[5] != MAT_VOID) row[CLS_RY] = (T)(c->cfg.d2 / c->mat_rho[f[5]]);
            if (f[6] != MAT_VOID) row[CLS_RZ] = (T)(c->cfg.d2 / c->mat_rho[f[6]]);
        }
        if (c->tab) { cudaFree(c->tab); c->tab = nullptr; }
        OK(dmalloc(c, &c->tab, tab.size() * sizeof(T), false));
        CU(cudaMemcpyAsync(c->tab, tab.data(), tab.size() * sizeof(T), cudaMemcpyHostToDevice, c->st));
        c->ncls = (int)keys.size();
        if (!c->code) OK(dmalloc(c, (void **)&c->code, (size_t)(c->cfg.nxl + 2) * c->ps));
        CU(cudaMalloc(&dkeys, keys.size() * sizeof(uint32_t)));
        CU(cudaMemcpyAsync(dkeys, keys.data(), keys.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->st));
        k_cls_assign<<<g, b, 0, c->st>>>(q, dkeys, c->ncls, c->code);
        c->launches += 2;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
        cudaFree(dkeys);
        CU(e);
        return make_maps();
    }

    int set_abc(const double *k) override {
        memcpy(c->abc, k, sizeof c->abc);
        c->have_abc = true;
        return 0;
    }

    // planes of component comp owned by this slab, in host arrays
    int owned_planes(int comp) const {
        const int hi = (comp == 0) ? c->cfg.nx - 1 : c->cfg.nx;
        int e = c->cfg.x0 + c->cfg.nxl;
        if (e > hi) e = hi;
        return e - c->cfg.x0 > 0 ? e - c->cfg.x0 : 0;
    }

    int xfer(bool to_dev, int which, double *ux, double *uy, double *uz) override {
        const int b = which == PHB_CUR ? b_cur() : b_old();
        double *h[3] = {ux, uy, uz};
        for (int comp = 0; comp < 3; ++comp) {
            if (!h[comp]) continue;
            const int np = owned_planes(comp);
            const int ey = c->cfg.ny - (comp == 1), ez = c->cfg.nz - (comp == 2);
            const size_t bytes = (size_t)np * ey * ez * sizeof(double);
            if (!bytes) continue;
            double *tmp = nullptr;
            CU(cudaMalloc(&tmp, bytes));
            dim3 bl = block_for(ez), g = grid3(ez, ey, np, bl);
            const bool as_delta = this->comp() && which == PHB_OLD;   // the slot holds u - u_old (set PHB_CUR first)
            const T *ucur = (const T *)c->buf[b_cur()][comp];
            if (to_dev) {
                CU(cudaMemcpyAsync(tmp, h[comp], bytes, cudaMemcpyHostToDevice, c->st));
                if (as_delta) k_scatter_delta<T><<<g, bl, 0, c->st>>>(tmp, ucur, (T *)c->buf[b][comp], np, ey, ez, 1, c->cfg.ny, c->nzp);
                else k_scatter<T><<<g, bl, 0, c->st>>>(tmp, (T *)c->buf[b][comp], np, ey, ez, 1, c->cfg.ny, c->nzp);
            } else {
                if (as_delta) k_gather_delta<T><<<g, bl, 0, c->st>>>(ucur, (const T *)c->buf[b][comp], tmp, np, ey, ez, 1, c->cfg.ny, c->nzp);
                else k_gather<T><<<g, bl, 0, c->st>>>((const T *)c->buf[b][comp], tmp, np, ey, ez, 1, c->cfg.ny, c->nzp);
                CU(cudaMemcpyAsync(h[comp], tmp, bytes, cudaMemcpyDeviceToHost, c->st));
            }
            c->launches++;
            cudaError_t e = cudaStreamSynchronize(c->st);
            cudaFree(tmp);
            CU(e);
            CU(cudaGetLastError());
        }
        return 0;
    }

    MatCls<T> mat() const { return MatCls<T>{c->code, (const T *)c->tab, c->ncls}; }
    template <class F>
    int dispatch(F &&f) {
        if (!c->code || !c->tab) return fail("material not set (table + ids)");
        if (c->cfg.arith == PHB_EXACT) return f.template operator()<Ar<T, true>>();
        if (c->cfg.arith == PHB_COMP) return f.template operator()<Ar<T, false, true>>();
        return f.template operator()<Ar<T, false>>();
    }

    int stress(int which, double *Th[6]) override {
        if (!c->have_spacing || !c->code) return fail("spacing/material not set");
        if (c->bloch_role) return fail("the stress dump is not available for a Bloch pair (the wrapped stresses mix both parts)");
        const int b = which == PHB_CUR ? b_cur() : b_old();
        const int nx = c->cfg.nx, ny = c->cfg.ny, nz = c->cfg.nz;
        const int n = c->cfg.nxl, n5 = owned_planes(0);
        const size_t cnt[6] = {(size_t)n * ny * nz, (size_t)n * ny * nz, (size_t)n * ny * nz,
                               (size_t)n * (ny - 1) * (nz - 1), (size_t)n5 * ny * (nz - 1),
                               (size_t)n5 * (ny - 1) * nz};
        double *d[6] = {};
        for (int q = 0; q < 6; ++q) {
            CU(cudaMalloc(&d[q], (cnt[q] ? cnt[q] : 1) * sizeof(double)));
            CU(cudaMemsetAsync(d[q], 0, cnt[q] * sizeof(double), c->st));
        }
        Geo<T> g = geo();
        Fld<T> u = fld(b);
        dim3 bl = block_for(nz), gr = grid3(nz, ny, n, bl);
        const int ib = c->cfg.x0, ie = c->cfg.x0 + n;
        (void)nx;
        MatCls<T> m = mat();
        auto run = [&]<class A>() -> int {
            k_stress_dump<A, MatCls<T>><<<gr, bl, 0, c->st>>>(g, u, m, ib, ie, d[0], d[1], d[2], d[3], d[4], d[5], periodic_y());
            return 0;
        };
        int r = dispatch(run);
        c->launches++;
        cudaError_t e = cudaGetLastError();
        for (int q = 0; q < 6 && !r && e == cudaSuccess; ++q)
            if (Th[q]) e = cudaMemcpyAsync(Th[q], d[q], cnt[q] * sizeof(double), cudaMemcpyDeviceToHost, c->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
        for (int q = 0; q < 6; ++q) cudaFree(d[q]);
        if (r) return r;
        CU(e);
        return 0;
    }

    // ---- one time step -----------------------------------------------------------------------
    // edge_b >= 0: edge launch of the two single planes ib and edge_b (marching kernel only)
    int physics(int ib, int ie, int edge_b = -1, int lane = 0) {
        if (ie <= ib) return 0;
        // launch lane: 0 = the step's main stream, 1 = the edge lane of a slab step
        cudaStream_t S = lane ? c->est : c->st, ZS = lane ? c->ezst : c->zst;
        cudaEvent_t EF = lane ? c->ev_efork : c->ev_fork, EJ = lane ? c->ev_ejoin : c->ev_join;
        StepArgs<T> p;
        p.edge_b = edge_b;
        for (int q = 0; q < 3; ++q) p.push_lo[q] = p.push_hi[q] = nullptr;
        if (c->halo == 2 && c->push_fused && use_march()) {
            const int bn = b_new();
            if (c->peer_buf[0][bn]) {       // left neighbour: its right ghost plane (local plane nxl_L + 1)
                const long long comp = (long long)(c->peer_nxl[0] + 2) * c->ps;
                for (int q = 0; q < 3; ++q) p.push_lo[q] = (T *)c->peer_buf[0][bn] + q * comp + (long long)(c->peer_nxl[0] + 1) * c->ps;
            }
            if (c->peer_buf[1][bn]) {       // right neighbour: its left ghost plane (local plane 0)
                const long long comp = (long long)(c->peer_nxl[1] + 2) * c->ps;
                for (int q = 0; q < 3; ++q) p.push_hi[q] = (T *)c->peer_buf[1][bn] + q * comp;
            }
        }
        {   // timing aid (single GPU): run the PUSH instantiation with device-local scratch planes as "neighbours" --
            // separates what the push code costs the kernel from what NVLink and the system fence cost
            static const int self_push = getenv("PHB_DEBUG_SELF_PUSH") ? atoi(getenv("PHB_DEBUG_SELF_PUSH")) : 0;
            if (self_push && c->nranks == 1 && use_march()) {
                if (!c->dbg_push) OK(dmalloc(c, &c->dbg_push, (size_t)6 * c->ps * c->esz));
                for (int q = 0; q < 3; ++q) {
                    if (self_push & 1) p.push_lo[q] = (T *)c->dbg_push + (long long)q * c->ps;
                    if (self_push & 2) p.push_hi[q] = (T *)c->dbg_push + (long long)(3 + q) * c->ps;
                }
            }
        }
        p.g = geo();
        p.cur = fld(b_cur()); p.old = fld(b_old()); p.nw = fld(b_new());
        p.line_save = (c->w && c->cfg.x0 == 0) ? (const T *)c->line_save : nullptr;
        p.i_begin = ib; p.i_end = ie;
        p.ztile0 = 0;
        const bool march = use_march();
        p.zface = (march && zface_fused()) ? 1 : 0;
        static const int zf_nop = getenv("PHB_DEBUG_ZF_NOP") ? atoi(getenv("PHB_DEBUG_ZF_NOP")) : 0;   // timing aid: ZF instantiation, face never applied
        if (zf_nop && march && !comp() && c->mRW == 2) p.zface = 2;
        p.zf_cl = (T)c->abc[6]; p.zf_ct = (T)c->abc[7];
        std::pair<cudaEvent_t, cudaEvent_t> *pe = nullptr;
        if (c->prof) {
            if (c->prof_used == c->prof_ev.size()) {
                if (c->prof_ev.size() >= 4096) OK(prof_collect(c));
                else {
                    cudaEvent_t a, b;
                    CU(cudaEventCreate(&a));
                    CU(cudaEventCreate(&b));
                    c->prof_ev.push_back({a, b});
                }
            }
            pe = &c->prof_ev[c->prof_used++];
            CU(cudaEventRecord(pe->first, S));
        }
        if (c->cfg.kernel == PHB_KERNEL_MARCH && !march)
            return fail("kernel=march requested but the marching kernel does not support this grid");
        MatCls<T> m = mat();
        auto run = [&]<class A>() -> int {
            if (march) {
                const MarchMaps &mp = c->mm[b_cur()];
                const int ch = edge_b >= 0 ? 1 : plan_chunks(ie - ib);      // an edge launch is two chunks of ie - ib planes
                int r = -2;
                // Split step: the instantiation with the z = -1 face code is slower for EVERY block (+22 us per step even
                // when the face code never runs: the extra path costs the plane loop scheduling freedom), while the work
                // itself concerns only the blocks of the last z-tile.  So the tile that owns the face is launched on its
                // own with the face instantiation, on a second (higher-priority) stream beside the launch for all other
                // z-tiles without it; both finish inside the same waves of blocks.
                const int V = VecOf<T>::V, nzt = (c->nzp + 32 * V - 1) / (32 * V);
                // auto: a wide cross-section (>= 200 k cells, i.e. every part fills its waves) and, for fp32, a long step
                const long long yz_cells = (long long)c->cfg.ny * c->cfg.nz;
                const bool want_split = c->zsplit > 0 || (c->zsplit < 0 && yz_cells >= 200000 &&
                                                          (sizeof(T) == 8 || (long long)(ie - ib) * yz_cells >= 100000000LL));
                if (p.zface == 1 && want_split && nzt >= 2 && (edge_b < 0 || ie - ib >= 16) && c->mR == 16 && c->mNST == 4 && c->mRW == 2) {
                    // parts: [face tile: ZF instantiation] [tile 0: with the k = 0 selects] [tiles between: neither] -- the
                    // middle part is the bulk and stays on the lane's main stream; PHB_ZSPLIT=2: face tile + the rest (round 2a)
                    const bool three = c->zsplit != 2 && nzt >= 3 && lane == 0;
                    CU(cudaEventRecord(EF, S));
                    CU(cudaStreamWaitEvent(ZS, EF, 0));
                    StepArgs<T> pz = p;
                    pz.ztile0 = nzt - 1;
                    const int rz = launch_march_cfg<A, 16, 4, 2>(pz, m, mp, ch, ZS, 1, 2);
                    CU(cudaEventRecord(EJ, ZS));
                    p.zface = 0;
                    int r0 = 0;
                    if (three) {
                        CU(cudaStreamWaitEvent(c->kst, EF, 0));
                        r0 = launch_march_cfg<A, 16, 4, 2>(p, m, mp, ch, c->kst, 1, 1);       // z-tile 0
                        CU(cudaEventRecord(c->ev_kjoin, c->kst));
                        p.ztile0 = 1;
                        r = launch_march_cfg<A, 16, 4, 2>(p, m, mp, ch, S, nzt - 2, 3);       // z-tiles 1 .. nzt-2
                        CU(cudaStreamWaitEvent(S, c->ev_kjoin, 0));
                    } else {
                        r = launch_march_cfg<A, 16, 4, 2>(p, m, mp, ch, S, nzt - 1, nzt == 2 ? 1 : 0);
                    }
                    CU(cudaStreamWaitEvent(S, EJ, 0));
                    if (rz < 0 || r < 0 || r0 < 0) return fail("marching kernel needs more shared memory than the device allows (%d classes)", c->ncls);
                    c->launches += rz + r + r0;
                    return 0;
                }
                if (c->mR == 16 && c->mNST == 4 && c->mRW == 2) r = launch_march_cfg<A, 16, 4, 2>(p, m, mp, ch, S);
                else if (c->mR == 16 && c->mNST == 4) r = launch_march_cfg<A, 16, 4>(p, m, mp, ch, S);
                else if (c->mR == 16 && c->mNST == 3) r = launch_march_cfg<A, 16, 3>(p, m, mp, ch, S);
                else if (c->mR == 8 && c->mNST == 4) r = launch_march_cfg<A, 8, 4>(p, m, mp, ch, S);
                if (r == -2) return fail("no marching-kernel instantiation for R=%d NST=%d", c->mR, c->mNST);
                if (r == -3) return fail("internal: fused z face requested for a marching-kernel configuration without it");
                if (r < 0 && (c->cfg.kernel == PHB_KERNEL_MARCH || (c->halo == 2 && c->push_fused)))
                    return fail("marching kernel needs more shared memory than the device allows (%d classes)", c->ncls);
                if (r >= 0) { c->launches += r; return 0; }
                c->maps_ok = false;      // kernel = auto: too many stencil classes for the shared-memory table -> naive kernel
            }
            {
                dim3 bl = block_for(c->cfg.nz), gr = grid3(c->cfg.nz, c->cfg.ny, ie - ib, bl);
                if (edge_b >= 0) return fail("internal: edge launch without the marching kernel");
                k_step_naive<A, MatCls<T>><<<gr, bl, 0, S>>>(p, m);
                c->launches++;
            }
            return 0;
        };
        OK(dispatch(run));
        if (pe) CU(cudaEventRecord(pe->second, S));
        CU(cudaGetLastError());
        return 0;
    }
    // the marching kernel applies the z = -1 face itself when the face points and their inner neighbours share a
    // lane: nz a multiple of the vector width, and not in the first lane of a tile (V = 2 reads the lane before)
    bool zface_fused() const {
        const int V = VecOf<T>::V, nz = c->cfg.nz;
        // measured at 512^3: fp32 gains 4 % (the separate strided face kernel costs 48 us per step), fp64 loses about as much
        // in the stencil kernel itself as the face kernel costs -> default on for fp32 only (PHB_ZFUSE=0/1 overrides)
        // round 2: with the split step (physics()) the fused face pays in fp64 too -> on by default for both types
        const bool want = c->zfuse < 0 ? true : (c->zfuse != 0);
        if (!want || comp() || c->mRW != 2 || nz < 2 * V || nz % V != 0) return false;
        return ((nz - 1) % (32 * V)) / V >= 1;
    }
    bool use_march() const {
        if (c->cfg.kernel == PHB_KERNEL_NAIVE) return false;
        // kernel = auto on a small slab: the x-march is a serial chain per (y, z) tile, and a grid of a few thousand
        // cells has too few tiles to fill the SMs -- one thread per cell is faster there (measured: 32^3 25 vs 33 us per
        // step, equal at 64^3, 2x slower at 96^3; EXACT arithmetic 8x faster on the 31x21x6 default.json grid)
        // (not for slabs that push their halos from inside the marching kernel)
        if (c->cfg.kernel == PHB_KERNEL_AUTO && !(c->halo == 2 && c->push_fused) && (long long)c->cfg.nxl * c->cfg.ny * c->cfg.nz < 200000) return false;
        return c->maps_ok && march_fits();
    }
    // x-chunks per launch: fill whole waves of (148 SMs x resident blocks) with the (y,z) tiles
    // the marching kernel addresses elements with 31-bit offsets
    bool march_fits() const { return (long long)(c->cfg.nxl + 2) * c->ps < (1LL << 31); }
    int plan_chunks(int np) const {
        if (c->mChunks > 0) return c->mChunks;
        const int V = VecOf<T>::V, TZ = 32 * V, TY = c->mR - 2;
        const long long tiles = (long long)((c->nzp + TZ - 1) / TZ) * ((c->cfg.ny + TY - 1) / TY);
        const long long slots = 148LL * (c->mR <= 8 ? 2 : 1);
        int best = 1;
        double best_eff = 0;
        // with a halo exchange in flight NCCL CTAs hold a few SMs when the interior launch starts: finer chunks let
        // the block scheduler rebalance instead of ending with a straggler wave (measured: 4 chunks)
        const int ch_min = (c->nranks > 1 && c->halo != 2 && np >= 64) ? 4 : 1;    // (the fused push has no NCCL kernel beside it)
        for (int ch = ch_min; ch <= 16 && ch * 8 <= np; ++ch) {
            if (ch < 16 && (np + ch - 1) / ch > 256) continue;    // x-spacing table lives in shared memory
            const long long blocks = tiles * ch, waves = (blocks + slots - 1) / slots;
            const double eff = (double)blocks / (double)(waves * slots) * (1.0 - 1.5 * ch / (double)np);
            if (eff > best_eff + 1e-9) { best_eff = eff; best = ch; }
        }
        return best;
    }
    int make_maps() override {
        c->maps_ok = false;
        if (c->cfg.kernel == PHB_KERNEL_NAIVE || !c->code) return 0;
        const int planes = c->cfg.nxl + 2;
        CUtensorMap cur[3], old[3], cls;
        bool ok = make_class_map<T>(&cls, c->code, c->nzp, c->cfg.ny, planes, c->mR);
        for (int b = 0; b < 3 && ok; ++b)      // the three components of a buffer are one allocation
            ok = make_field_maps<T>(&cur[b], &old[b], c->buf[b][0], c->nzp, c->cfg.ny, planes, c->mR);
        if (!ok) {
            if (c->cfg.kernel == PHB_KERNEL_MARCH) return fail("cuTensorMapEncodeTiled failed");
            return 0;
        }
        for (int b = 0; b < 3; ++b) {
            const int bo = comp() ? 2 : (b + 2) % 3;
            c->mm[b].u = cur[b];
            c->mm[b].o = old[bo];
            c->mm[b].c = cls;
        }
        c->maps_ok = true;
        return 0;
    }

    const char *kernel_name() override { return use_march() ? march_name<T>() : "naive"; }

    AbcArgs<T> abc_args(int ib, int ie) {
        AbcArgs<T> a;
        a.g = geo(); a.cur = fld(b_cur()); a.nw = fld(b_new()); a.dl = fld(b_old());
        a.clx = (T)c->abc[0]; a.ctx = (T)c->abc[1]; a.cly0 = (T)c->abc[2]; a.cty0 = (T)c->abc[3];
        a.cly1 = (T)c->abc[4]; a.cty1 = (T)c->abc[5]; a.clz = (T)c->abc[6]; a.ctz = (T)c->abc[7];
        a.i_begin = ib; a.i_end = ie;
        a.z_edges_only = 0;
        return a;
    }
    int abc_x() {
        AbcArgs<T> a = abc_args(0, 0);
        dim3 bl = block_for(c->cfg.nz), gr = grid3(c->cfg.nz, c->cfg.ny, 1, bl);
        if (c->cfg.arith == PHB_EXACT) k_abc_x<Ar<T, true>><<<gr, bl, 0, c->st>>>(a);
        else if (comp()) k_abc_x<Ar<T, false, true>><<<gr, bl, 0, c->st>>>(a);
        else k_abc_x<Ar<T, false>><<<gr, bl, 0, c->st>>>(a);
        c->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    // x face (if this rank owns it) + y faces + z face of planes [ib, ie): one launch when the stencil kernel has applied the
    // z face itself (k_faces_fused), else the ordered launches x, y, z
    int faces(int ib, int ie, bool with_x) {
        if (ie <= ib) return 0;
        if (c->faces_fused && use_march() && zface_fused() && !periodic_y() && !comp() && !getenv("PHB_DEBUG_ZF_NOP") && c->cfg.ny >= 5) {
            AbcArgs<T> a = abc_args(ib, ie);
            const long long n = (with_x ? (long long)c->cfg.ny * c->cfg.nz : 0) + 2LL * (ie - ib) * c->cfg.nz + 3LL * (ie - ib) +
                                (with_x ? 2LL * (c->cfg.ny - 3) : 0);
            const unsigned gr = (unsigned)((n + 255) / 256);
            if (c->cfg.arith == PHB_EXACT) k_faces_fused<Ar<T, true>><<<gr, 256, 0, c->st>>>(a, with_x ? 1 : 0);
            else k_faces_fused<Ar<T, false>><<<gr, 256, 0, c->st>>>(a, with_x ? 1 : 0);
            c->launches++;
            CU(cudaGetLastError());
            return 0;
        }
        if (with_x) OK(abc_x());
        return abc_yz(ib, ie);
    }
    bool periodic_y() const { return c->cfg.bc_y == PHB_BC_PERIODIC; }
    // periodic y boundaries: recompute the rows the wrapped stresses reach + the displacement copies (k_pbc_y); must run
    // right after the stencil launch of the same planes, BEFORE the x face (the reference order: pbc, then apply_u_abc)
    int pbc_y(int ib, int ie) {
        if (ie <= ib || !periodic_y() || c->bloch_role) return 0;      // (a Bloch pair runs k_pbc_y_bloch once for both parts)
        StepArgs<T> p;
        for (int q = 0; q < 3; ++q) p.push_lo[q] = p.push_hi[q] = nullptr;
        p.edge_b = -1; p.zface = 0; p.ztile0 = 0;
        p.zf_cl = p.zf_ct = (T)0;
        p.g = geo();
        p.cur = fld(b_cur()); p.old = fld(b_old()); p.nw = fld(b_new());
        p.line_save = (c->w && c->cfg.x0 == 0) ? (const T *)c->line_save : nullptr;
        p.i_begin = ib; p.i_end = ie;
        MatCls<T> m = mat();
        dim3 bl = block_for(c->cfg.nz), gr = grid3(c->cfg.nz, ie - ib, 1, bl);
        if (c->cfg.arith == PHB_EXACT) k_pbc_y<Ar<T, true>, MatCls<T>><<<gr, bl, 0, c->st>>>(p, m);
        else k_pbc_y<Ar<T, false>, MatCls<T>><<<gr, bl, 0, c->st>>>(p, m);
        c->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    int abc_yz(int ib, int ie) {
        if (ie <= ib) return 0;
        AbcArgs<T> a = abc_args(ib, ie);
        a.z_edges_only = (use_march() && zface_fused() && !getenv("PHB_DEBUG_ZF_NOP")) ? 1 : 0;
        if (!periodic_y()) {
            dim3 bl = block_for(c->cfg.nz), gr = grid3(c->cfg.nz, ie - ib, 1, bl);
            gr.z = 2;
            if (c->cfg.arith == PHB_EXACT) k_abc_y<Ar<T, true>><<<gr, bl, 0, c->st>>>(a);
            else if (comp()) k_abc_y<Ar<T, false, true>><<<gr, bl, 0, c->st>>>(a);
            else k_abc_y<Ar<T, false>><<<gr, bl, 0, c->st>>>(a);
        }
        {
            dim3 bl(32, 8, 1), gr = grid3(c->cfg.ny, ie - ib, 1, bl);
            if (c->cfg.arith == PHB_EXACT) k_abc_z<Ar<T, true>><<<gr, bl, 0, c->st>>>(a);
            else if (comp()) k_abc_z<Ar<T, false, true>><<<gr, bl, 0, c->st>>>(a);
            else k_abc_z<Ar<T, false>><<<gr, bl, 0, c->st>>>(a);
        }
        c->launches += periodic_y() ? 1 : 2;
        CU(cudaGetLastError());
        return 0;
    }

    int exchange() {
        // send new[x0] to the left neighbour (its right ghost), new[x0+nxl-1] to the right
        // neighbour (its left ghost); receive the mirror images.  3 components per direction.
        const ncclDataType_t dt = sizeof(T) == 8 ? ncclDouble : ncclFloat;
        const size_t n = (size_t)c->ps;
        const int bn = b_new();
        NC(g_nccl.GroupStart());
        for (int comp = 0; comp < 3; ++comp) {
            T *base = (T *)c->buf[bn][comp];
            if (c->rank > 0) {
                NC(g_nccl.Send(base + 1 * c->ps, n, dt, c->rank - 1, c->comm, c->cst));
                NC(g_nccl.Recv(base + 0 * c->ps, n, dt, c->rank - 1, c->comm, c->cst));
            }
            if (c->rank < c->nranks - 1) {
                NC(g_nccl.Send(base + (long long)c->cfg.nxl * c->ps, n, dt, c->rank + 1, c->comm, c->cst));
                NC(g_nccl.Recv(base + (long long)(c->cfg.nxl + 1) * c->ps, n, dt, c->rank + 1, c->comm, c->cst));
            }
        }
        NC(g_nccl.GroupEnd());
        c->launches++;
        return 0;
    }

    int step() override {
        if (c->w && c->cfg.x0 == 0 && c->tt - c->w_base >= c->nw)
            return fail("source table covers steps %lld..%lld, step %lld requested", c->w_base, c->w_base + c->nw - 1, c->tt);
        // single GPU: replay the step as a CUDA graph (captured per rotation phase after a few plain steps)
        // (slabs too when the halos are pushed from inside the stencil kernel: the flag kernels take the step number from
        // device memory; the NCCL path is launched call by call)
        const bool graphable = c->graph_mode && (c->nranks == 1 || c->halo == 2) && !c->prof;
        if (graphable && c->plain_steps >= 3) {
            const int ph = c->cur;
            if (!c->gexec[ph]) {
                cudaGraph_t gr = nullptr;
                const long long l0 = c->launches.load();
                CU(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
                const int rc = launch_step();
                const cudaError_t ce = cudaStreamEndCapture(c->st, &gr);
                c->gnodes[ph] = (int)(c->launches.load() - l0);
                c->launches -= c->gnodes[ph];
                if (rc != 0 || ce != cudaSuccess || !gr) {
                    if (gr) cudaGraphDestroy(gr);
                    cudaGetLastError();
                    c->graph_mode = 0;                 // capture refused: keep launching kernel by kernel
                    if (rc != 0) return rc;
                } else {
                    const cudaError_t ie = cudaGraphInstantiate(&c->gexec[ph], gr, 0);
                    cudaGraphDestroy(gr);
                    if (ie != cudaSuccess) { cudaGetLastError(); c->gexec[ph] = nullptr; c->graph_mode = 0; }
                }
            }
            if (c->gexec[ph]) {
                CU(cudaGraphLaunch(c->gexec[ph], c->st));
                c->launches += c->gnodes[ph];
                advance();
                if (c->bloch_role == 1) c->partner->eng->advance();
                return 0;
            }
        }
        OK(launch_step());
        c->plain_steps++;
        advance();          // rotate: old <- cur, cur <- new
        if (c->bloch_role == 1) c->partner->eng->advance();
        return 0;
    }
    void advance() override {
        c->cur = b_new();
        c->tt++;
    }
    // ---- Bloch pair ------------------------------------------------------------------------------------------
    // The follower's launches go to the PRIMARY's streams (program order = execution order, and the primary's CUDA graph
    // of the step contains them): its stream / event handles are swapped for the primary's for the duration of the call.
    struct StreamSwap {
        phb_ctx *f, *p;
        cudaStream_t st, zst, kst, est, ezst;
        cudaEvent_t a, b, k;
        StreamSwap(phb_ctx *f_, phb_ctx *p_) : f(f_), p(p_), st(f_->st), zst(f_->zst), kst(f_->kst), est(f_->est), ezst(f_->ezst),
                                               a(f_->ev_fork), b(f_->ev_join), k(f_->ev_kjoin) {
            f->st = p->st; f->zst = p->zst; f->kst = p->kst; f->est = p->est; f->ezst = p->ezst;
            f->ev_fork = p->ev_fork; f->ev_join = p->ev_join; f->ev_kjoin = p->ev_kjoin;
        }
        ~StreamSwap() { f->st = st; f->zst = zst; f->kst = kst; f->est = est; f->ezst = ezst; f->ev_fork = a; f->ev_join = b; f->ev_kjoin = k; }
    };
    int follower_physics() override {
        StreamSwap sw(c, c->partner);
        return physics(c->cfg.x0, c->cfg.x0 + c->cfg.nxl);
    }
    int follower_faces() override {
        StreamSwap sw(c, c->partner);
        return faces(c->cfg.x0, c->cfg.x0 + c->cfg.nxl, true);
    }
    StepArgs<T> pbc_args() {
        StepArgs<T> p;
        for (int q = 0; q < 3; ++q) p.push_lo[q] = p.push_hi[q] = nullptr;
        p.edge_b = -1; p.zface = 0; p.ztile0 = 0;
        p.zf_cl = p.zf_ct = (T)0;
        p.g = geo();
        p.cur = fld(b_cur()); p.old = fld(b_old()); p.nw = fld(b_new());
        p.line_save = (c->w && c->cfg.x0 == 0) ? (const T *)c->line_save : nullptr;
        p.i_begin = c->cfg.x0; p.i_end = c->cfg.x0 + c->cfg.nxl;
        return p;
    }
    int launch_step_bloch() {
        auto *fe = static_cast<Engine<T> *>(c->partner->eng);
        const int x0 = c->cfg.x0, xe = x0 + c->cfg.nxl;
        if (c->w) {
            k_source<T><<<1, 256, 0, c->st>>>(geo(), (T *)c->buf[b_cur()][2], (T *)c->line_save, c->w, c->src_idx);
            c->launches++;
        }
        OK(physics(x0, xe));
        OK(fe->follower_physics());
        {
            StepArgs<T> pa = pbc_args(), pb = fe->pbc_args();
            MatCls<T> m = mat();
            dim3 bl = block_for(c->cfg.nz), gr = grid3(c->cfg.nz, xe - x0, 1, bl);
            if (c->cfg.arith == PHB_EXACT) k_pbc_y_bloch<Ar<T, true>, MatCls<T>><<<gr, bl, 0, c->st>>>(pa, pb, m, (T)c->bloch_c, (T)c->bloch_s);
            else k_pbc_y_bloch<Ar<T, false>, MatCls<T>><<<gr, bl, 0, c->st>>>(pa, pb, m, (T)c->bloch_c, (T)c->bloch_s);
            c->launches++;
            CU(cudaGetLastError());
        }
        OK(faces(x0, xe, true));
        OK(fe->follower_faces());
        return 0;
    }

    // the launches of one time step on c->st (does not advance tt / rotate)
    int launch_step() {
        if (c->bloch_role == 2) return fail("this context is the imaginary part of a Bloch pair: run the real part");
        if (c->bloch_role == 1) return launch_step_bloch();
        const int x0 = c->cfg.x0, xe = c->cfg.x0 + c->cfg.nxl;
        const bool last = (xe == c->cfg.nx);
        if (c->w && x0 == 0) {
            k_source<T><<<1, 256, 0, c->st>>>(geo(), (T *)c->buf[b_cur()][2], (T *)c->line_save, c->w, c->src_idx);
            c->launches++;
        }
        if (c->halo == 2 && c->nranks > 1) {
            // halo exchange over NVLink peer memory (CUDA IPC), no collective call; two modes (phb_p2p_mode)
            const bool hasL = c->rank > 0, hasR = c->rank < c->nranks - 1;
            if (!c->push_fused) {
                // default: wait for the neighbours' flags of the previous step (their edge planes are in my ghost planes) ->
                // stencil (one launch, or the split parts) -> periodic fix-up -> all faces of the owned planes -> push the two
                // finished edge planes into the neighbours' ghost planes, publish the step number, count the step
                int *const fl = hasL ? c->flags + 0 : nullptr, *const fr = hasR ? c->flags + 1 : nullptr;
                const int nxl = xe - x0, cl = (nxl + 3) / 4;
                if (c->overlap && use_march() && nxl >= 384) {
                    // The first and the last quarter of the slab read a ghost plane: they run on the edge lane, behind the wait
                    // for the neighbours' flags of the previous step; the middle half does not and starts at once, so the
                    // neighbours' push, the flag round trip and the skew between the ranks hide behind half a step of work.
                    cudaStream_t L = c->prof ? c->st : c->est;       // (profiling pass: one lane, so the per-launch times add up)
                    if (!c->prof) {
                        CU(cudaEventRecord(c->ev_lane, c->st));
                        CU(cudaStreamWaitEvent(c->est, c->ev_lane, 0));
                    }
                    k_wait_flags<<<1, 1, 0, L>>>(fl, fr, c->flags + 8, 0, 0, 0ULL);
                    OK(physics(x0, x0 + cl, xe - cl, c->prof ? 0 : 1));
                    if (!c->prof) CU(cudaEventRecord(c->ev_lane_done, c->est));
                    OK(physics(x0 + cl, xe - cl));
                    if (!c->prof) CU(cudaStreamWaitEvent(c->st, c->ev_lane_done, 0));
                } else {
                    k_wait_flags<<<1, 1, 0, c->st>>>(fl, fr, c->flags + 8, 0, 0, 0ULL);
                    OK(physics(x0, xe));
                }
                c->launches++;
                OK(pbc_y(x0, xe));
                OK(faces(x0, xe, last));
                PushArgs a{};
                const int bn = b_new();
                for (int q = 0; q < 3; ++q) {
                    const char *mine = (const char *)c->buf[bn][q];
                    a.src[0][q] = mine + (size_t)1 * c->ps * c->esz;                      // first owned plane (local plane 1)
                    a.src[1][q] = mine + (size_t)c->cfg.nxl * c->ps * c->esz;             // last owned plane
                    if (hasL) a.dst[0][q] = (char *)c->peer_buf[0][bn] + ((size_t)q * (c->peer_nxl[0] + 2) + (c->peer_nxl[0] + 1)) * c->ps * c->esz;
                    if (hasR) a.dst[1][q] = (char *)c->peer_buf[1][bn] + ((size_t)q * (c->peer_nxl[1] + 2)) * c->ps * c->esz;
                }
                a.vecs = (long long)c->ps * (long long)c->esz / 16;
                a.flag[0] = hasL ? c->peer_flags[0] + 1 : nullptr;
                a.flag[1] = hasR ? c->peer_flags[1] + 0 : nullptr;
                a.steps_done = c->flags + 8;
                a.arrive = (unsigned *)(c->flags + 9);
                k_push_signal<<<444, 256, 0, c->st>>>(a);      // 3 blocks per SM: ~5 vectors per thread and plane, loads in flight together
                c->launches++;
                CU(cudaGetLastError());
                return 0;
            }
            // mode "fused": ONE stencil launch pushes the edge planes into the neighbours' ghost planes itself; then a flag
            // write / wait, and the y / z absorbing faces are applied to the owned planes AND to the received ghost planes
            // (same formula and inputs as on the owning rank, so the result stays bit-identical)
            if (!use_march()) return fail("halo=fused needs the marching kernel");
            if (periodic_y()) return fail("periodic y boundaries: use halo mode p2p or nccl (the fix-up rows are final only after the stencil kernel)");
            OK(physics(x0, xe));
            k_signal<<<1, 1, 0, c->st>>>(hasL ? c->peer_flags[0] + 1 : nullptr, hasR ? c->peer_flags[1] + 0 : nullptr, c->flags + 8);
            k_wait_flags<<<1, 1, 0, c->st>>>(hasL ? c->flags + 0 : nullptr, hasR ? c->flags + 1 : nullptr, c->flags + 8, 1, 1, 0ULL);
            c->launches += 2;
            OK(faces(x0 - (hasL ? 1 : 0), xe + (hasR ? 1 : 0), last));
            return 0;
        }
        static const int fake_edges = getenv("PHB_DEBUG_FAKE_EDGES") ? atoi(getenv("PHB_DEBUG_FAKE_EDGES")) : 0;   // timing aid
        if ((c->comm && c->nranks > 1) || fake_edges) {
            // edge planes first so that the exchange overlaps the interior update
            const bool hasL = c->rank > 0 || (fake_edges & 1), hasR = c->rank < c->nranks - 1 || (fake_edges & 2);
            int ib = x0, ie = xe;
            if (hasL && hasR && use_march()) {
                OK(physics(x0, x0 + 1, xe - 1));      // both edge planes in one launch
                OK(pbc_y(x0, x0 + 1));
                OK(pbc_y(xe - 1, xe));
                OK(abc_yz(x0, x0 + 1));
                OK(abc_yz(xe - 1, xe));
                ib = x0 + 1; ie = xe - 1;
            } else {
                if (hasL) { OK(physics(x0, x0 + 1)); OK(pbc_y(x0, x0 + 1)); OK(abc_yz(x0, x0 + 1)); ib = x0 + 1; }
                if (hasR) { OK(physics(xe - 1, xe)); OK(pbc_y(xe - 1, xe)); OK(abc_yz(xe - 1, xe)); ie = xe - 1; }
            }
            CU(cudaEventRecord(c->ev_edge, c->st));
            CU(cudaStreamWaitEvent(c->cst, c->ev_edge, 0));
            if (c->comm) OK(exchange());
            CU(cudaEventRecord(c->ev_comm, c->cst));
            OK(physics(ib, ie));
            OK(pbc_y(ib, ie));
            OK(faces(ib, ie, last));
            CU(cudaStreamWaitEvent(c->st, c->ev_comm, 0));
        } else {
            OK(physics(x0, xe));
            OK(pbc_y(x0, xe));
            OK(faces(x0, xe, last));
        }
        return 0;
    }
};

// ------------------------------------------------------------------------------------------
// recorder
// ------------------------------------------------------------------------------------------
static int record_frame(phb_ctx *c) {
    // flow control: wait for a free slot -- bounded, and the consumer (or phb_cancel) can break it: the reference
    // bounds both sides of its writer queue the same way (base_solver.py:89-92,148,274)
    switch (c->rec.wait_free(&c->cancel)) {
        case 0: break;
        case 1: return fail("recording aborted: %s", c->rec.why().c_str());
        case 3: return 0;      // cancelled: run_locked reports it
        default: return fail("recorder ring full for %d ms: the frame consumer has stalled (phb_record_timeout)", c->rec.timeout_ms.load());
    }
    const long long f = c->rec.produced.load();
    const int s = (int)(f % c->rec.slots);
    double *slot = c->ring_dev + (long long)s * c->rec.frame_doubles;
    const int npx = std::max(0, std::min(c->cfg.x0 + c->cfg.nxl, c->cfg.nx - 1) - c->cfg.x0), npyz = c->cfg.nxl;
    if (c->cfg.record_mask & PHB_REC_FULL) {
        // whole arrays in the reference's shapes (or every (sx, sy, sz)-th entry of them), one gather per component
        const int ny = c->cfg.ny, nz = c->cfg.nz;
        const int sx = c->cfg.record_stride[0], sy = c->cfg.record_stride[1], sz = c->cfg.record_stride[2];
        const int np[3] = {npx, npyz, npyz}, ey[3] = {ny, ny - 1, ny}, ez[3] = {nz, nz, nz - 1};
        double *o = slot;
        for (int comp = 0; comp < 3; ++comp) {
            if (!(c->cfg.record_mask & (1 << comp))) continue;
            const int npd = (np[comp] + sx - 1) / sx, eyd = (ey[comp] + sy - 1) / sy, ezd = (ez[comp] + sz - 1) / sz;
            const long long cnt = (long long)npd * eyd * ezd;
            if (cnt > 0) {
                dim3 bl = block_for(ezd), gr = grid3(ezd, eyd, npd, bl);
                if (sx == 1 && sy == 1 && sz == 1) {
                    if (c->cfg.dtype == PHB_F64)
                        k_gather<double><<<gr, bl, 0, c->st>>>((const double *)c->buf[c->cur][comp], o, np[comp], ey[comp], ez[comp], 1, ny, c->nzp);
                    else
                        k_gather<float><<<gr, bl, 0, c->st>>>((const float *)c->buf[c->cur][comp], o, np[comp], ey[comp], ez[comp], 1, ny, c->nzp);
                } else if (c->cfg.dtype == PHB_F64)
                    k_gather_strided<double><<<gr, bl, 0, c->st>>>((const double *)c->buf[c->cur][comp], o, npd, eyd, ezd, 1, ny, c->nzp, sx, sy, sz);
                else
                    k_gather_strided<float><<<gr, bl, 0, c->st>>>((const float *)c->buf[c->cur][comp], o, npd, eyd, ezd, 1, ny, c->nzp, sx, sy, sz);
                c->launches++;
            }
            o += cnt;
        }
    } else {
        dim3 bl(128, 1, 1), gr((c->cfg.ny + 127) / 128, c->cfg.nxl, 1);
        if (c->cfg.dtype == PHB_F64) {
            auto *e = static_cast<Engine<double> *>(c->eng);
            k_record<double><<<gr, bl, 0, c->st>>>(e->geo(), e->fld(c->cur), c->cfg.record_mask, slot, npx, npyz);
        } else {
            auto *e = static_cast<Engine<float> *>(c->eng);
            k_record<float><<<gr, bl, 0, c->st>>>(e->geo(), e->fld(c->cur), c->cfg.record_mask, slot, npx, npyz);
        }
        c->launches++;
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->stage_ev[s], c->st));
    CU(cudaStreamWaitEvent(c->rst, c->stage_ev[s], 0));
    CU(cudaMemcpyAsync(c->rec.host + (long long)s * c->rec.frame_doubles, slot, (size_t)c->rec.frame_doubles * sizeof(double),
                       cudaMemcpyDeviceToHost, c->rst));
    CU(cudaEventRecord(c->rec.slot_ev[s], c->rst));
    c->rec.slot_tt[s] = c->tt - 1;
    c->rec.produced.fetch_add(1);
    return 0;
}

// ------------------------------------------------------------------------------------------
// line probes
// ------------------------------------------------------------------------------------------
template <class T> static inline const T *fld_comp(const Fld<T> &f, int comp) { return comp == 0 ? f.ux : comp == 1 ? f.uy : f.uz; }
static int probe_sample(phb_ctx *c) {
    for (auto &pr : c->probes) {
        if (pr.rows <= 0) continue;
        if (pr.frames >= pr.cap) return fail("probe is full (%lld samples)", pr.cap);
        double *dst = pr.trace + pr.frames * pr.rows;
        const int bl = 128, gr = (pr.rows + bl - 1) / bl;
        if (c->cfg.dtype == PHB_F64) {
            auto *e = static_cast<Engine<double> *>(c->eng);
            k_probe_sample<double><<<gr, bl, 0, c->st>>>(e->geo(), fld_comp(e->fld(c->cur), pr.comp), pr.j, pr.k, pr.rows, dst);
        } else {
            auto *e = static_cast<Engine<float> *>(c->eng);
            k_probe_sample<float><<<gr, bl, 0, c->st>>>(e->geo(), fld_comp(e->fld(c->cur), pr.comp), pr.j, pr.k, pr.rows, dst);
        }
        c->launches++;
        pr.frames++;
    }
    CU(cudaGetLastError());
    return 0;
}

// windowed DFT along t of rows [row0, row0 + nrows) -> device A[nrows][nf]; the caller frees *A_out, *tw_out
static int probe_dft_t_dev(phb_ctx *c, const phb_ctx::Probe &pr, const double *window, long long nf, long long row0,
                           long long nrows, double2 **A_out) {
    const long long n = pr.frames;
    if (n < 1 || nf < 1 || nf > n) return fail("bad spectrum size: %lld frames, %lld frequencies", n, nf);
    if (row0 < 0 || nrows < 0 || row0 + nrows > pr.rows) return fail("probe rows [%lld, %lld) outside [0, %d)", row0, row0 + nrows, pr.rows);
    if (n >= (1LL << 30)) return fail("too many frames");
    double *win = nullptr;
    double2 *tw = nullptr, *A = nullptr;
    CU(cudaMallocAsync((void **)&win, (size_t)n * sizeof(double), c->st));
    CU(cudaMallocAsync((void **)&tw, (size_t)n * sizeof(double2), c->st));
    CU(cudaMallocAsync((void **)&A, (size_t)std::max(1LL, nrows * nf) * sizeof(double2), c->st));
    CU(cudaMemcpyAsync(win, window, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->st));
    k_twiddle<<<(unsigned)((n + 255) / 256), 256, 0, c->st>>>(tw, (int)n);
    if (nrows > 0) {
        dim3 bl(32, 8), gr((unsigned)((nrows + 31) / 32), (unsigned)((nf + 7) / 8));
        k_dft_t<<<gr, bl, 0, c->st>>>(pr.trace, pr.rows, (int)n, win, tw, (int)nf, (int)row0, (int)nrows, A);
    }
    c->launches += 2;
    CU(cudaGetLastError());
    CU(cudaFreeAsync(win, c->st));
    CU(cudaFreeAsync(tw, c->st));
    *A_out = A;
    return 0;
}

static int split_to_host(phb_ctx *c, const double2 *dev, long long count, double *re, double *im) {
    std::vector<double2> h((size_t)count);
    CU(cudaMemcpyAsync(h.data(), dev, (size_t)count * sizeof(double2), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    for (long long q = 0; q < count; ++q) { re[q] = h[(size_t)q].x; im[q] = h[(size_t)q].y; }
    return 0;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int phb_version(void) { return PHB_VERSION; }
const char *phb_last_error(void) { return g_err.c_str(); }

int phb_device_count(int *n) {
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) { *n = 0; return fail("cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    return 0;
}

int phb_create(const phb_cfg *cfg, phb_ctx **out) {
    if (!cfg || !out) return fail("null argument");
    *out = nullptr;
    if (cfg->nx < 4 || cfg->ny < 4 || cfg->nz < 4) return fail("grid must be at least 4 points per axis (got %d x %d x %d)", cfg->nx, cfg->ny, cfg->nz);
    if (cfg->x0 < 0 || cfg->nxl < 1 || cfg->x0 + cfg->nxl > cfg->nx) return fail("bad slab [%d, %d) of %d", cfg->x0, cfg->x0 + cfg->nxl, cfg->nx);
    if (cfg->dtype != PHB_F32 && cfg->dtype != PHB_F64) return fail("bad dtype %d", cfg->dtype);
    if (cfg->arith != PHB_FAST && cfg->arith != PHB_EXACT && cfg->arith != PHB_COMP) return fail("bad arith %d", cfg->arith);
    if (cfg->bc_y != PHB_BC_ABSORBING && cfg->bc_y != PHB_BC_PERIODIC) return fail("bad bc_y %d", cfg->bc_y);
    if (cfg->bc_y == PHB_BC_PERIODIC && cfg->arith == PHB_COMP) return fail("periodic y boundaries are not available with the compensated state");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail("no CUDA device available (%s); libphb200 has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail("device %d out of range (%d devices)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail("device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor);

    phb_ctx *c = new phb_ctx();
    c->cfg = *cfg;
    if (c->cfg.record_every < 1) c->cfg.record_every = 1;
    c->device = cfg->device;
    c->esz = cfg->dtype == PHB_F64 ? 8 : 4;
    c->nzp = (cfg->nz + 31) / 32 * 32;
    c->ps = (long long)cfg->ny * c->nzp;
    auto cleanup = [&](int r) { phb_destroy(c); return r; };
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) return cleanup(fail("stream create failed"));
    {   // the halo exchange must get SMs ahead of the interior stencil blocks that become runnable at the same moment
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&c->cst, cudaStreamNonBlocking, hi) != cudaSuccess) return cleanup(fail("stream create failed"));
    }
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&c->zst, cudaStreamNonBlocking, hi) != cudaSuccess) return cleanup(fail("stream create failed"));
        if (cudaStreamCreateWithPriority(&c->kst, cudaStreamNonBlocking, hi) != cudaSuccess) return cleanup(fail("stream create failed"));
        cudaEventCreateWithFlags(&c->ev_kjoin, cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithFlags(&c->est, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithPriority(&c->ezst, cudaStreamNonBlocking, hi) != cudaSuccess) return cleanup(fail("stream create failed"));
    }
    for (cudaEvent_t *ev : {&c->ev_efork, &c->ev_ejoin, &c->ev_lane, &c->ev_lane_done}) cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    if (const char *e = getenv("PHB_OVERLAP")) c->overlap = atoi(e) != 0;
    // three specialised launches per step (fp64 1.666 -> 1.607 ms, fp32 0.842 -> 0.829 ms at 512^3, with 64-byte TMA promotion) -- but
    // only where the two extra launches and their fork / join (~10-15 us) pay: measured on a 512 x 512 cross-section fp64 gains down
    // to 128-plane slabs (0.4345 vs 0.4379 ms), fp32 only at 512 planes (256: 0.445 vs 0.430); on 384^2 / 320^2 cross-sections the
    // parts no longer fill their waves (320^3: 0.471 vs 0.447 ms) -> auto, see physics()
    c->zsplit = -1;
    if (const char *e = getenv("PHB_ZSPLIT")) c->zsplit = atoi(e);
    if (const char *e = getenv("PHB_FACES_FUSED")) c->faces_fused = atoi(e) != 0;
    cudaEventCreateWithFlags(&c->ev_edge, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming);
    cudaEventCreate(&c->ev_t0);
    cudaEventCreate(&c->ev_t1);
    const size_t fb = (size_t)(cfg->nxl + 2) * c->ps * c->esz;
    for (int b = 0; b < 3; ++b) {          // ux, uy, uz of a buffer back to back (one 4-D TMA descriptor covers them)
        if (dmalloc(c, &c->buf[b][0], 3 * fb)) return cleanup(1);
        c->buf[b][1] = (char *)c->buf[b][0] + fb;
        c->buf[b][2] = (char *)c->buf[b][0] + 2 * fb;
    }
    if (dmalloc(c, &c->line_save, (size_t)cfg->ny * c->esz)) return cleanup(1);
    if (cfg->dtype == PHB_F64) c->eng = new Engine<double>(c); else c->eng = new Engine<float>(c);
    if (const char *e = getenv("PHB_MARCH_R")) c->mR = atoi(e);
    if (const char *e = getenv("PHB_MARCH_NST")) c->mNST = atoi(e);
    // EXACT arithmetic is bound by its IEEE divisions, not by latency: 16 warps of one row hide them better than 8 warps
    // of two rows (512^3 fp64: 22.5 vs 26.9 ms per step)
    if (cfg->arith == PHB_EXACT) c->mRW = 1;
    if (const char *e = getenv("PHB_MARCH_RW")) c->mRW = atoi(e);
    if (const char *e = getenv("PHB_ZFUSE")) c->zfuse = atoi(e);
    if (const char *e = getenv("PHB_GRAPH")) c->graph_mode = atoi(e) != 0;
    if (c->mRW != 1 && !(c->mRW == 2 && c->mR == 16 && c->mNST == 4)) c->mRW = 1;   // two rows per warp exist for R = 16, NST = 4
    if (const char *e = getenv("PHB_MARCH_CHUNKS")) c->mChunks = atoi(e);
    if (c->mR != 8 && c->mR != 16) return cleanup(fail("PHB_MARCH_R must be 8 or 16"));
    // recorder ring
    if (cfg->record_mask) {
        const int npx = std::max(0, std::min(cfg->x0 + cfg->nxl, cfg->nx - 1) - cfg->x0);
        long long fd = 0;
        const bool full = (cfg->record_mask & PHB_REC_FULL) != 0;      // whole arrays: Grid.freezeData (grid.py:68-77)
        int *rs = c->cfg.record_stride;
        for (int q = 0; q < 3; ++q) if (rs[q] < 1 || !full) rs[q] = 1;
        if (cfg->x0 % rs[0] != 0) return cleanup(fail("record_stride[0] = %d does not divide the slab origin x0 = %d", rs[0], cfg->x0));
        auto cd = [](long long n, int s) { return (n + s - 1) / s; };
        if (cfg->record_mask & PHB_REC_UX) fd += full ? cd(npx, rs[0]) * cd(cfg->ny, rs[1]) * cd(cfg->nz, rs[2]) : (long long)npx * cfg->ny;
        if (cfg->record_mask & PHB_REC_UY) fd += full ? cd(cfg->nxl, rs[0]) * cd(cfg->ny - 1, rs[1]) * cd(cfg->nz, rs[2]) : (long long)cfg->nxl * (cfg->ny - 1);
        if (cfg->record_mask & PHB_REC_UZ) fd += full ? cd(cfg->nxl, rs[0]) * cd(cfg->ny, rs[1]) * cd(cfg->nz - 1, rs[2]) : (long long)cfg->nxl * cfg->ny;
        if (fd <= 0) return cleanup(fail("record_mask %d selects no component", cfg->record_mask));
        c->rec.frame_doubles = fd;
        c->rec.slots = cfg->ring_slots > 0 ? cfg->ring_slots : 16;
        c->rec.device = c->device;
        if (const char *e = getenv("PHB_REC_TIMEOUT_MS")) c->rec.timeout_ms.store(atoi(e));
        if (cudaHostAlloc((void **)&c->rec.host, (size_t)c->rec.slots * fd * sizeof(double), cudaHostAllocDefault) != cudaSuccess)
            return cleanup(fail("cudaHostAlloc of the %d-slot recorder ring failed", c->rec.slots));
        if (dmalloc(c, (void **)&c->ring_dev, (size_t)c->rec.slots * fd * sizeof(double), false)) return cleanup(1);
        if (cudaStreamCreateWithFlags(&c->rst, cudaStreamNonBlocking) != cudaSuccess) return cleanup(fail("stream create failed"));
        c->rec.slot_ev.resize(c->rec.slots);
        c->stage_ev.resize(c->rec.slots);
        for (auto &ev : c->stage_ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        c->rec.slot_tt.assign(c->rec.slots, -1);
        c->rec.slot_done.assign(c->rec.slots, 0);
        for (auto &ev : c->rec.slot_ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    }
    if (cudaStreamSynchronize(c->st) != cudaSuccess) return cleanup(fail("device init failed: %s", cudaGetErrorString(cudaGetLastError())));
    *out = c;
    return 0;
}

int phb_destroy(phb_ctx *c) {
    if (!c) return 0;
    // writer threads read the pinned ring: stop and join them before anything is freed (a context destroyed with a
    // live consumer was a use-after-free); a caller blocked in phb_record_next sees the abort and returns
    if (c->partner) {           // unlink a Bloch pair: the other part becomes an ordinary periodic context again
        cudaStreamSynchronize(c->st);
        if (c->partner->st) cudaStreamSynchronize(c->partner->st);
        graph_invalidate(c->partner);
        c->partner->partner = nullptr;
        c->partner->bloch_role = 0;
        c->partner = nullptr;
    }
    if (c->ev_pair) cudaEventDestroy(c->ev_pair);
    c->rec.abort("context destroyed");
    c->wr.finish(2000);
    for (int spins = 0; c->rec.in_api.load() > 0 && spins < 40000; ++spins)      // a consumer inside phb_record_next sees the abort and leaves (<= 2 s)
        std::this_thread::sleep_for(std::chrono::microseconds(50));
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->cst) cudaStreamSynchronize(c->cst);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (int s = 0; s < 2; ++s) {
        for (int b = 0; b < 3; ++b) if (c->peer_buf[s][b]) cudaIpcCloseMemHandle(c->peer_buf[s][b]);
        if (c->peer_flags[s]) cudaIpcCloseMemHandle(c->peer_flags[s]);
    }
    cudaFree(c->flags);
    for (int b = 0; b < 3; ++b) cudaFree(c->buf[b][0]);
    for (int a = 0; a < 6; ++a) cudaFree(c->sp[a]);
    cudaFree(c->tab); cudaFree(c->ids); cudaFree(c->code); cudaFree(c->line_save); cudaFree(c->dbg_push);
    graph_invalidate(c);
    if (c->w) cudaFree(c->w);
    cudaFree(c->src_idx);
    if (c->rec.host) cudaFreeHost(c->rec.host);
    for (auto &ev : c->rec.slot_ev) cudaEventDestroy(ev);
    for (auto &ev : c->stage_ev) cudaEventDestroy(ev);
    if (c->rst) { cudaStreamSynchronize(c->rst); cudaStreamDestroy(c->rst); }
    cudaFree(c->ring_dev);
    for (auto &pr : c->probes) cudaFree(pr.trace);
    for (auto &pe : c->prof_ev) { cudaEventDestroy(pe.first); cudaEventDestroy(pe.second); }
    if (c->zst) { cudaStreamSynchronize(c->zst); cudaStreamDestroy(c->zst); }
    if (c->kst) { cudaStreamSynchronize(c->kst); cudaStreamDestroy(c->kst); }
    if (c->ev_kjoin) cudaEventDestroy(c->ev_kjoin);
    for (cudaStream_t q : {c->est, c->ezst}) if (q) { cudaStreamSynchronize(q); cudaStreamDestroy(q); }
    for (cudaEvent_t ev : {c->ev_efork, c->ev_ejoin, c->ev_lane, c->ev_lane_done}) if (ev) cudaEventDestroy(ev);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_edge) cudaEventDestroy(c->ev_edge);
    if (c->ev_comm) cudaEventDestroy(c->ev_comm);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->st) cudaStreamDestroy(c->st);
    if (c->cst) cudaStreamDestroy(c->cst);
    delete c->eng;
    delete c;
    return 0;
}

#define ENTER(c)                              \
    if (!(c)) return fail("null context");    \
    std::lock_guard<std::mutex> lk_((c)->mu); \
    CU(cudaSetDevice((c)->device));

int phb_set_spacing(phb_ctx *c, const double *fdx, const double *fdy, const double *fdz, const double *sdx,
                    const double *sdy, const double *sdz) {
    ENTER(c);
    graph_invalidate(c);
    if (!fdx || !fdy || !fdz || !sdx || !sdy || !sdz) return fail("null spacing array");
    return c->eng->set_spacing(fdx, fdy, fdz, sdx, sdy, sdz);
}

int phb_set_material_table(phb_ctx *c, int32_t nmat, const double *c12, const double *rho) {
    ENTER(c);
    graph_invalidate(c);
    if (nmat < 1 || nmat > MAX_MAT) return fail("nmat must be 1..%d (got %d)", (int)MAX_MAT, nmat);
    for (int m = 0; m < nmat; ++m)
        if (!(rho[m] > 0)) return fail("density of material %d is not positive", m);
    OK(c->eng->set_table(nmat, c12, rho));
    if (c->ids) OK(c->eng->build_codes());
    return 0;
}

static int ids_planes(const phb_ctx *c, int *ib, int *ie) {
    *ib = c->cfg.x0;
    *ie = std::min(c->cfg.x0 + c->cfg.nxl + 1, c->cfg.nx);
    return *ie - *ib;
}

int phb_set_material_ids(phb_ctx *c, const uint8_t *ids, int64_t nplanes) {
    ENTER(c);
    graph_invalidate(c);
    int ib, ie;
    const int np = ids_planes(c, &ib, &ie);
    if (nplanes != np) return fail("expected %d id planes [%d, %d), got %lld", np, ib, ie, (long long)nplanes);
    const size_t bytes = (size_t)np * c->cfg.ny * c->cfg.nz;
    if (!c->ids) OK(dmalloc(c, (void **)&c->ids, bytes, false));
    c->ids_ib = ib; c->ids_ie = ie;
    CU(cudaMemcpyAsync(c->ids, ids, bytes, cudaMemcpyHostToDevice, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (c->nmat) OK(c->eng->build_codes());
    return 0;
}

int phb_set_material_dense(phb_ctx *c, const double *C, const double *P, int64_t nplanes) {
    if (!c) return fail("null context");
    if (!C || !P) return fail("null argument");
    int ib, ie;
    const int np = ids_planes(c, &ib, &ie);
    if (nplanes != np) return fail("expected %d material planes [%d, %d), got %lld", np, ib, ie, (long long)nplanes);
    const size_t cells = (size_t)np * c->cfg.ny * c->cfg.nz;
    // the 12 entries of the 6x6 block the path reads: rows 0..2 x columns 0..2, then (3,3), (4,4), (5,5)
    static const int pick[12] = {0, 1, 2, 6, 7, 8, 12, 13, 14, 21, 28, 35};
    std::vector<double> tab;            // nmat x 13 (12 stiffness entries, density)
    std::vector<uint8_t> ids(cells);
    int nmat = 0, last = 0;
    double key[13];
    for (size_t q = 0; q < cells; ++q) {
        const double *cq = C + q * 36;
        for (int a = 0; a < 12; ++a) key[a] = cq[pick[a]];
        key[12] = P[q];
        int id = -1;
        if (nmat && memcmp(key, &tab[(size_t)last * 13], sizeof key) == 0) id = last;
        for (int m = 0; id < 0 && m < nmat; ++m)
            if (memcmp(key, &tab[(size_t)m * 13], sizeof key) == 0) id = m;
        if (id < 0) {
            if (nmat == MAX_MAT) return fail("more than %d distinct materials in C / P (cell %zu)", (int)MAX_MAT, q);
            tab.insert(tab.end(), key, key + 13);
            id = nmat++;
        }
        ids[q] = (uint8_t)(last = id);
    }
    std::vector<double> c12((size_t)nmat * 12), rho((size_t)nmat);
    for (int m = 0; m < nmat; ++m) {
        std::copy(&tab[(size_t)m * 13], &tab[(size_t)m * 13] + 12, &c12[(size_t)m * 12]);
        rho[(size_t)m] = tab[(size_t)m * 13 + 12];
    }
    if (int r = phb_set_material_table(c, nmat, c12.data(), rho.data())) return r;
    return phb_set_material_ids(c, ids.data(), nplanes);
}

int phb_gen_material_ids(phb_ctx *c, const float *tg, int32_t n, const double *x, const double *y, const double *z) {
    ENTER(c);
    graph_invalidate(c);
    if (n < 0 || (n && !tg) || !x || !y || !z) return fail("bad arguments");
    const int nx = c->cfg.nx, ny = c->cfg.ny, nz = c->cfg.nz;
    int ib, ie;
    const int np = ids_planes(c, &ib, &ie);
    // z profile per inclusion, with the re-binding of `z` to the index array (grid.py:173)
    std::vector<double> zc(z, z + nz);
    std::vector<std::vector<int>> prof(n);
    for (int t = 0; t < n; ++t) {
        std::vector<int> idx;
        for (int m = 0; m < (int)zc.size(); ++m)
            if (zc[m] <= (double)tg[4 * t + 2]) idx.push_back(m);
        prof[t] = idx;
        zc.assign(idx.begin(), idx.end());
    }
    std::vector<std::vector<int>> groups;
    std::vector<int> group(n, 0);
    for (int t = 0; t < n; ++t) {
        int gi = -1;
        for (size_t q = 0; q < groups.size(); ++q)
            if (groups[q] == prof[t]) { gi = (int)q; break; }
        if (gi < 0) {
            if (groups.size() == 32) return fail("more than 32 distinct inclusion depth profiles; upload ids with phb_set_material_ids instead");
            groups.push_back(prof[t]);
            gi = (int)groups.size() - 1;
        }
        group[t] = gi;
    }
    std::vector<unsigned> zbits(nz, 0u);
    for (size_t q = 0; q < groups.size(); ++q)
        for (int m : groups[q])
            if (m >= 0 && m < nz) zbits[m] |= 1u << q;
    const size_t bytes = (size_t)np * ny * nz;
    if (!c->ids) OK(dmalloc(c, (void **)&c->ids, bytes, false));
    c->ids_ib = ib; c->ids_ie = ie;
    double *dx = nullptr, *dy = nullptr;
    float *dt = nullptr;
    int *dg = nullptr;
    unsigned *dcol = nullptr, *dz = nullptr;
    CU(cudaMalloc(&dx, nx * sizeof(double)));
    CU(cudaMalloc(&dy, ny * sizeof(double)));
    CU(cudaMalloc(&dt, (n ? n : 1) * 4 * sizeof(float)));
    CU(cudaMalloc(&dg, (n ? n : 1) * sizeof(int)));
    CU(cudaMalloc(&dcol, (size_t)np * ny * sizeof(unsigned)));
    CU(cudaMalloc(&dz, nz * sizeof(unsigned)));
    CU(cudaMemcpyAsync(dx, x, nx * sizeof(double), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(dy, y, ny * sizeof(double), cudaMemcpyHostToDevice, c->st));
    if (n) {
        CU(cudaMemcpyAsync(dt, tg, (size_t)n * 4 * sizeof(float), cudaMemcpyHostToDevice, c->st));
        CU(cudaMemcpyAsync(dg, group.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->st));
    }
    CU(cudaMemcpyAsync(dz, zbits.data(), nz * sizeof(unsigned), cudaMemcpyHostToDevice, c->st));
    k_incl_columns<<<dim3((ny + 127) / 128, np, 1), 128, 0, c->st>>>(dx, dy, ib, ie, ny, dt, dg, n, dcol);
    dim3 bl = block_for(nz);
    k_incl_fill<<<grid3(nz, ny, np, bl), bl, 0, c->st>>>(dcol, dz, c->ids, np, ny, nz);
    c->launches += 2;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
    cudaFree(dx); cudaFree(dy); cudaFree(dt); cudaFree(dg); cudaFree(dcol); cudaFree(dz);
    CU(e);
    if (c->nmat) OK(c->eng->build_codes());
    return 0;
}

int phb_get_material_ids(phb_ctx *c, uint8_t *ids) {
    ENTER(c);
    if (!c->ids) return fail("material ids not set");
    CU(cudaMemcpyAsync(ids, c->ids, (size_t)c->cfg.nxl * c->cfg.ny * c->cfg.nz, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    return 0;
}

int phb_set_abc(phb_ctx *c, const double coef[8]) {
    ENTER(c);
    graph_invalidate(c);
    return c->eng->set_abc(coef);
}

int phb_set_source_table(phb_ctx *c, const double *w, int64_t n) {
    ENTER(c);
    // The previous table may still be read by enqueued steps: everything here is ordered on the launch stream.
    // The buffer (and the captured graphs that hold its address) is kept while the new table fits.
    if (!w || n <= 0) {
        if (c->w) { CU(cudaFreeAsync(c->w, c->st)); c->w = nullptr; graph_invalidate(c); }
        c->nw = c->w_cap = 0;
        return 0;
    }
    if (!c->src_idx) {
        CU(cudaMalloc((void **)&c->src_idx, sizeof(long long)));
        graph_invalidate(c);
    }
    if (n > c->w_cap) {
        if (c->w) CU(cudaFreeAsync(c->w, c->st));
        c->w = nullptr;
        c->w_cap = std::max<long long>(n, 1024);
        CU(cudaMallocAsync((void **)&c->w, (size_t)c->w_cap * sizeof(double), c->st));
        graph_invalidate(c);
    }
    CU(cudaMemcpyAsync(c->w, w, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->st));   // pageable source: staged before return
    CU(cudaMemsetAsync(c->src_idx, 0, sizeof(long long), c->st));
    c->nw = n;
    c->w_base = c->tt;
    return 0;
}

int phb_set_fields(phb_ctx *c, int32_t which, const double *ux, const double *uy, const double *uz) {
    ENTER(c);
    if (which != PHB_CUR && which != PHB_OLD) return fail("bad buffer selector %d", which);
    return c->eng->xfer(true, which, (double *)ux, (double *)uy, (double *)uz);
}
int phb_get_fields(phb_ctx *c, int32_t which, double *ux, double *uy, double *uz) {
    ENTER(c);
    if (which != PHB_CUR && which != PHB_OLD) return fail("bad buffer selector %d", which);
    return c->eng->xfer(false, which, ux, uy, uz);
}
int phb_get_stress(phb_ctx *c, int32_t which, double *T1, double *T2, double *T3, double *T4, double *T5,
                   double *T6) {
    ENTER(c);
    double *T[6] = {T1, T2, T3, T4, T5, T6};
    return c->eng->stress(which, T);
}

static int run_locked(phb_ctx *c, int64_t nsteps) {
    if (!c->have_spacing) return fail("phb_set_spacing not called");
    if (!c->code) return fail("material not set (table + ids)");
    if (!c->have_abc) return fail("phb_set_abc not called");
    if (c->nranks > 1 && !c->comm && c->halo != 2) return fail("slab context without halo exchange: call phb_comm_init or phb_p2p_import");
    if (c->cfg.record_mask && c->rec.aborted.load()) return fail("recording aborted: %s", c->rec.why().c_str());
    if (c->bloch_role == 2) return fail("this context is the imaginary part of a Bloch pair: run the real part");
    if (c->bloch_role == 1) {        // whatever the follower has queued on its own stream (field uploads) comes first
        if (c->partner->tt != c->tt || c->partner->cur != c->cur) return fail("Bloch pair out of step");
        CU(cudaEventRecord(c->ev_pair, c->partner->st));
        CU(cudaStreamWaitEvent(c->st, c->ev_pair, 0));
    }
    struct PairJoin {      // ... and what the follower does next on its own stream (field downloads) waits for this run
        phb_ctx *c;
        ~PairJoin() {
            if (c->bloch_role == 1 && c->partner) {
                cudaEventRecord(c->ev_pair, c->st);
                cudaStreamWaitEvent(c->partner->st, c->ev_pair, 0);
            }
        }
    } pair_join{c};
    for (int64_t s = 0; s < nsteps; ++s) {
        if (c->cancel.load()) {        // phb_cancel from another thread (BaseSolver.cancel, base_solver.py:246-248,282-284)
            c->cancel.store(0);
            g_err = "cancelled";
            return 3;
        }
        OK(c->eng->step());
        if ((c->tt % c->cfg.record_every) == 0) {
            if (c->cfg.record_mask) OK(record_frame(c));
            if (!c->probes.empty()) OK(probe_sample(c));
        }
    }
    return 0;
}

int phb_run(phb_ctx *c, int64_t nsteps) {
    ENTER(c);
    return run_locked(c, nsteps);
}
int phb_sync(phb_ctx *c) {
    ENTER(c);
    if (c->halo == 2 && !c->push_fused && c->nranks > 1 && c->flags) {
        // the neighbours' pushes of every completed step have landed in my ghost planes (so that a neighbour that is a
        // little behind never stores into memory I have released); bounded: a neighbour that is gone must not hang us
        k_wait_flags<<<1, 1, 0, c->st>>>(c->rank > 0 ? c->flags + 0 : nullptr, c->rank < c->nranks - 1 ? c->flags + 1 : nullptr,
                                         c->flags + 8, 0, 0, 10000000000ULL);
    }
    CU(cudaStreamSynchronize(c->st));
    CU(cudaStreamSynchronize(c->cst));
    if (c->rst) CU(cudaStreamSynchronize(c->rst));
    return 0;
}
int phb_run_timed(phb_ctx *c, int64_t nsteps, float *ms) {
    ENTER(c);
    CU(cudaStreamSynchronize(c->st));
    CU(cudaStreamSynchronize(c->cst));
    CU(cudaEventRecord(c->ev_t0, c->st));
    OK(run_locked(c, nsteps));
    CU(cudaEventRecord(c->ev_t1, c->st));
    CU(cudaEventSynchronize(c->ev_t1));
    CU(cudaStreamSynchronize(c->cst));
    CU(cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
    return 0;
}
int phb_steps_done(phb_ctx *c, int64_t *tt) {
    if (!c) return fail("null context");
    *tt = c->tt;
    return 0;
}
int phb_launch_count(phb_ctx *c, int64_t *n) {
    if (!c) return fail("null context");
    *n = c->launches.load();
    return 0;
}
int phb_info(phb_ctx *c, char *name, int32_t len, int64_t *bytes) {
    if (!c) return fail("null context");
    if (name && len > 0) snprintf(name, len, "%s", c->eng->kernel_name());
    if (bytes) *bytes = c->dev_bytes;
    return 0;
}

int phb_profile(phb_ctx *c, int32_t enable, double *kernel_ms, int64_t *kernel_launches) {
    ENTER(c);
    graph_invalidate(c);
    OK(prof_collect(c));
    if (kernel_ms) *kernel_ms = c->prof_ms;
    if (kernel_launches) *kernel_launches = c->prof_n;
    if (enable >= 0) {
        c->prof = enable != 0;
        c->prof_ms = 0;
        c->prof_n = 0;
    }
    return 0;
}

int phb_comm_unique_id(char id[128]) {
    OK(nccl_load());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId u;
    NC(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, 128);
    return 0;
}
int phb_comm_init(phb_ctx *c, const char id[128], int32_t rank, int32_t nranks) {
    ENTER(c);
    graph_invalidate(c);
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail("bad rank %d of %d", rank, nranks);
    c->rank = rank; c->nranks = nranks;
    if (nranks == 1) return 0;
    if (c->cfg.nxl < 4) return fail("a slab needs at least 4 planes (got %d)", c->cfg.nxl);
    c->halo = 1;
    OK(nccl_load());
    ncclUniqueId u;
    memcpy(&u, id, 128);
    NC(g_nccl.CommInitRank(&c->comm, nranks, u, rank));
    // NCCL sets its peer connections up lazily on first use: do one tiny exchange with both neighbours now so
    // that the first time step does not pay for it (Solver.init absorbs it, Solver.run does not)
    {
        double *tmp = nullptr;
        CU(cudaMalloc(&tmp, 4 * sizeof(double)));
        NC(g_nccl.GroupStart());
        if (rank > 0) {
            NC(g_nccl.Send(tmp + 0, 1, ncclDouble, rank - 1, c->comm, c->cst));
            NC(g_nccl.Recv(tmp + 1, 1, ncclDouble, rank - 1, c->comm, c->cst));
        }
        if (rank < nranks - 1) {
            NC(g_nccl.Send(tmp + 2, 1, ncclDouble, rank + 1, c->comm, c->cst));
            NC(g_nccl.Recv(tmp + 3, 1, ncclDouble, rank + 1, c->comm, c->cst));
        }
        NC(g_nccl.GroupEnd());
        cudaError_t e = cudaStreamSynchronize(c->cst);
        cudaFree(tmp);
        CU(e);
    }
    return 0;
}

int phb_p2p_export(phb_ctx *c, char handles[256], int32_t *nxl) {
    ENTER(c);
    if (!c->flags) {
        OK(dmalloc(c, (void **)&c->flags, 64));
        CU(cudaStreamSynchronize(c->st));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    for (int b = 0; b < 3; ++b) CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)(handles + 64 * b), c->buf[b][0]));
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)(handles + 192), c->flags));
    *nxl = c->cfg.nxl;
    return 0;
}

int phb_p2p_import(phb_ctx *c, int32_t rank, int32_t nranks, const char *left, int32_t left_nxl, const char *right,
                   int32_t right_nxl) {
    ENTER(c);
    graph_invalidate(c);
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail("bad rank %d of %d", rank, nranks);
    if (!c->flags) return fail("call phb_p2p_export first");

    c->rank = rank; c->nranks = nranks;
    if (nranks == 1) return 0;
    if (c->cfg.nxl < 4) return fail("a slab needs at least 4 planes (got %d)", c->cfg.nxl);
    const char *src[2] = {rank > 0 ? left : nullptr, rank < nranks - 1 ? right : nullptr};
    const int nx2[2] = {left_nxl, right_nxl};
    for (int s = 0; s < 2; ++s) {
        if (!src[s]) continue;
        cudaIpcMemHandle_t h;
        for (int b = 0; b < 3; ++b) {
            memcpy(&h, src[s] + 64 * b, 64);
            CU(cudaIpcOpenMemHandle(&c->peer_buf[s][b], h, cudaIpcMemLazyEnablePeerAccess));
        }
        memcpy(&h, src[s] + 192, 64);
        CU(cudaIpcOpenMemHandle((void **)&c->peer_flags[s], h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_nxl[s] = nx2[s];
    }
    {   // the flag kernels count steps in device memory from what has been run so far
        const int done = (int)c->tt;
        CU(cudaMemcpyAsync(c->flags + 8, &done, sizeof(int), cudaMemcpyHostToDevice, c->st));
        CU(cudaStreamSynchronize(c->st));
    }
    c->halo = 2;
    return 0;
}

int phb_bloch_pair(phb_ctx *re, phb_ctx *im, double phase) {
    if (!re || !im || re == im) return fail("two distinct contexts are needed");
    std::lock_guard<std::mutex> l1(re->mu);
    std::lock_guard<std::mutex> l2(im->mu);
    const phb_cfg &a = re->cfg, &b = im->cfg;
    if (a.nx != b.nx || a.ny != b.ny || a.nz != b.nz || a.x0 != b.x0 || a.nxl != b.nxl || a.dtype != b.dtype || a.arith != b.arith ||
        a.kernel != b.kernel || a.device != b.device)
        return fail("the two parts of a Bloch pair must be created with the same grid, type, arithmetic, kernel and device");
    if (a.bc_y != PHB_BC_PERIODIC || b.bc_y != PHB_BC_PERIODIC) return fail("both parts must be created with bc_y = PHB_BC_PERIODIC");
    if (a.x0 != 0 || a.nxl != a.nx || re->nranks > 1 || im->nranks > 1) return fail("a Bloch pair runs on one GPU (whole grid)");
    if (re->partner || im->partner) return fail("context already paired");
    if (re->tt != 0 || im->tt != 0) return fail("pair the contexts before the first step");
    if (im->cfg.record_mask) return fail("the imaginary part cannot record (read it with phb_get_fields)");
    CU(cudaSetDevice(re->device));
    if (!re->ev_pair) CU(cudaEventCreateWithFlags(&re->ev_pair, cudaEventDisableTiming));
    graph_invalidate(re);
    graph_invalidate(im);
    re->partner = im; im->partner = re;
    re->bloch_role = 1; im->bloch_role = 2;
    re->bloch_c = im->bloch_c = cos(phase);
    re->bloch_s = im->bloch_s = sin(phase);
    return 0;
}

int phb_p2p_mode(phb_ctx *c, int32_t fused_in_kernel) {
    ENTER(c);
    graph_invalidate(c);
    if (fused_in_kernel && c->cfg.bc_y == PHB_BC_PERIODIC) return fail("periodic y boundaries cannot use the in-kernel halo push");
    c->push_fused = fused_in_kernel ? 1 : 0;
    return 0;
}

int phb_record_frame_doubles(phb_ctx *c, int64_t *n) {
    if (!c) return fail("null context");
    *n = c->rec.frame_doubles;
    return 0;
}
struct InApi {      // marks a consumer call in progress so that phb_destroy does not free the ring under it
    std::atomic<int> &n;
    explicit InApi(std::atomic<int> &n_) : n(n_) { n.fetch_add(1); }
    ~InApi() { n.fetch_sub(1); }
};
int phb_record_next(phb_ctx *c, const double **frame, int64_t *tt, int32_t timeout_ms) {
    if (!c || !c->rec.host) return fail("recording not enabled");
    InApi guard(c->rec.in_api);
    if (c->wr.started) return fail("the native writer is draining the ring (phb_writer_start)");
    RecRing &r = c->rec;
    if (r.consumed.load() != r.released.load()) return fail("previous frame not released");
    auto t0 = std::chrono::steady_clock::now();
    while (r.produced.load() <= r.consumed.load()) {
        if (r.aborted.load()) return fail("recording aborted: %s", r.why().c_str());
        if (std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() >= timeout_ms) {
            *tt = -1; *frame = nullptr;
            g_err = "timeout";
            return 2;   // timeout, not an error
        }
        std::this_thread::sleep_for(std::chrono::microseconds(100));
    }
    const int s = (int)(r.consumed.load() % r.slots);
    cudaSetDevice(c->device);
    CU(cudaEventSynchronize(r.slot_ev[s]));
    *frame = r.host + (long long)s * r.frame_doubles;
    *tt = r.slot_tt[s];
    r.consumed.fetch_add(1);
    return 0;
}
int phb_record_release(phb_ctx *c) {
    if (!c || !c->rec.host) return fail("recording not enabled");
    if (c->rec.released.load() < c->rec.consumed.load()) c->rec.released.fetch_add(1);
    return 0;
}
int phb_record_abort(phb_ctx *c, const char *why) {
    if (!c) return fail("null context");
    c->rec.abort(why && *why ? why : "aborted by the consumer");
    return 0;
}
int phb_record_timeout(phb_ctx *c, int32_t timeout_ms) {
    if (!c) return fail("null context");
    c->rec.timeout_ms.store(timeout_ms > 0 ? timeout_ms : 1);
    return 0;
}
int phb_cancel(phb_ctx *c) {
    if (!c) return fail("null context");
    c->cancel.store(1);
    return 0;
}

// ---- native writer ---------------------------------------------------------------------------------
static int writer_setup(NativeWriter &w, RecRing &r, int fd, int32_t ncomp, const int64_t *base, const int64_t *bytes, int64_t stride,
                        int64_t frames) {
    if (fd < 0 || ncomp < 1 || ncomp > 3 || !base || !bytes || frames < 0) return fail("bad writer arguments");
    long long off = 0;
    for (int q = 0; q < ncomp; ++q) {
        if (bytes[q] <= 0 || base[q] < 0) return fail("bad extent for component %d", q);
        w.base[q] = base[q]; w.bytes[q] = bytes[q]; w.off[q] = off;
        off += bytes[q];
    }
    if (off != r.frame_doubles * 8) return fail("components add up to %lld bytes, a ring frame has %lld", off, r.frame_doubles * 8);
    w.fd = fd; w.ncomp = ncomp; w.stride = stride; w.frames = frames;
    w.wait_us.store(0); w.write_us.store(0);
    return 0;
}
int phb_writer_start(phb_ctx *c, int32_t fd, int32_t ncomp, const int64_t *base, const int64_t *bytes, int64_t stride,
                     int64_t frames, int32_t nthreads, int32_t flags) {
    if (!c || !c->rec.host) return fail("recording not enabled");
    if (c->wr.started) return fail("writer already started");
    if (c->rec.produced.load() != 0) return fail("frames were recorded before the writer started");
    OK(writer_setup(c->wr, c->rec, fd, ncomp, base, bytes, stride, frames));
    if (flags & PHB_WRITER_MMAP) c->wr.map_extents((flags & PHB_WRITER_POPULATE) != 0);     // falls back to pwrite when the mapping is refused
    return c->wr.start(&c->rec, nthreads > 0 ? nthreads : 4);
}
int phb_writer_mapped(phb_ctx *c, int32_t *mapped) {
    if (!c || !mapped) return fail("null argument");
    *mapped = c->wr.map != nullptr;
    return 0;
}
int phb_writer_finish(phb_ctx *c, int32_t timeout_ms, int64_t *written, double *wait_s, double *write_s) {
    if (!c) return fail("null context");
    const bool ok = c->wr.finish(timeout_ms > 0 ? timeout_ms : 300000);      // reference: join(300), base_solver.py:89-92,274
    if (written) *written = c->wr.written.load();
    if (wait_s) *wait_s = 1e-6 * (double)c->wr.wait_us.load();
    if (write_s) *write_s = 1e-6 * (double)c->wr.write_us.load();
    if (!ok) return fail("writer threads did not finish within %d ms", timeout_ms);
    if (c->rec.aborted.load()) return fail("%s", c->rec.why().c_str());
    return 0;
}
// CPU-only exercise of the ring + writer threads (tests): a host thread produces `frames` frames of `frame_doubles`
// doubles, frame f element q = f * 1e6 + q, through a `slots`-deep ring; the writer threads store them at
// base[c] + f * stride.  abort_at >= 0: the producer stops there as if cancelled.  No CUDA call is made.
int phb_writer_selftest(int32_t fd, int32_t ncomp, const int64_t *base, const int64_t *bytes, int64_t stride, int64_t frames,
                        int32_t slots, int32_t nthreads, int32_t timeout_ms, int32_t flags, int64_t *written) {
    if (slots < 1 || frames < 0) return fail("bad arguments");
    RecRing r;
    long long fb = 0;
    for (int q = 0; q < ncomp && q < 3; ++q) fb += bytes ? bytes[q] : 0;
    if (fb <= 0 || fb % 8) return fail("bad frame size");
    r.frame_doubles = fb / 8;
    r.slots = slots;
    r.timeout_ms.store(timeout_ms > 0 ? timeout_ms : 1000);
    std::vector<double> host((size_t)(r.frame_doubles * slots));
    r.host = host.data();
    r.slot_tt.assign(slots, -1);
    NativeWriter w;
    OK(writer_setup(w, r, fd, ncomp, base, bytes, stride, frames));
    if ((flags & PHB_WRITER_MMAP) && !w.map_extents((flags & PHB_WRITER_POPULATE) != 0)) return fail("mmap of the frame extents failed: %s", strerror(errno));
    w.start(&r, nthreads > 0 ? nthreads : 2);
    int rc = 0;
    for (long long f = 0; f < frames && !rc; ++f) {
        const int wf = r.wait_free();
        if (wf == 1) rc = fail("recording aborted: %s", r.why().c_str());
        else if (wf) rc = fail("recorder ring full for %d ms", r.timeout_ms.load());
        else {
            double *slot = r.host + (f % slots) * r.frame_doubles;
            for (long long q = 0; q < r.frame_doubles; ++q) slot[q] = (double)f * 1e6 + (double)q;
            r.slot_tt[(size_t)(f % slots)] = f;
            r.produced.fetch_add(1);
        }
    }
    const bool ok = w.finish(timeout_ms > 0 ? timeout_ms : 1000);
    if (written) *written = w.written.load();
    if (rc) return rc;
    if (!ok) return fail("writer threads did not finish");
    if (r.aborted.load()) return fail("%s", r.why().c_str());
    return 0;
}

int phb_probe_add(phb_ctx *c, int32_t comp, int32_t j, int32_t k, int64_t capacity, int32_t *id) {
    ENTER(c);
    if (!id) return fail("null argument");
    if (comp < 0 || comp > 2) return fail("bad component %d", comp);
    const int nyc = c->cfg.ny - (comp == 1), nzc = c->cfg.nz - (comp == 2), nxc = c->cfg.nx - (comp == 0);
    if (j < 0 || j >= nyc || k < 0 || k >= nzc) return fail("probe line (%d, %d) outside the %d x %d field", j, k, nyc, nzc);
    if (capacity < 1) return fail("bad probe capacity");
    phb_ctx::Probe pr{};
    pr.comp = comp; pr.j = j; pr.k = k;
    pr.rows = std::max(0, std::min(c->cfg.x0 + c->cfg.nxl, nxc) - c->cfg.x0);
    pr.cap = capacity; pr.frames = 0;
    if (dmalloc(c, (void **)&pr.trace, (size_t)std::max(1LL, (long long)pr.rows * capacity) * sizeof(double), false)) return 1;
    c->probes.push_back(pr);
    *id = (int32_t)c->probes.size() - 1;
    return 0;
}
#define PROBE(c, id) \
    if ((id) < 0 || (size_t)(id) >= (c)->probes.size()) return fail("no probe %d", (int)(id)); \
    const phb_ctx::Probe &pr = (c)->probes[(size_t)(id)]
int phb_probe_shape(phb_ctx *c, int32_t id, int64_t *rows, int64_t *frames) {
    ENTER(c);
    PROBE(c, id);
    if (rows) *rows = pr.rows;
    if (frames) *frames = pr.frames;
    return 0;
}
int phb_probe_read(phb_ctx *c, int32_t id, double *out) {
    ENTER(c);
    PROBE(c, id);
    if (!out) return fail("null argument");
    const long long n = pr.frames, rows = pr.rows;
    if (n == 0 || rows == 0) return 0;
    std::vector<double> h((size_t)(n * rows));
    CU(cudaMemcpyAsync(h.data(), pr.trace, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    for (long long t = 0; t < n; ++t)
        for (long long p = 0; p < rows; ++p) out[p * n + t] = h[(size_t)(t * rows + p)];
    return 0;
}
int phb_probe_dft_t(phb_ctx *c, int32_t id, const double *window, int64_t nf, int64_t row0, int64_t nrows,
                    double *out_re, double *out_im) {
    ENTER(c);
    PROBE(c, id);
    if (!window || !out_re || !out_im) return fail("null argument");
    double2 *A = nullptr;
    OK(probe_dft_t_dev(c, pr, window, nf, row0, nrows, &A));
    int r = split_to_host(c, A, nrows * nf, out_re, out_im);
    cudaFreeAsync(A, c->st);
    return r;
}
int phb_probe_dft_xt(phb_ctx *c, int32_t id, const double *window, int64_t nf, int64_t nx_total,
                     double *out_re, double *out_im) {
    ENTER(c);
    PROBE(c, id);
    if (!window || !out_re || !out_im) return fail("null argument");
    if (nx_total < 1 || nx_total >= (1LL << 30) || c->cfg.x0 + pr.rows > nx_total)
        return fail("bad total row count %lld", (long long)nx_total);
    double2 *A = nullptr, *F = nullptr, *twx = nullptr;
    OK(probe_dft_t_dev(c, pr, window, nf, 0, pr.rows, &A));
    CU(cudaMallocAsync((void **)&F, (size_t)(nx_total * nf) * sizeof(double2), c->st));
    CU(cudaMallocAsync((void **)&twx, (size_t)nx_total * sizeof(double2), c->st));
    k_twiddle<<<(unsigned)((nx_total + 255) / 256), 256, 0, c->st>>>(twx, (int)nx_total);
    dim3 bl(32, 8), gr((unsigned)((nf + 31) / 32), (unsigned)((nx_total + 7) / 8));
    k_dft_x<<<gr, bl, 0, c->st>>>(A, pr.rows, (int)nf, c->cfg.x0, (int)nx_total, twx, F);
    c->launches += 2;
    CU(cudaGetLastError());
    int r = split_to_host(c, F, nx_total * nf, out_re, out_im);
    cudaFreeAsync(A, c->st); cudaFreeAsync(F, c->st); cudaFreeAsync(twx, c->st);
    return r;
}

}  // extern "C"

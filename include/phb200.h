/* phb200.h -- C ABI of libphb200.so, the B200 (sm_100a) FDTD time-stepping engine that
 * replaces the NumPy hot path of Phonomena's solver plugin.
 *
 * The reference has no FFI: its path is Python/NumPy inside BaseSolver
 * (phonomena/simulation/base_solver.py).  Each entry point below names the reference
 * code it replaces ("replaces:").  The Python plugin `phonomena_b200/solver_b200.py`
 * (class Solver, same interface as base_solver.BaseSolver) is the only intended caller
 * and binds these with ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - All functions return 0 on success, non-zero on error; phb_last_error() gives the
 *    message for the calling thread.  There is NO CPU fallback: without a CUDA device
 *    phb_create fails.
 *  - Host arrays are float64 (the reference's DTYPE, grid.py:21) in the reference's
 *    C-order shapes: ux (Nx-1,Ny,Nz), uy (Nx,Ny-1,Nz), uz (Nx,Ny,Nz-1)  (grid.py:88-102).
 *    The library owns its device memory and copies during the call; the caller may
 *    free its arrays as soon as the call returns.
 *  - A context may own the whole grid or one x-slab [x0, x0+nxl) of it (one process per
 *    GPU).  Slab-local host arrays hold only the owned planes: for ux the planes
 *    x0 .. min(x0+nxl, Nx-1)-1, for uy/uz the planes x0 .. x0+nxl-1.
 *  - Thread-safety: a context may be used from any host thread, one call at a time
 *    (the GUI calls init() and run() from different threads, main_widget.py:119-131).
 */
#ifndef PHB200_H
#define PHB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHB_VERSION 100

typedef struct phb_ctx phb_ctx;

enum { PHB_F32 = 0, PHB_F64 = 1 };
/* arithmetic mode.
 *  PHB_FAST : reciprocal spacings, FMA contraction allowed (fp64: <= 1e-12 rel-L2 of the reference)
 *  PHB_EXACT: true IEEE division, the reference's expression order, no contraction:
 *             fp64 results are BIT-IDENTICAL to the reference's NumPy evaluation. */
enum { PHB_FAST = 0, PHB_EXACT = 1, PHB_COMP = 2 };
/*  PHB_COMP : FAST arithmetic on the compensated state (u, delta = u - u_old) instead of (u, u_old):
 *             delta += dt^2/rho * div T;  u_new = u + delta.  Same step algebraically, several times less
 *             fp32 round-off drift on long runs, at 12 instead of 9 field words of traffic per cell.
 *             Measured on data/default.json (tests/test_gpu_long.py): rel-L2 vs the reference after 10^3 /
 *             10^4 steps: plain fp32 2.9e-6 / 5.9e-5, PHB_COMP 3.4e-7 / 1.45e-5 -- i.e. fp32 meets the 1e-5
 *             north-star tolerance up to a few thousand steps (plain) or ~7000 steps (PHB_COMP), NOT over
 *             10^4; use fp64 (1e-13 after 10^4 steps) when 1e-5 must hold on longer runs. */
/* stencil kernel selection: MARCH = the TMA-fed x-marching kernel (k_march.cuh), NAIVE = one thread per cell
 * (k_naive.cuh, the on-device specification); AUTO = MARCH except on slabs below 200 000 cells (too few tiles
 * for the serial x-march) and when the stencil-class table does not fit beside the shared-memory rings */
enum { PHB_KERNEL_AUTO = 0, PHB_KERNEL_NAIVE = 1, PHB_KERNEL_MARCH = 2 };
/* which displacement buffer */
enum { PHB_CUR = 0, PHB_OLD = 1 };
/* surface components to record (bit mask) */
enum { PHB_REC_UX = 1, PHB_REC_UY = 2, PHB_REC_UZ = 4,
       PHB_REC_FULL = 8 };   /* frames hold the whole arrays (reference shapes, z fastest) instead of their k = 0 planes */

/* y boundaries.  ABSORBING: first-order Mur faces at y = 0 and y = -1 (base_solver.py:543-550), what the reference
 * runs.  PERIODIC: the reference's archived stubs apply_T_pbc / apply_u_pbc (base_solver.py:383-400,
 * 475-486; zero Bloch phase), each applied after the corresponding traction-free update, and no Mur face in y; the x and
 * z faces stay.  Not available with PHB_COMP; slab contexts exchange halos through NCCL in this mode. */
enum { PHB_BC_ABSORBING = 0, PHB_BC_PERIODIC = 1 };

typedef struct phb_cfg {
    int32_t nx, ny, nz;      /* global grid points: grid.x.size, grid.y.size, grid.z.size        */
    int32_t x0, nxl;         /* owned x-planes [x0, x0+nxl); whole grid: x0 = 0, nxl = nx         */
    int32_t dtype;           /* PHB_F32 | PHB_F64: storage and arithmetic type on the device      */
    int32_t arith;           /* PHB_FAST | PHB_EXACT                                              */
    int32_t device;          /* CUDA device ordinal                                               */
    int32_t kernel;          /* PHB_KERNEL_*                                                      */
    int32_t record_mask;     /* PHB_REC_* bits; 0 = no surface recording                          */
    int32_t record_every;    /* record the k=0 plane after every n-th step (>=1)                  */
    int32_t ring_slots;      /* pinned-host ring depth (frames); 0 = default                      */
    int32_t bc_y;            /* PHB_BC_ABSORBING (0, the reference's behaviour) | PHB_BC_PERIODIC                */
    int32_t record_stride[3]; /* PHB_REC_FULL only: keep every sx-th plane / sy-th row / sz-th level (global indices that
                                * are multiples of the stride; 0 = 1).  A decimated full-volume snapshot of a large grid
                                * fits the recorder ring where the whole arrays would not; a slab needs x0 % sx == 0. */
    double  dt;              /* material.dt                                      (material.py:80-93) */
    double  d2;              /* dt**2 as evaluated by the host (base_solver.py:443: self.m.dt**2) */
} phb_cfg;

int         phb_version(void);
const char *phb_last_error(void);
int         phb_device_count(int *n);

/* replaces: BaseSolver.init's field allocation via Grid.update (grid.py:79-110): all
 * displacement buffers start at zero. */
int phb_create(const phb_cfg *cfg, phb_ctx **out);
int phb_destroy(phb_ctx *ctx);

/* replaces: use of grid.fdx/fdy/fdz/sdx/sdy/sdz (grid.py:118-127) in every update.
 * Global arrays of lengths nx-1, ny-1, nz-1, nx-2, ny-2, nz-2. */
int phb_set_spacing(phb_ctx *ctx, const double *fdx, const double *fdy, const double *fdz,
                    const double *sdx, const double *sdy, const double *sdz);

/* replaces: Material.C / Material.P as a table: nmat materials, each the 12 stiffness
 * entries the path reads -- C[r][c] for r,c in 0..2 (row-major, NOT symmetrised), then
 * C[3][3], C[4][4], C[5][5] -- and the density (material.py:55-63). */
int phb_set_material_table(phb_ctx *ctx, int32_t nmat, const double *c12, const double *rho);

/* replaces: Material.C / Material.P handed over as the reference stores them (material.py:48-63;
 * SURVEY 8a2): C (planes, Ny, Nz, 6, 6) and P (planes, Ny, Nz) float64 for the same planes as
 * phb_set_material_ids.  The distinct (12 entries read + density) tuples become the material table
 * and the id map in one call; more distinct cells than the table holds (15) is an error. */
int phb_set_material_dense(phb_ctx *ctx, const double *C, const double *P, int64_t nplanes);

/* per-cell material id (index into the table), reference layout (planes, Ny, Nz) uint8,
 * for planes x0 .. min(x0+nxl+1, nx)-1  (one extra plane to the right when it exists). */
int phb_set_material_ids(phb_ctx *ctx, const uint8_t *ids, int64_t nplanes);

/* replaces: Grid.inclusionIndices + Material.setConstants (grid.py:158-175,
 * material.py:55-63) evaluated on the device with the same float32/float64 mixing and the
 * same z re-binding quirk: id 0 everywhere, id 1 inside the cylinders.
 * targets: n x 4 float32 (x, y, z, r); x, y, z: global mesh lines. */
int phb_gen_material_ids(phb_ctx *ctx, const float *targets, int32_t n,
                         const double *x, const double *y, const double *z);
/* owned planes x0..x0+nxl-1, layout (nxl, Ny, Nz) */
int phb_get_material_ids(phb_ctx *ctx, uint8_t *ids);

/* replaces: coefficient evaluation at the top of apply_u_abc (base_solver.py:525-537).
 * coef = { clx, ctx, cly0, cty0, cly1, cty1, clz, ctz } computed by the host exactly as there. */
int phb_set_abc(phb_ctx *ctx, const double coef[8]);

/* replaces: wave_fn(tt) (base_solver.py:251,294-312): the source samples of the NEXT n steps
 * (w[0] belongs to the step phb_steps_done() reports now), evaluated by the host with the
 * reference's own expression.  May be called between phb_run calls to stream the samples: the copy is ordered
 * on the library's launch stream behind the steps already enqueued (no synchronisation) and the caller's
 * array may be reused as soon as the call returns. */
int phb_set_source_table(phb_ctx *ctx, const double *w, int64_t n);

/* field transfer (tests, restart, full-field output); which = PHB_CUR | PHB_OLD.
 * Any pointer may be NULL to skip that component. */
int phb_set_fields(phb_ctx *ctx, int32_t which, const double *ux, const double *uy, const double *uz);
int phb_get_fields(phb_ctx *ctx, int32_t which, double *ux, double *uy, double *uz);

/* debug/parity: the six stress arrays the reference would hold for the displacement
 * buffer `which` (T1..T3 (n,Ny,Nz), T4 (n,Ny-1,Nz-1), T5 (n',Ny,Nz-1), T6 (n',Ny-1,Nz);
 * owned planes).  After k steps, which = PHB_OLD reproduces grid.T1..T6. */
int phb_get_stress(phb_ctx *ctx, int32_t which, double *T1, double *T2, double *T3,
                   double *T4, double *T5, double *T6);

/* replaces: the body of the time loop (base_solver.py:251-256) for nsteps steps:
 * source, update_T, update_T_BC, update_u, update_u_BC, time_step [, halo exchange, record].
 * Enqueues on the context's stream and returns; phb_sync waits. */
int phb_run(phb_ctx *ctx, int64_t nsteps);
int phb_sync(phb_ctx *ctx);
/* same, bracketed by CUDA events on the launching stream; returns elapsed ms */
int phb_run_timed(phb_ctx *ctx, int64_t nsteps, float *ms);
int phb_steps_done(phb_ctx *ctx, int64_t *tt);
/* number of kernels launched by this context so far */
int phb_launch_count(phb_ctx *ctx, int64_t *n);
/* name of the stencil kernel variant in use, and device bytes allocated */
int phb_info(phb_ctx *ctx, char *kernel_name, int32_t len, int64_t *device_bytes);

/* Bloch-periodic y boundaries with a phase (beyond the reference, whose archived stubs are the
 * phase-0 case): u(y + L) = u(y) e^{i phase}, L = ny - 2 rows.  The field is complex = two contexts created with the same
 * configuration and bc_y = PHB_BC_PERIODIC, `re` and `im`; after pairing, phb_run on `re` steps both (the real stencil
 * advances each part on its own, the periodic copies mix them with cos / sin of the phase), `im` refuses phb_run.
 * Set up both parts (spacing, material, abc) before the first step; the source drives the real part only (give `im`
 * no source table); fields are read from each context with phb_get_fields.  One GPU, not with PHB_COMP.  Destroying
 * either context dissolves the pair. */
int phb_bloch_pair(phb_ctx *re, phb_ctx *im, double phase);

/* per-kernel timing of the fused stencil launches (the dominant kernel), for the roofline in
 * bench.py: CUDA event pairs on the launching stream around every stencil launch.
 * Returns the accumulated milliseconds and launch count since profiling was last (re)enabled;
 * enable = 1 / 0 switches it on / off and resets the accumulators, enable < 0 only reads. */
int phb_profile(phb_ctx *ctx, int32_t enable, double *kernel_ms, int64_t *kernel_launches);

/* multi-GPU (one process per GPU): NCCL communicator over the slab chain.
 * replaces: nothing (the reference has no domain decomposition). */
int phb_comm_unique_id(char id[128]);
int phb_comm_init(phb_ctx *ctx, const char id[128], int32_t rank, int32_t nranks);

/* Fused halo exchange over NVLink peer memory (preferred over the NCCL path): every rank exports CUDA-IPC
 * handles of its three displacement buffers and of its flag word (4 x 64 bytes) plus its nxl; the host
 * hands each rank the exports of its left and right neighbours (NULL at the ends).  The stencil kernel then
 * stores the slab's first / last plane straight into the neighbours' ghost planes; a stream-ordered flag
 * write / wait replaces the collective; each rank applies the y / z absorbing faces to its ghost planes.
 * replaces: nothing in the reference.
 * Two ways to move the planes (phb_p2p_mode; default 0): 0 = after the faces of a step one small kernel copies the two
 * finished edge planes into the neighbours' ghost planes (16-byte peer stores) and publishes the step number in their
 * flag words; 1 = "fused": the stencil kernel itself stores the edge planes while it computes them, the receiver
 * applies the y / z faces to its ghost planes.  Both are bit-identical to a single-GPU run; mode 0 is faster here
 * (the transfer is 7 us of NVLink time per step, the in-kernel stores cost the memory-bound stencil 38 us). */
int phb_p2p_export(phb_ctx *ctx, char handles[256], int32_t *nxl);
int phb_p2p_mode(phb_ctx *ctx, int32_t fused_in_kernel);
int phb_p2p_import(phb_ctx *ctx, int32_t rank, int32_t nranks, const char *left, int32_t left_nxl,
                   const char *right, int32_t right_nxl);

/* surface recording: replaces Grid.freezeData + Writer.put (grid.py:68-77,
 * base_solver.py:97-100,258-260) for the k=0 plane.  Frames are written by the device
 * into a pinned host ring; the consumer takes them in order.
 * phb_record_next: blocks until the next frame is complete (or timeout_ms elapses ->
 * returns 2 with *tt = -1); on success *frame points at ring memory holding, for each
 * recorded component in the order ux, uy, uz, a (planes, Ny') float64 array, and *tt is
 * the step index the frame belongs to.  The slot stays valid until phb_record_release. */
int phb_record_next(phb_ctx *ctx, const double **frame, int64_t *tt, int32_t timeout_ms);
int phb_record_release(phb_ctx *ctx);
int phb_record_frame_doubles(phb_ctx *ctx, int64_t *n);
/* Both sides of the ring are bounded, like the reference's writer queue (queue.get(timeout=120) and
 * join(300), base_solver.py:89-92,148,274): phb_run waits at most `timeout_ms` (default 120 000, or
 * $PHB_REC_TIMEOUT_MS) for a free slot and then fails; a consumer that cannot go on (e.g. ENOSPC) calls
 * phb_record_abort, which makes the waiting / next phb_run fail with that message instead of hanging.
 * Neither call takes the context lock: they are meant to be called while phb_run is blocked. */
int phb_record_abort(phb_ctx *ctx, const char *why);
int phb_record_timeout(phb_ctx *ctx, int32_t timeout_ms);

/* replaces: BaseSolver.cancel (base_solver.py:246-248,282-284: a threading.Event polled once per step).
 * Callable from any thread while phb_run is executing: phb_run returns 3 ("cancelled") after the step in
 * flight -- also when it is waiting for a free recorder slot.  Does not take the context lock. */
int phb_cancel(phb_ctx *ctx);

/* replaces: Writer.run (base_solver.py:135-160), the consumer thread / process that stores one time slab
 * per queue item.  `nthreads` native threads drain the pinned ring: frame f (the f-th recorded frame) has
 * its component c written with pwrite(fd, ...) from the pinned slot to file offset base[c] + f * stride
 * (bytes[c] bytes; components in the ring's order ux, uy, uz as selected by record_mask) -- the chunk
 * addresses the host-side HDF5 writer reserved before the run.  `fd` stays owned by the caller and must
 * stay open until phb_writer_finish has returned.  A write error aborts the recording (phb_run then fails
 * with the errno text).  phb_writer_finish: no further frame will be produced; drain, join (at most
 * timeout_ms; 0 = 300 s), report the number of frames written and the seconds the threads spent waiting
 * for frames / writing them.  phb_destroy joins a writer that is still running.
 * flags: PHB_WRITER_MMAP -- the caller has ALLOCATED every frame extent (posix_fallocate) and opened `fd` read-write:
 * the extents are mapped MAP_SHARED and the threads copy into the page cache in parallel (pwrite on one file
 * serialises on the inode lock); PHB_WRITER_POPULATE pre-faults the mapping.  If the mapping is refused the threads
 * fall back to pwrite (phb_writer_mapped tells which). */
enum { PHB_WRITER_MMAP = 1, PHB_WRITER_POPULATE = 2 };
int phb_writer_start(phb_ctx *ctx, int32_t fd, int32_t ncomp, const int64_t *base, const int64_t *bytes,
                     int64_t stride, int64_t frames, int32_t nthreads, int32_t flags);
int phb_writer_mapped(phb_ctx *ctx, int32_t *mapped);
int phb_writer_finish(phb_ctx *ctx, int32_t timeout_ms, int64_t *written, double *wait_s, double *write_s);
/* host-only exercise of the ring + writer threads (CPU tests; makes no CUDA call): a producer thread
 * pushes `frames` synthetic frames (element q of frame f = f * 1e6 + q) through a `slots`-deep ring. */
int phb_writer_selftest(int32_t fd, int32_t ncomp, const int64_t *base, const int64_t *bytes, int64_t stride,
                        int64_t frames, int32_t slots, int32_t nthreads, int32_t timeout_ms, int32_t flags, int64_t *written);

/* line probes and on-device spectra.
 * replaces: simulation/analysis.py:59-66 -- the reference re-reads the (x, t) matrix
 * u[:, y, z, :] from the HDF5 file; a probe keeps it in device memory while the run goes on.
 * phb_probe_add: probe component comp (0 ux, 1 uy, 2 uz) on the line (:, j, k), room for
 * `capacity` samples.  One sample is taken after every step tt with tt % record_every == 0 --
 * the same instants as the recorded frames (base_solver.py:256-260: the state AFTER time_step).
 * Rows are this slab's planes on which the component exists (ux has Nx-1 planes). */
int phb_probe_add(phb_ctx *ctx, int32_t comp, int32_t j, int32_t k, int64_t capacity, int32_t *id);
int phb_probe_shape(phb_ctx *ctx, int32_t id, int64_t *rows, int64_t *frames);
/* out[row][frame], float64 */
int phb_probe_read(phb_ctx *ctx, int32_t id, double *out);
/* replaces: simulation/analysis.py:67-86 (np.fft.fft / np.fft.fft2 of the windowed matrix).
 * window: the `frames` weights (np.hanning(frames) in the reference); nf: frequencies kept
 * (frames // 2).  Outputs are UNNORMALISED complex sums; the caller applies norm="ortho" and abs.
 * phb_probe_dft_t : out[(row - row0)][kt] = sum_t u[row][t] w[t] exp(-2 pi i kt t / frames),
 *                   rows row0 .. row0 + nrows - 1 of this slab.
 * phb_probe_dft_xt: out[kx][kt] = sum_{rows of this slab} (the above)[row][kt] *
 *                   exp(-2 pi i kx (x0 + row) / nx_total): this slab's share of the 2-D
 *                   transform over nx_total rows; slabs add. */
int phb_probe_dft_t(phb_ctx *ctx, int32_t id, const double *window, int64_t nf, int64_t row0, int64_t nrows,
                    double *out_re, double *out_im);
int phb_probe_dft_xt(phb_ctx *ctx, int32_t id, const double *window, int64_t nf, int64_t nx_total,
                     double *out_re, double *out_im);

#ifdef __cplusplus
}
#endif
#endif /* PHB200_H */

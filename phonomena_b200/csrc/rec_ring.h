// rec_ring.h -- host side of the recorder: the pinned frame ring the device fills and the native
// writer threads that drain it into the output file.
//
// replaces: base_solver.Writer.put / Writer.run (base_solver.py:97-100,135-160): the reference pushes
// (FrozenGrid, tt) through a queue to a writer thread / process that stores one time slab per item with
// h5py.  Here the producer is the stepping loop (record_frame in phb200.cu: gather kernel -> device
// staging slot -> async D2H into a pinned slot), the consumers are C++ threads that pwrite() each
// frame's components straight from the pinned slot to their final offsets in the file (the chunk
// addresses are fixed up front by the host-side HDF5 writer, h5lite.reserve_frames) -- no GIL, no
// intermediate copy.  Both sides are bounded like the reference's (queue.get(timeout=120), join(300),
// base_solver.py:89-92,148,274): a full ring makes the producer wait at most `timeout_ms`, and either
// side can abort the other (a writer that hits ENOSPC makes phb_run fail instead of hanging it).
//
// Host-only C++ (no CUDA types except the optional per-slot events), so the same code is exercised
// on a CPU-only machine through phb_writer_selftest.
#pragma once
#include <cuda_runtime.h>
#include <errno.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace phb {

struct RecRing {
    double *host = nullptr;            // slots x frame_doubles (pinned when a device feeds it)
    long long frame_doubles = 0;
    int slots = 0;
    int device = -1;                   // >= 0: consumers wait on slot_ev before reading a slot
    std::vector<cudaEvent_t> slot_ev;  // D2H copy of the slot complete
    std::vector<long long> slot_tt;    // step index of the frame in the slot
    std::vector<char> slot_done;       // native writer: slot written, waiting for in-order release
    std::atomic<long long> produced{0}, consumed{0}, released{0};
    std::atomic<int> aborted{0};
    std::atomic<int> in_api{0};        // consumer threads currently inside phb_record_next / phb_record_release (phb_destroy waits for 0)
    std::atomic<int> timeout_ms{120000};   // producer: longest wait for a free slot (reference: queue timeout 120 s)
    std::mutex mu;                     // abort message, in-order release
    std::string abort_msg;

    void abort(const char *why) {
        std::lock_guard<std::mutex> lk(mu);
        if (!aborted.load()) abort_msg = why ? why : "aborted";
        aborted.store(1);
    }
    std::string why() {
        std::lock_guard<std::mutex> lk(mu);
        return abort_msg;
    }
    // producer: wait for a free slot.  0 = got one, 1 = aborted, 2 = timed out, 3 = cancelled
    int wait_free(const std::atomic<int> *cancel = nullptr) {
        const auto t0 = std::chrono::steady_clock::now();
        int spins = 0;
        while (produced.load() - released.load() >= slots) {
            if (aborted.load()) return 1;
            if (cancel && cancel->load()) return 3;
            if ((++spins & 63) == 0 &&
                std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() >= timeout_ms.load())
                return 2;
            std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
        return aborted.load() ? 1 : 0;
    }
    // consumer side of the in-order release: frame f (slot f % slots) has been consumed
    void release_in_order(long long f) {
        std::lock_guard<std::mutex> lk(mu);
        slot_done[(size_t)(f % slots)] = 1;
        long long r = released.load();
        while (r < produced.load() && slot_done[(size_t)(r % slots)]) {
            slot_done[(size_t)(r % slots)] = 0;
            ++r;
        }
        released.store(r);
    }
};

// The native writer: `nthreads` threads; each claims the next frame index, waits until the device has
// delivered it, writes every recorded component to base[c] + frame * stride, releases the slot in order.
struct NativeWriter {
    RecRing *ring = nullptr;
    int fd = -1;
    int ncomp = 0;
    long long base[3] = {0, 0, 0}, bytes[3] = {0, 0, 0}, off[3] = {0, 0, 0};   // file offset of frame 0, bytes, byte offset inside a ring frame
    long long stride = 0;              // file distance between consecutive frames of one component
    long long frames = 0;              // frames the file has room for
    std::vector<std::thread> th;
    std::atomic<long long> claim{0}, written{0};
    std::atomic<int> stop{0};          // no frame beyond `produced` will come: drain and exit
    std::atomic<long long> wait_us{0}, write_us{0};
    bool started = false;
    // Optional: the frame extents mapped MAP_SHARED (the caller has ALLOCATED them, posix_fallocate): the threads then
    // memcpy into the page cache in parallel.  pwrite() on one file serialises on the inode lock (tmpfs, ext4, xfs ...),
    // so four threads wrote one 6 MB frame per 1.07 ms between them; mapped, they do not share a lock.
    char *map = nullptr;
    long long map_off = 0, map_len = 0;

    bool map_extents(bool populate) {
        long long lo = base[0], hi = 0;
        for (int c = 0; c < ncomp; ++c) {
            lo = std::min(lo, base[c]);
            hi = std::max(hi, base[c] + (frames - 1) * stride + bytes[c]);
        }
        if (frames <= 0 || hi <= lo) return false;
        const long long pg = sysconf(_SC_PAGESIZE);
        map_off = lo / pg * pg;
        map_len = hi - map_off;
        void *p = mmap(nullptr, (size_t)map_len, PROT_READ | PROT_WRITE, MAP_SHARED | (populate ? MAP_POPULATE : 0), fd, (off_t)map_off);
        if (p == MAP_FAILED) { map = nullptr; return false; }
        map = static_cast<char *>(p);
        return true;
    }
    // Tearing down a populated multi-GB mapping takes tens of milliseconds (page-table walk); nothing waits for it: the
    // data is in the page cache, the file can be closed while the mapping still exists -> a detached thread unmaps.
    void unmap() {
        if (map) {
            char *m = map;
            const size_t len = (size_t)map_len;
            std::thread([m, len] { munmap(m, len); }).detach();
        }
        map = nullptr;
    }

    static long long now_us() {
        return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }

    void worker() {
        RecRing &r = *ring;
        if (r.device >= 0) cudaSetDevice(r.device);
        for (;;) {
            const long long f = claim.fetch_add(1);
            if (f >= frames) return;
            long long t0 = now_us();
            while (r.produced.load() <= f) {
                if (r.aborted.load()) return;
                if (stop.load() && r.produced.load() <= f) return;
                std::this_thread::sleep_for(std::chrono::microseconds(50));
            }
            const int s = (int)(f % r.slots);
            if (r.device >= 0 && !r.slot_ev.empty()) {
                const cudaError_t e = cudaEventSynchronize(r.slot_ev[(size_t)s]);
                if (e != cudaSuccess) {
                    r.abort((std::string("recorder: device copy of a frame failed: ") + cudaGetErrorString(e)).c_str());
                    return;
                }
            }
            long long t1 = now_us();
            wait_us += t1 - t0;
            const char *src = reinterpret_cast<const char *>(r.host + (long long)s * r.frame_doubles);
            for (int c = 0; c < ncomp; ++c) {
                long long left = bytes[c], pos = base[c] + f * stride;
                const char *p = src + off[c];
                if (map) {
                    memcpy(map + (pos - map_off), p, (size_t)left);
                    continue;
                }
                while (left > 0) {
                    const ssize_t n = pwrite(fd, p, (size_t)left, (off_t)pos);
                    if (n < 0) {
                        if (errno == EINTR) continue;
                        char msg[256];
                        snprintf(msg, sizeof msg, "writer: pwrite of frame %lld failed: %s", f, strerror(errno));
                        r.abort(msg);
                        return;
                    }
                    if (n == 0) { r.abort("writer: pwrite wrote nothing"); return; }
                    left -= n; pos += n; p += n;
                }
            }
            write_us += now_us() - t1;
            written.fetch_add(1);
            r.release_in_order(f);
        }
    }

    int start(RecRing *r, int nthreads) {
        ring = r;
        claim.store(0); written.store(0); stop.store(0);
        r->slot_done.assign((size_t)r->slots, 0);
        if (nthreads < 1) nthreads = 1;
        if (nthreads > r->slots) nthreads = r->slots;
        started = true;
        for (int q = 0; q < nthreads; ++q) th.emplace_back([this] { worker(); });
        return 0;
    }
    // Ask the threads to drain what has been produced and wait for them.  false: still running after timeout_ms
    // (the ring is then aborted so that they stop at the next poll, and they are joined anyway).
    bool finish(long long timeout_ms) {
        if (!started) return true;
        stop.store(1);
        const auto t0 = std::chrono::steady_clock::now();
        bool ok = true;
        // the workers exit once every produced frame is claimed; a stuck pwrite is the only way to exceed the timeout
        while (written.load() < std::min(ring->produced.load(), frames) && !ring->aborted.load()) {
            if (std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() >= timeout_ms) {
                ring->abort("writer threads did not finish in time");
                ok = false;
                break;
            }
            std::this_thread::sleep_for(std::chrono::microseconds(100));
        }
        for (auto &t : th) if (t.joinable()) t.join();
        th.clear();
        unmap();
        started = false;
        return ok;
    }
};

}  // namespace phb

// k_probe.cuh -- line probes and their Hann-windowed spectra on the device.
//
// The reference's post-processing (simulation/analysis.py:44-96) re-reads u[:, y, z, :] from the
// HDF5 file -- an (x, t) matrix -- multiplies by np.hanning(N) along t and takes the magnitude of
// a 1-D (t) or 2-D (x, t) DFT with norm="ortho", keeping the first N // 2 frequencies.  A probe
// keeps that (x, t) matrix in HBM while the simulation runs (one tiny gather per recorded step),
// so the spectrum needs neither the file nor a host transfer of the frames.
//
// The transform lengths are the step count and the grid extent -- arbitrary integers, a few
// hundred to a few thousand -- and the whole job is ~1e8..1e9 multiply-adds, so it is evaluated
// as a direct DFT in fp64 with an exact integer phase index (k * n mod N into a twiddle table);
// no radix restrictions, error ~ sqrt(N) ulp.
#pragma once
#include "fd_common.cuh"

namespace phb {

// trace[f * rows + p] = u_comp(x0 + p, j, k)
template <class T>
__global__ void k_probe_sample(Geo<T> g, const T *u, int j, int k, int rows, double *dst) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < rows) dst[p] = (double)u[g.idx(g.x0 + p, j, k)];
}

// tw[m] = exp(-2 pi i m / n), m = 0 .. n-1
__global__ void k_twiddle(double2 *tw, int n) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    double s, c;
    sincospi(2.0 * (double)m / (double)n, &s, &c);
    tw[m] = make_double2(c, -s);
}

// A[p][kt] = sum_t trace[t][p] * win[t] * tw[(kt * t) mod n]      p < rows, kt < nf
// threads: x = row p (coalesced trace reads), y = frequency kt
__global__ void k_dft_t(const double *trace, int rows, int n, const double *win, const double2 *tw, int nf,
                        int row0, int nrows, double2 *A) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;     // output row
    const int kt = blockIdx.y * blockDim.y + threadIdx.y;
    if (q >= nrows || kt >= nf) return;
    const int p = row0 + q;
    double re = 0.0, im = 0.0;
    int m = 0;                                               // (kt * t) mod n, exact
    for (int t = 0; t < n; ++t) {
        const double v = trace[(long long)t * rows + p] * win[t];
        const double2 w = tw[m];
        re = fma(v, w.x, re);
        im = fma(v, w.y, im);
        m += kt;
        if (m >= n) m -= n;
    }
    A[(long long)q * nf + kt] = make_double2(re, im);
}

// F[kx][kt] = sum_p A[p][kt] * twx[(kx * (x0 + p)) mod nxt]      kx < nxt, kt < nf
// (this slab's partial sum of the x transform; slabs add)
__global__ void k_dft_x(const double2 *A, int rows, int nf, int x0, int nxt, const double2 *twx, double2 *F) {
    const int kt = blockIdx.x * blockDim.x + threadIdx.x;    // coalesced along kt
    const int kx = blockIdx.y * blockDim.y + threadIdx.y;
    if (kt >= nf || kx >= nxt) return;
    double re = 0.0, im = 0.0;
    int m = (int)(((long long)kx * x0) % nxt);
    for (int p = 0; p < rows; ++p) {
        const double2 a = A[(long long)p * nf + kt];
        const double2 w = twx[m];
        re = fma(a.x, w.x, fma(-a.y, w.y, re));
        im = fma(a.x, w.y, fma(a.y, w.x, im));
        m += kx;
        if (m >= nxt) m -= nxt;
    }
    F[(long long)kx * nf + kt] = make_double2(re, im);
}

}  // namespace phb

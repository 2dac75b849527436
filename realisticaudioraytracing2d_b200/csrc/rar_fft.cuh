// rar_fft.cuh -- building blocks of the hand-written shared-memory FFT used by the partitioned
// overlap-save convolution (the B200 replacement for AudioConvolve.compute:13-31).
//
// A real FFT of size 2M (M = 256: block 256, window 512) is computed as one complex FFT of size M on
// the packed sequence z[n] = x[2n] + i x[2n+1], followed by a split step.  The complex FFT is a
// Stockham auto-sort radix-4 transform: log4(M) = 4 passes, M/4 = 64 butterflies per pass, ping-pong
// between two shared-memory buffers, no bit reversal.
//
// Spectra are stored as "packed half spectra": M complex values, bin 0 = (X[0], X[M]) (DC and Nyquist
// are both real), bins 1..M-1 = X[k].
//
// Every function takes the index of the butterfly ("thread") it performs, so the same code runs in a
// CUDA block (one call per thread, __syncthreads between passes) and in the host test harness
// (tests/host_emulation.cpp: a loop over the index), which checks it against numpy.fft without a GPU.
//
// The reference only has an unused single-thread radix-2 FFT for N=128
// (RaytraceOcclusion2D.compute:352-425, never dispatched), so there is nothing to match bit-for-bit
// here; the contract is the 1e-4 relative-L2 tolerance on the convolved audio.
#pragma once

#include "rar_ray.cuh"  // f2, RAR_HD

namespace rar {

constexpr int kFftM = 256;       // complex transform size
constexpr int kFftLogM = 8;
constexpr int kFftThreads = 64;  // butterflies per pass

RAR_HD f2 cmul(f2 a, f2 b) { return f2{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
RAR_HD f2 cadd(f2 a, f2 b) { return f2{a.x + b.x, a.y + b.y}; }
RAR_HD f2 csub(f2 a, f2 b) { return f2{a.x - b.x, a.y - b.y}; }
RAR_HD f2 conj2(f2 a) { return f2{a.x, -a.y}; }

// One radix-4 Stockham pass.  i in [0, M/4); p = 1, 4, 16, 64.  tw[k] = exp(-2*pi*i*k/M), k in [0, M).
// inverse: conjugated twiddles and butterflies (no scaling).
RAR_HD void fft_pass_r4(const f2 *in, f2 *out, int i, int p, const f2 *tw, bool inverse) {
    const int t = kFftM / 4;
    const int k = i & (p - 1);
    const int j = ((i - k) << 2) + k;
    f2 u0 = in[i], u1 = in[i + t], u2 = in[i + 2 * t], u3 = in[i + 3 * t];
    // twiddle m: exp(-2*pi*i * m*k / (4p)) = tw[m * k * (M / (4p))]
    const int step = k * (kFftM / (4 * p));
    f2 w1 = tw[step], w2 = tw[2 * step], w3 = tw[3 * step];
    if (inverse) { w1 = conj2(w1); w2 = conj2(w2); w3 = conj2(w3); }
    u1 = cmul(u1, w1);
    u2 = cmul(u2, w2);
    u3 = cmul(u3, w3);
    f2 a0 = cadd(u0, u2), a1 = csub(u0, u2), a2 = cadd(u1, u3), a3 = csub(u1, u3);
    // forward: multiply a3 by -i; inverse: by +i
    f2 b3 = inverse ? f2{-a3.y, a3.x} : f2{a3.y, -a3.x};
    out[j] = cadd(a0, a2);
    out[j + p] = cadd(a1, b3);
    out[j + 2 * p] = csub(a0, a2);
    out[j + 3 * p] = csub(a1, b3);
}

// Split step after the forward complex FFT: Z (size M) -> packed half spectrum P (size M).
// k in [0, M/2]; handles bins k and M-k.  tw2[k] = exp(-2*pi*i*k/(2M)), k in [0, M/2].
RAR_HD void rfft_split(const f2 *Z, f2 *P, int k, const f2 *tw2) {
    if (k == 0) {
        P[0] = f2{Z[0].x + Z[0].y, Z[0].x - Z[0].y};
        return;
    }
    const f2 zk = Z[k], zm = Z[kFftM - k];
    // Xe = (zk + conj(zm))/2 ; Xo = (zk - conj(zm))/(2i)
    const f2 xe = f2{0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y)};
    const f2 xo = f2{0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x)};
    const f2 w = tw2[k];
    const f2 wxo = cmul(w, xo);
    P[k] = cadd(xe, wxo);
    // bin M-k: Xe[M-k] = conj(Xe[k]), Xo[M-k] = conj(Xo[k]), w[M-k] = -conj(w[k])
    P[kFftM - k] = f2{xe.x - wxo.x, -(xe.y - wxo.y)};
}

// Merge step before the inverse complex FFT: packed half spectrum P -> Z (size M), so that
// ifft_M(Z)[n] * (1/M) = x[2n] + i x[2n+1].
RAR_HD void irfft_merge(const f2 *P, f2 *Z, int k, const f2 *tw2) {
    if (k == 0) {
        const float x0 = P[0].x, xm = P[0].y;
        Z[0] = f2{0.5f * (x0 + xm), 0.5f * (x0 - xm)};
        return;
    }
    const f2 xk = P[k], xmk = P[kFftM - k];
    // Xe = (X[k] + conj(X[M-k]))/2 ; Xo = (X[k] - conj(X[M-k]))/2 * conj(w)
    const f2 xe = f2{0.5f * (xk.x + xmk.x), 0.5f * (xk.y - xmk.y)};
    const f2 d = f2{0.5f * (xk.x - xmk.x), 0.5f * (xk.y + xmk.y)};
    const f2 xo = cmul(d, conj2(tw2[k]));
    // Z[k] = Xe + i*Xo ; Z[M-k] = conj(Xe) + i*conj(Xo)
    Z[k] = f2{xe.x - xo.y, xe.y + xo.x};
    Z[kFftM - k] = f2{xe.x + xo.y, -xe.y + xo.x};
}

}  // namespace rar

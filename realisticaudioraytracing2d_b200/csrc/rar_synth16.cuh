// rar_synth16.cuh -- register-resident 256-point transforms for the filter-bank synthesis kernel (band_synth.cu).
//
// The first synthesis kernel (conv_kernels.cu, band_synth_kernel) ran every band's transform through the four-pass
// shared-memory FFT of rar_fft.cuh and was bound by shared-memory wavefronts (72 % of peak, 0.16 of the HBM roofline,
// profiles/r02_ncu_band_synth.txt).  Here a transform of M = 256 complex points is split 16 x 16 over 16 threads that
// hold 16 points each: a 16-point transform in registers, one twiddle multiplication, ONE transposition through
// shared memory, a second 16-point transform in registers.  With n = t + 16 r and k = k1 + 16 k2,
//        Z[k1 + 16 k2] = sum_t W16^(t k2) * ( W256^(t k1) * sum_r z[t + 16 r] W16^(r k1) ).
// Thread t computes the inner sum for its residue t; after the transposition thread k1 holds Z[k1 + 16 k2] for all k2.
//
// The synthesis needs S[k] = sum_b A_b[k] P_b[k], P_b the 512-point real spectrum of band b's segment, obtained from
// the packed transform Z_b by the split step P[k] = alpha_k Z[k] + beta_k conj(Z[M-k]).  Z[M-k] lives in another
// thread, so instead of splitting every band the kernel accumulates two sums with REAL weights on its own elements,
//        U[k] = sum_b A_b[k] Z_b[k],      W[k] = sum_b A_b[M-k] Z_b[k],
// and does the exchange with the partner thread once per segment: S[k] = alpha_k U[k] + beta_k conj(W[M-k]); composed
// with the merge step of the inverse transform this collapses to three real multiples per bin (synth_merge1).
// A_b is the zero-phase amplitude of band b's symmetric 255-tap filter (the filters are used centred; the 512-point
// circular convolution then wraps the 127 samples before the segment to the end of the window, which the output
// step unwraps), so the weights are real and a band costs 4 FMAs per element.
//
// Like rar_fft.cuh these are RAR_HD functions of the thread index so that tests/host_emulation.cpp can run them
// without a GPU.  Floating-point contract: 1e-4 relative L2 on the synthesised response (tests), not bit-exactness.
#pragma once

#include <math.h>
#include <stddef.h>

#include "rar_fft.cuh"

namespace rar {

constexpr int kSyn = 16;  // threads per transform = points per thread

// (a, b, c, d) <- 4-point DFT; INV: conjugated kernel, no scaling
template <bool INV>
RAR_HD void dft4(f2 &a, f2 &b, f2 &c, f2 &d) {
    const f2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = csub(b, d);
    const f2 r = INV ? f2{-s3.y, s3.x} : f2{s3.y, -s3.x};  // (-/+ i) s3
    a = cadd(s0, s2);
    b = cadd(s1, r);
    c = csub(s0, s2);
    d = csub(s1, r);
}

// x * W16^m, W16 = exp(-2 pi i / 16) (INV: its conjugate)
template <bool INV, int m>
RAR_HD f2 mul_w16(f2 x) {
    constexpr float c = 0.92387953251128674f, s = 0.38268343236508977f, r = 0.70710678118654752f;
    if (m == 0) return x;
    if (m == 4) return INV ? f2{-x.y, x.x} : f2{x.y, -x.x};
    const float wr = m == 1 ? c : m == 2 ? r : m == 3 ? s : m == 6 ? -r : -c;                   // m == 9: -c
    const float wi_f = m == 1 ? -s : m == 2 ? -r : m == 3 ? -c : m == 6 ? -r : s;              // forward imaginary part
    const float wi = INV ? -wi_f : wi_f;
    return f2{x.x * wr - x.y * wi, x.x * wi + x.y * wr};
}

// 16-point DFT, natural order in and out.  HALF: x[8..15] are taken as zero (and not read).
template <bool INV, bool HALF>
RAR_HD void fft16(f2 (&x)[16]) {
    // n = n1 + 4 n2, k = k2 + 4 k1:  a[n1][k2] = sum_n2 x[n1 + 4 n2] W4^(n2 k2), kept in x[n1 + 4 k2]
#pragma unroll
    for (int n1 = 0; n1 < 4; n1++) {
        if (HALF) {
            const f2 u = x[n1], v = x[n1 + 4];
            const f2 rv = INV ? f2{-v.y, v.x} : f2{v.y, -v.x};
            x[n1] = cadd(u, v);
            x[n1 + 4] = cadd(u, rv);
            x[n1 + 8] = csub(u, v);
            x[n1 + 12] = csub(u, rv);
        } else {
            dft4<INV>(x[n1], x[n1 + 4], x[n1 + 8], x[n1 + 12]);
        }
    }
    // twiddles W16^(n1 k2)
    x[1 + 4] = mul_w16<INV, 1>(x[1 + 4]);
    x[1 + 8] = mul_w16<INV, 2>(x[1 + 8]);
    x[1 + 12] = mul_w16<INV, 3>(x[1 + 12]);
    x[2 + 4] = mul_w16<INV, 2>(x[2 + 4]);
    x[2 + 8] = mul_w16<INV, 4>(x[2 + 8]);
    x[2 + 12] = mul_w16<INV, 6>(x[2 + 12]);
    x[3 + 4] = mul_w16<INV, 3>(x[3 + 4]);
    x[3 + 8] = mul_w16<INV, 6>(x[3 + 8]);
    x[3 + 12] = mul_w16<INV, 9>(x[3 + 12]);
    // X[k2 + 4 k1] = sum_n1 a[n1][k2] W4^(n1 k1)
    f2 y[16];
#pragma unroll
    for (int k2 = 0; k2 < 4; k2++) {
        f2 a = x[4 * k2], b = x[4 * k2 + 1], c = x[4 * k2 + 2], d = x[4 * k2 + 3];
        dft4<INV>(a, b, c, d);
        y[k2] = a;
        y[k2 + 4] = b;
        y[k2 + 8] = c;
        y[k2 + 12] = d;
    }
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = y[k];
}

// U += a Z, W += b Z with the real weights (a, b) = (A[k], A[M-k]) of this thread's 16 bins
RAR_HD void synth_accumulate(f2 (&U)[16], f2 (&W)[16], const f2 (&Z)[16], const f2 (&ab)[16]) {
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) {
        U[k2].x = fmaf(ab[k2].x, Z[k2].x, U[k2].x);
        U[k2].y = fmaf(ab[k2].x, Z[k2].y, U[k2].y);
        W[k2].x = fmaf(ab[k2].y, Z[k2].x, W[k2].x);
        W[k2].y = fmaf(ab[k2].y, Z[k2].y, W[k2].y);
    }
}

// split step on a pair: alpha z + beta conj(zm) with w = exp(-2 pi i k / 2M)   (rfft_split of rar_fft.cuh for one bin)
RAR_HD f2 synth_split(f2 z, f2 zm, f2 w) {
    const f2 xe = f2{0.5f * (z.x + zm.x), 0.5f * (z.y - zm.y)};
    const f2 xo = f2{0.5f * (z.y + zm.y), -0.5f * (z.x - zm.x)};
    return cadd(xe, cmul(w, xo));
}

// Spectrum of the segment's synthesised window, ready for the inverse packed transform, one bin k = t + 16 k2.
// Composing the split step (rfft_split: S[k] = alpha U[k] + beta conj(W[M-k])) with the merge step of the inverse
// (irfft_merge: Z'[k] = conj(alpha) S[k] + conj(beta) conj(S[M-k])), alpha = (1 - i w)/2, beta = (1 + i w)/2,
// w = exp(-i theta), theta = 2 pi k / 512, and alpha_{M-k} = conj(alpha), beta_{M-k} = conj(beta):
//        Z'[k] = |alpha|^2 U[k] + |beta|^2 W[k] + conj(alpha) beta conj(W[M-k]) + alpha conj(beta) conj(U[M-k])
//              = (1 - sin theta)/2 U[k] + (1 + sin theta)/2 W[k] + i (cos theta)/2 conj(W[M-k] - U[M-k]).
// Ep = W[M-k] - U[M-k] comes from the partner thread; w = (cos theta, -sin theta) from the table.
// Bin 0 packs (DC, Nyquist): DC = sum_b A_b[0] (Z.x + Z.y) = U.x + U.y, Nyquist = sum_b A_b[M] (Z.x - Z.y) = W.x - W.y.
RAR_HD f2 synth_merge1(bool first, f2 U, f2 W, f2 Ep, f2 w) {
    const float a = fmaf(0.5f, w.y, 0.5f), b = 1.0f - a, h = 0.5f * w.x;
    f2 z = f2{fmaf(a, U.x, fmaf(b, W.x, h * Ep.y)), fmaf(a, U.y, fmaf(b, W.y, h * Ep.x))};
    if (first) {
        const float x0 = U.x + U.y, xm = W.x - W.y;
        z = f2{0.5f * (x0 + xm), 0.5f * (x0 - xm)};
    }
    return z;
}

// Tables.  A table row belongs to a thread t and holds 16 elements e; it is stored so that the 16 threads of a
// transform read it with coalesced 16-byte loads: element e of thread t at f2 index synth_tab(t, e) of its 256-entry
// table (the float4 of elements 2j, 2j+1 of all t is contiguous over t).
// T[0..256): tw(t, k1) = exp(-2 pi i t k1 / 256); T[256..512): w2(t, k2) = exp(-2 pi i (t + 16 k2) / 512);
// T[512 + b*256 ...): (A_b[k], A_b[256 - k]) at (t, k2), k = t + 16 k2, with A_b[k] = sum_j g_b[j] cos(2 pi k (j - 127) / 512)
// the zero-phase amplitude of band b's filter (taps: [bands][256], 255 taps and a zero).  synth_table_len(bands) f2 elements.
RAR_HD int synth_tab(int t, int e) { return (((e >> 1) * 16 + t) << 1) + (e & 1); }
inline size_t synth_table_len(int bands) { return 512 + (size_t)bands * 256; }
inline void synth_tables(const float *taps, int bands, f2 *T) {
    const double two_pi = 6.283185307179586476925286766559;
    for (int t = 0; t < 16; t++)
        for (int j = 0; j < 16; j++) {
            const double a = two_pi * (double)(t * j) / 256.0, b = two_pi * (double)(t + 16 * j) / 512.0;
            T[synth_tab(t, j)] = f2{(float)cos(a), (float)-sin(a)};
            T[256 + synth_tab(t, j)] = f2{(float)cos(b), (float)-sin(b)};
        }
    double c512[512];
    for (int i = 0; i < 512; i++) c512[i] = cos(two_pi * (double)i / 512.0);
    for (int b = 0; b < bands; b++) {
        double A[257];
        for (int k = 0; k <= 256; k++) {
            double acc = 0.0;
            for (int j = 0; j < 255; j++) acc += (double)taps[(size_t)b * 256 + j] * c512[(k * (j - 127 + 512)) & 511];
            A[k] = acc;
        }
        for (int t = 0; t < 16; t++)
            for (int k2 = 0; k2 < 16; k2++) {
                const int k = t + 16 * k2;
                T[512 + (size_t)b * 256 + synth_tab(t, k2)] = f2{(float)A[k], (float)A[256 - k]};
            }
    }
}

}  // namespace rar

// conv_kernels.cu -- uniformly partitioned overlap-save FFT convolution for sm_100a.
//
// Replaces AudioConvolve.compute:13-31 (direct-form O(N*M), one thread per output sample).  Block
// B = 256 samples, window 2B = 512, IR split into partitions of B taps.  With X[j] the spectrum of
// the input window ending at block j and H[p] the spectrum of IR partition p,
//        Y[j] = sum_p X[j-p] * H[p],     output block j = last B samples of irfft(Y[j]).
// The FFTs are hand-written shared-memory radix-4 Stockham transforms (rar_fft.cuh); the
// complex-multiply-accumulate over partitions streams both spectra from HBM with 16-byte loads and
// is the bandwidth-bound kernel of the streaming convolver (BASELINE config 5: 2 x 1875 x 2 KB per
// stream per block).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include <mutex>

#include "rar_fft.cuh"
#include "rar_internal.h"

namespace rar {
namespace {

constexpr int kB = 256;            // block
constexpr int kSub = 4;            // transforms per CTA in the FFT kernels
constexpr float kConvEps = 1e-4f;  // Common.hlsl:4, used by AudioConvolve.compute:25

__device__ f2 g_tw[kFftM];           // exp(-2 pi i k / 256)
__device__ f2 g_tw2[kFftM / 2 + 1];  // exp(-2 pi i k / 512)

struct FftSmem {
    f2 a[kSub][kFftM];
    f2 b[kSub][kFftM];
    f2 tw[kFftM];
    f2 tw2[kFftM / 2 + 1];
};

__device__ __forceinline__ void load_tables(FftSmem &s) {
    for (int k = threadIdx.x; k < kFftM; k += blockDim.x) s.tw[k] = g_tw[k];
    for (int k = threadIdx.x; k <= kFftM / 2; k += blockDim.x) s.tw2[k] = g_tw2[k];
}

// Four radix-4 passes a -> b -> a -> b -> a; the result is in a.  Block-wide barriers between passes.
__device__ __forceinline__ void fft256_inplace(f2 *a, f2 *b, int i, const f2 *tw, bool inverse) {
    fft_pass_r4(a, b, i, 1, tw, inverse);
    __syncthreads();
    fft_pass_r4(b, a, i, 4, tw, inverse);
    __syncthreads();
    fft_pass_r4(a, b, i, 16, tw, inverse);
    __syncthreads();
    fft_pass_r4(b, a, i, 64, tw, inverse);
    __syncthreads();
}

// Loads the real window w[m] = src(first + m), m in [0, 512), packed as z[n] = (w[2n], w[2n+1]), forward
// transforms it and leaves the packed half spectrum in s.b[sub].  `fetch(t)` returns sample t or 0.
template <class Fetch>
__device__ __forceinline__ void rfft512_to_b(FftSmem &s, int sub, int i, Fetch fetch) {
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int n = i + 64 * m;
        s.a[sub][n] = f2{fetch(2 * n), fetch(2 * n + 1)};
    }
    __syncthreads();
    fft256_inplace(s.a[sub], s.b[sub], i, s.tw, false);
    for (int k = i; k <= kFftM / 2; k += 64) rfft_split(s.a[sub], s.b[sub], k, s.tw2);
    __syncthreads();
}

// Packed half spectrum in s.b[sub] -> real window; afterwards s.a[sub][n] = (w[2n], w[2n+1]) * M.
__device__ __forceinline__ void irfft512_from_b(FftSmem &s, int sub, int i) {
    for (int k = i; k <= kFftM / 2; k += 64) irfft_merge(s.b[sub], s.a[sub], k, s.tw2);
    __syncthreads();
    fft256_inplace(s.a[sub], s.b[sub], i, s.tw, true);
}

__global__ void fixed_to_float_kernel(const long long *__restrict__ hist, float *__restrict__ out, long long n, float scale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = ((float)hist[i] * 9.094947017729282e-13f) * scale;
}

// All-reduce step of the single-process multi-GPU path: the root device sums its peers' histograms, read
// straight from peer memory (NVLink P2P loads).
__global__ void peer_reduce_kernel(long long *__restrict__ hist, const PeerHists peers, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long acc = hist[i];
        for (int k = 0; k < peers.n; k++) acc += peers.p[k][i];
        hist[i] = acc;
    }
}

__global__ void float_to_fixed_kernel(const float *__restrict__ in, long long *__restrict__ hist, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        hist[i] = quantize_energy(in[i]);
}

// H[p] = rfft([ir[pB .. pB+B), 0 x B]); blockIdx.y = response of a batch (ir and H advance by their strides)
__global__ void __launch_bounds__(64 * kSub) ir_spectra_kernel(const float *__restrict__ ir, long long ir_stride, int ir_len,
                                                                f2 *__restrict__ H, long long h_stride, int n_part) {
    __shared__ FftSmem s;
    const int sub = threadIdx.x >> 6, i = threadIdx.x & 63;
    const int p = blockIdx.x * kSub + sub;
    ir += (long long)blockIdx.y * ir_stride;
    H += (long long)blockIdx.y * h_stride;
    load_tables(s);
    const long long first = (long long)p * kB;
    rfft512_to_b(s, sub, i, [&](int t) -> float {
        const long long g = first + t;
        return (p < n_part && t < kB && g < ir_len) ? ir[g] : 0.0f;
    });
    if (p < n_part) {
#pragma unroll
        for (int m = 0; m < 4; m++) H[(size_t)p * kFftM + i + 64 * m] = s.b[sub][i + 64 * m];
    }
}

// ---- filter-bank synthesis of a broadband response from a banded slot ----------------------------
//
// A banded slot holds IR[bin * bands + band] (RaytraceOcclusion2D.compute:241-248); band b carries the energy that
// arrives in the frequency range [edge_b, edge_b+1).  The broadband response a convolution needs is
//        h[n] = sum_b (g_b * h_b)[n + D],   h_b[n] = IR[(n / stride) * bands + b] if stride divides n, else 0,
// with g_b the 255-tap windowed-sinc band-pass filter of band b (linear phase, delay D = 127).  The ideal band-pass
// responses of contiguous bands telescope to a unit impulse at D and the window is 1 there, so sum_b g_b is exactly
// that impulse: a slot whose bands are all equal synthesises to that common response, sample for sample.
// Computed per 256-sample segment as overlap-add in the frequency domain: S[p] = sum_b G_b . rfft512(h_b segment p),
// irfft512, and the 512 results added into out[256 p - D ...].  Every output sample receives exactly two
// contributions (segments p and p-1) onto a zeroed buffer, and a + b = b + a, so the atomic adds are deterministic.
//
// One launch handles a batch of slots of equal shape (blockIdx.y).  VEC (bands a multiple of 4, stride 1): a thread
// first fetches ALL the words it will need of a chunk of 4 bands -- its four samples x 32 contiguous bytes, eight
// independent 16-byte loads, every fetched sector fully used -- and then runs the four transforms out of registers
// (chunks of 8 bands needed 75 registers: 3 CTAs per SM instead of 5); otherwise each band's samples are gathered
// with strided scalar loads.
struct BandAcc {
    f2 v[4];
};
__device__ __forceinline__ void band_accumulate(BandAcc &acc, const FftSmem &s, int sub, int i, const f2 *__restrict__ gb) {
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int k = i + 64 * m;
        const f2 x = s.b[sub][k], g = gb[k];
        if (k == 0) {  // packed bin 0 = (DC, Nyquist), both real
            acc.v[m].x = fmaf(x.x, g.x, acc.v[m].x);
            acc.v[m].y = fmaf(x.y, g.y, acc.v[m].y);
        } else {
            acc.v[m].x += x.x * g.x - x.y * g.y;
            acc.v[m].y += x.x * g.y + x.y * g.x;
        }
    }
}

template <bool VEC>
__global__ void __launch_bounds__(64 * kSub, 4) band_synth_kernel(const __grid_constant__ BandSynthBatch batch, int bins, int bands, int stride,
                                                                const f2 *__restrict__ G, int out_len, int n_seg, int delay) {
    __shared__ FftSmem s;
    const int sub = threadIdx.x >> 6, i = threadIdx.x & 63;
    const int p = blockIdx.x * kSub + sub;
    const BandSynthItem it = batch.items[blockIdx.y];
    const long long *__restrict__ hist = it.hist;
    const float scale = it.scale;
    load_tables(s);
    const long long first = (long long)p * kB;
    const long long n_samples = (long long)bins * stride;
    BandAcc acc;
#pragma unroll
    for (int m = 0; m < 4; m++) acc.v[m] = f2{0.f, 0.f};
    if (VEC) {
        // this thread's samples of the segment: t = 2i, 2i+1 (n = i) and 2i+128, 2i+129 (n = i+64); n >= 128 is padding
        constexpr int CH = 4;
        for (int c0 = 0; c0 < bands; c0 += CH) {
            float f[4][CH];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const long long g = first + 2 * i + (q & 1) + 128 * (q >> 1);
                const bool on = p < n_seg && g < n_samples;
                const longlong2 *src = reinterpret_cast<const longlong2 *>(hist + (on ? g : 0) * bands + c0);
#pragma unroll
                for (int k = 0; k < CH / 2; k++) {
                    longlong2 w = on ? __ldcs(src + k) : make_longlong2(0, 0);
                    f[q][2 * k] = ((float)w.x * 9.094947017729282e-13f) * scale;
                    f[q][2 * k + 1] = ((float)w.y * 9.094947017729282e-13f) * scale;
                }
            }
#pragma unroll
            for (int b = 0; b < CH; b++) {
                s.a[sub][i] = f2{f[0][b], f[1][b]};
                s.a[sub][i + 64] = f2{f[2][b], f[3][b]};
                s.a[sub][i + 128] = f2{0.f, 0.f};
                s.a[sub][i + 192] = f2{0.f, 0.f};
                __syncthreads();
                fft256_inplace(s.a[sub], s.b[sub], i, s.tw, false);
                for (int k = i; k <= kFftM / 2; k += 64) rfft_split(s.a[sub], s.b[sub], k, s.tw2);
                __syncthreads();
                band_accumulate(acc, s, sub, i, G + (size_t)(c0 + b) * kFftM);
                __syncthreads();  // s.a / s.b are reused by the next band
            }
        }
    } else {
        for (int b = 0; b < bands; b++) {
            rfft512_to_b(s, sub, i, [&](int t) -> float {
                const long long g = first + t;
                if (p >= n_seg || t >= kB || g >= n_samples) return 0.0f;
                long long bin = g;
                if (stride > 1) {
                    bin = g / stride;
                    if (bin * stride != g) return 0.0f;
                }
                return ((float)__ldg(hist + bin * bands + b) * 9.094947017729282e-13f) * scale;
            });
            band_accumulate(acc, s, sub, i, G + (size_t)b * kFftM);
            __syncthreads();  // s.a / s.b are reused by the next band
        }
    }
#pragma unroll
    for (int m = 0; m < 4; m++) s.b[sub][i + 64 * m] = acc.v[m];
    __syncthreads();
    irfft512_from_b(s, sub, i);
    if (p < n_seg) {
        const float inv = 1.0f / (float)kFftM;
        float *__restrict__ out = it.out;
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const int n = i + 64 * m;
            const f2 v = s.a[sub][n];
            const long long o = first + 2 * n - delay;
            if (o >= 0 && o < out_len) atomicAdd(out + o, v.x * inv);
            if (o + 1 >= 0 && o + 1 < out_len) atomicAdd(out + o + 1, v.y * inv);
        }
    }
}

// X[j] = rfft(x[(j-1)B .. (j+1)B)) with out-of-range samples and |x| <= eps samples as zero.
__global__ void __launch_bounds__(64 * kSub) input_spectra_kernel(const float *__restrict__ x, int x_len, f2 *__restrict__ X,
                                                                   int n_xwin) {
    __shared__ FftSmem s;
    const int sub = threadIdx.x >> 6, i = threadIdx.x & 63;
    const int j = blockIdx.x * kSub + sub;
    load_tables(s);
    const long long first = ((long long)j - 1) * kB;
    rfft512_to_b(s, sub, i, [&](int t) -> float {
        const long long g = first + t;
        if (j >= n_xwin || g < 0 || g >= x_len) return 0.0f;
        const float v = x[g];
        return fabsf(v) > kConvEps ? v : 0.0f;
    });
    if (j < n_xwin) {
#pragma unroll
        for (int m = 0; m < 4; m++) X[(size_t)j * kFftM + i + 64 * m] = s.b[sub][i + 64 * m];
    }
}

// Accumulators of one float4 (two packed bins).  Bin A keeps a.x*b.x and a.y*b.y apart so that packed
// bin 0 (DC, Nyquist) can be finished as two real products.
struct Acc4 {
    float ac, bd, im, re1, im1;
};
__device__ __forceinline__ void cmac4(Acc4 &s, const float4 x, const float4 h) {
    s.ac = fmaf(x.x, h.x, s.ac);
    s.bd = fmaf(x.y, h.y, s.bd);
    s.im = fmaf(x.x, h.y, fmaf(x.y, h.x, s.im));
    s.re1 = fmaf(x.z, h.z, fmaf(-x.w, h.w, s.re1));
    s.im1 = fmaf(x.z, h.w, fmaf(x.w, h.z, s.im1));
}
__device__ __forceinline__ float4 finish4(const Acc4 &s, bool packed_bin0) {
    return packed_bin0 ? make_float4(s.ac, s.bd, s.re1, s.im1) : make_float4(s.ac - s.bd, s.im, s.re1, s.im1);
}

// Y[j] = sum_{p} X[j-p] * H[p] over the valid p; one CTA of 128 threads per output block.
__global__ void __launch_bounds__(128) block_cmac_kernel(const float4 *__restrict__ X, int n_xwin, const float4 *__restrict__ H,
                                                          int n_part, float4 *__restrict__ Y) {
    const int j = blockIdx.x, t = threadIdx.x;
    int p_lo = j - (n_xwin - 1);
    if (p_lo < 0) p_lo = 0;
    int p_hi = j < n_part - 1 ? j : n_part - 1;
    Acc4 s = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int p = p_lo; p <= p_hi; p++) cmac4(s, __ldg(X + (size_t)(j - p) * 128 + t), __ldg(H + (size_t)p * 128 + t));
    Y[(size_t)j * 128 + t] = finish4(s, t == 0);
}

// out[jB + i] = irfft(Y[j])[B + i] * scale
__global__ void __launch_bounds__(64 * kSub) output_blocks_kernel(const f2 *__restrict__ Y, int n_out_blocks, float *__restrict__ out,
                                                                   int out_len, float scale) {
    __shared__ FftSmem s;
    const int sub = threadIdx.x >> 6, i = threadIdx.x & 63;
    const int j = blockIdx.x * kSub + sub;
    load_tables(s);
#pragma unroll
    for (int m = 0; m < 4; m++)
        s.b[sub][i + 64 * m] = j < n_out_blocks ? Y[(size_t)j * kFftM + i + 64 * m] : f2{0.f, 0.f};
    __syncthreads();
    irfft512_from_b(s, sub, i);
    if (j < n_out_blocks) {
#pragma unroll
        for (int m = 0; m < 2; m++) {
            const int n = 128 + i + 64 * m;  // second half of the window
            const long long o = (long long)j * kB + 2 * (n - 128);
            const f2 v = s.a[sub][n];
            // sample out_len-1 lies beyond the N+M-1 samples of the linear convolution: the reference's
            // loop is empty there (AudioConvolve.compute:19-20) and writes an exact 0.
            if (o < out_len) out[o] = o == out_len - 1 ? 0.0f : v.x * scale;
            if (o + 1 < out_len) out[o + 1] = o + 1 == out_len - 1 ? 0.0f : v.y * scale;
        }
    }
}

// ---- streaming convolver ------------------------------------------------------------------------------

// Window = [prev block | new block] per stream -> fdl[s][head]; prev <- new (after eps zeroing).
__global__ void __launch_bounds__(64 * kSub) stream_input_kernel(const float *__restrict__ in, float *__restrict__ prev,
                                                                  f2 *__restrict__ fdl, int n_streams, int n_part, int head) {
    __shared__ FftSmem s;
    const int sub = threadIdx.x >> 6, i = threadIdx.x & 63;
    const int st = blockIdx.x * kSub + sub;
    load_tables(s);
    const bool on = st < n_streams;
    const float *pv = prev + (size_t)st * kB;
    const float *nw = in + (size_t)st * kB;
    rfft512_to_b(s, sub, i, [&](int t) -> float {
        if (!on) return 0.0f;
        if (t < kB) return pv[t];
        const float v = nw[t - kB];
        return fabsf(v) > kConvEps ? v : 0.0f;
    });
    if (on) {
        f2 *dst = fdl + ((size_t)st * n_part + head) * kFftM;
#pragma unroll
        for (int m = 0; m < 4; m++) dst[i + 64 * m] = s.b[sub][i + 64 * m];
    }
    __syncthreads();  // every read of prev is done
    if (on) {
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const float v = nw[i + 64 * m];
            prev[(size_t)st * kB + i + 64 * m] = fabsf(v) > kConvEps ? v : 0.0f;
        }
    }
}

constexpr int kCmacTy = 4;

// partial[s][split] = sum over the split's partitions of fdl[s][(head - p) mod P] * H[s][p].
// grid (n_split, S), block (128, kCmacTy).  This is the HBM-bound kernel.
// LIST: blockIdx.y indexes `list`, the streams whose impulse response is being cross-faded (second pass with the new
// spectra); otherwise it is the stream itself and the code is the plain hot kernel.
template <bool LIST>
__global__ void __launch_bounds__(128 * kCmacTy) stream_cmac_kernel(const float4 *__restrict__ fdl, const float4 *__restrict__ H,
                                                                     float4 *__restrict__ partial, int n_part, int per_split,
                                                                     int head, const int *__restrict__ list) {
    __shared__ float4 red[kCmacTy][128];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int st = LIST ? list[blockIdx.y] : (int)blockIdx.y, split = blockIdx.x;
    const int p0 = split * per_split;
    const int p1 = min(p0 + per_split, n_part);
    const float4 *fs = fdl + (size_t)st * n_part * 128 + tx;
    const float4 *hs = H + (size_t)st * n_part * 128 + tx;
    Acc4 s = {0.f, 0.f, 0.f, 0.f, 0.f};
    int p = p0 + ty;
    for (; p + 3 * kCmacTy < p1; p += 4 * kCmacTy) {
        float4 x[4], h[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int pp = p + u * kCmacTy;
            int slot = head - pp;
            if (slot < 0) slot += n_part;
            x[u] = __ldcs(fs + (size_t)slot * 128);
            h[u] = __ldcs(hs + (size_t)pp * 128);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) cmac4(s, x[u], h[u]);
    }
    for (; p < p1; p += kCmacTy) {
        int slot = head - p;
        if (slot < 0) slot += n_part;
        cmac4(s, __ldcs(fs + (size_t)slot * 128), __ldcs(hs + (size_t)p * 128));
    }
    red[ty][tx] = finish4(s, tx == 0);
    __syncthreads();
    if (ty == 0) {
        float4 r = red[0][tx];
#pragma unroll
        for (int k = 1; k < kCmacTy; k++) {
            const float4 v = red[k][tx];
            r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
        }
        partial[((size_t)st * gridDim.x + split) * 128 + tx] = r;
    }
}

// out[s] = last B samples of irfft(sum over splits of partial[s][split]) * (1/M)
__global__ void __launch_bounds__(64 * kSub) stream_output_kernel(const f2 *__restrict__ partial, int n_split, float *__restrict__ out,
                                                                   int n_streams) {
    __shared__ FftSmem s;
    const int sub = threadIdx.x >> 6, i = threadIdx.x & 63;
    const int st = blockIdx.x * kSub + sub;
    load_tables(s);
    const bool on = st < n_streams;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        f2 acc = f2{0.f, 0.f};
        if (on) {
            const f2 *src = partial + (size_t)st * n_split * kFftM + i + 64 * m;
            for (int q = 0; q < n_split; q++) acc = cadd(acc, src[(size_t)q * kFftM]);
        }
        s.b[sub][i + 64 * m] = acc;
    }
    __syncthreads();
    irfft512_from_b(s, sub, i);
    if (on) {
        const float scale = 1.0f / (float)kFftM;
#pragma unroll
        for (int m = 0; m < 2; m++) {
            const int n = 128 + i + 64 * m;
            const f2 v = s.a[sub][n];
            float2 *o = reinterpret_cast<float2 *>(out + (size_t)st * kB + 2 * (n - 128));
            *o = make_float2(v.x * scale, v.y * scale);
        }
    }
}

// The output step of a block in which some streams change their impulse response: those streams get
// out[n] = y_old[n] + w[n] (y_new[n] - y_old[n]), w[n] = (n + 1) / B, y_old from `partial` (spectra in use so far) and
// y_new from `partial2` (the new spectra, same input history); the others are written as stream_output_kernel does.
__global__ void __launch_bounds__(64 * kSub) stream_output_fade_kernel(const f2 *__restrict__ partial, const f2 *__restrict__ partial2,
                                                                        const int *__restrict__ fade, int n_split,
                                                                        float *__restrict__ out, int n_streams) {
    __shared__ FftSmem s;
    const int sub = threadIdx.x >> 6, i = threadIdx.x & 63;
    const int st = blockIdx.x * kSub + sub;
    load_tables(s);
    const bool on = st < n_streams;
    bool any = false;  // block-uniform: the second transform has block-wide barriers
    for (int k = 0; k < kSub; k++) {
        const int t = blockIdx.x * kSub + k;
        if (t < n_streams && fade[t] != 0) any = true;
    }
    const bool mine = on && fade[st] != 0;
    const float scale = 1.0f / (float)kFftM;
    float y[2][2];
#pragma unroll
    for (int m = 0; m < 4; m++) {
        f2 acc = f2{0.f, 0.f};
        if (on) {
            const f2 *src = partial + (size_t)st * n_split * kFftM + i + 64 * m;
            for (int q = 0; q < n_split; q++) acc = cadd(acc, src[(size_t)q * kFftM]);
        }
        s.b[sub][i + 64 * m] = acc;
    }
    __syncthreads();
    irfft512_from_b(s, sub, i);
#pragma unroll
    for (int m = 0; m < 2; m++) {
        const f2 v = s.a[sub][128 + i + 64 * m];
        y[m][0] = v.x * scale;
        y[m][1] = v.y * scale;
    }
    if (any) {
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 4; m++) {
            f2 acc = f2{0.f, 0.f};
            if (mine) {
                const f2 *src = partial2 + (size_t)st * n_split * kFftM + i + 64 * m;
                for (int q = 0; q < n_split; q++) acc = cadd(acc, src[(size_t)q * kFftM]);
            }
            s.b[sub][i + 64 * m] = acc;
        }
        __syncthreads();
        irfft512_from_b(s, sub, i);
        if (mine) {
#pragma unroll
            for (int m = 0; m < 2; m++) {
                const int k0 = 2 * (i + 64 * m);  // sample index inside the block
                const f2 v = s.a[sub][128 + i + 64 * m];
                const float w0 = (float)(k0 + 1) * (1.0f / (float)kB), w1 = (float)(k0 + 2) * (1.0f / (float)kB);
                y[m][0] += w0 * (v.x * scale - y[m][0]);
                y[m][1] += w1 * (v.y * scale - y[m][1]);
            }
        }
    }
    if (on) {
#pragma unroll
        for (int m = 0; m < 2; m++) {
            float2 *o = reinterpret_cast<float2 *>(out + (size_t)st * kB + 2 * (i + 64 * m));
            *o = make_float2(y[m][0], y[m][1]);
        }
    }
}

inline int blocks_for(long long n, int per) { return (int)((n + per - 1) / per); }

}  // namespace

void conv_init_tables(cudaStream_t stream) {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return;
    f2 tw[kFftM], tw2[kFftM / 2 + 1];
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < kFftM; k++) tw[k] = f2{(float)cos(two_pi * k / kFftM), (float)-sin(two_pi * k / kFftM)};
    for (int k = 0; k <= kFftM / 2; k++)
        tw2[k] = f2{(float)cos(two_pi * k / (2 * kFftM)), (float)-sin(two_pi * k / (2 * kFftM))};
    // on the caller's stream (the context streams are non-blocking: nothing orders them against the legacy stream);
    // the host arrays are on this stack frame, so the copies are awaited before returning
    cudaMemcpyToSymbolAsync(g_tw, tw, sizeof tw, 0, cudaMemcpyHostToDevice, stream);
    cudaMemcpyToSymbolAsync(g_tw2, tw2, sizeof tw2, 0, cudaMemcpyHostToDevice, stream);
    cudaStreamSynchronize(stream);
    if (dev >= 0 && dev < 64) done[dev] = true;
}

cudaError_t launch_fixed_to_float(const long long *hist, float *out, long long n, float scale, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    fixed_to_float_kernel<<<min(blocks_for(n, 256), 4096), 256, 0, s>>>(hist, out, n, scale);
    return cudaGetLastError();
}

cudaError_t launch_peer_reduce(long long *hist, PeerHists peers, long long n, cudaStream_t s) {
    if (n <= 0 || peers.n <= 0) return cudaSuccess;
    peer_reduce_kernel<<<min(blocks_for(n, 256), 1184), 256, 0, s>>>(hist, peers, n);
    return cudaGetLastError();
}

cudaError_t launch_float_to_fixed(const float *in, long long *hist, long long n, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    float_to_fixed_kernel<<<min(blocks_for(n, 256), 4096), 256, 0, s>>>(in, hist, n);
    return cudaGetLastError();
}

cudaError_t launch_ir_spectra(const float *ir_f, int ir_len, float2 *H, int n_part, int block, cudaStream_t s) {
    return launch_ir_spectra_batch(ir_f, 0, ir_len, H, 0, n_part, 1, block, s);
}

cudaError_t launch_ir_spectra_batch(const float *ir_f, long long ir_stride, int ir_len, float2 *H, long long h_stride, int n_part,
                                    int n_items, int block, cudaStream_t s) {
    if (block != kB) return cudaErrorInvalidValue;
    if (n_part <= 0 || n_items <= 0) return cudaSuccess;
    static const bool no_fast = [] { const char *e = getenv("RAR_NO_FAST_SYNTH"); return e && e[0] == '1'; }();
    if (!no_fast && ir_spectra16_applicable(ir_f, ir_stride)) return launch_ir_spectra16(ir_f, ir_stride, ir_len, H, h_stride, n_part, n_items, s);
    ir_spectra_kernel<<<dim3(blocks_for(n_part, kSub), n_items), 64 * kSub, 0, s>>>(ir_f, ir_stride, ir_len, reinterpret_cast<f2 *>(H),
                                                                                     h_stride, n_part);
    return cudaGetLastError();
}

cudaError_t launch_band_synth(const BandSynthBatch &batch, int n_items, int bins, int bands, int stride, const float2 *G, int out_len,
                              cudaStream_t s) {
    if (n_items <= 0 || bins <= 0 || bands <= 0 || stride <= 0 || out_len <= 0) return cudaSuccess;
    if (n_items > kBandSynthBatch) return cudaErrorInvalidValue;
    const long long n_samples = (long long)bins * stride;
    const int n_seg = (int)((n_samples + kB - 1) / kB);
    const dim3 grid(blocks_for(n_seg, kSub), n_items);
    if (bands % 4 == 0 && stride == 1)
        band_synth_kernel<true><<<grid, 64 * kSub, 0, s>>>(batch, bins, bands, stride, reinterpret_cast<const f2 *>(G), out_len, n_seg,
                                                           kBandFilterDelay);
    else
        band_synth_kernel<false><<<grid, 64 * kSub, 0, s>>>(batch, bins, bands, stride, reinterpret_cast<const f2 *>(G), out_len, n_seg,
                                                            kBandFilterDelay);
    return cudaGetLastError();
}

cudaError_t launch_input_spectra(const float *x, int x_len, float2 *X, int n_xwin, int block, cudaStream_t s) {
    if (block != kB) return cudaErrorInvalidValue;
    if (n_xwin <= 0) return cudaSuccess;
    input_spectra_kernel<<<blocks_for(n_xwin, kSub), 64 * kSub, 0, s>>>(x, x_len, reinterpret_cast<f2 *>(X), n_xwin);
    return cudaGetLastError();
}

cudaError_t launch_block_cmac(const float2 *X, int n_xwin, const float2 *H, int n_part, float2 *Y, int n_out_blocks,
                              int block, cudaStream_t s) {
    if (block != kB) return cudaErrorInvalidValue;
    if (n_out_blocks <= 0) return cudaSuccess;
    block_cmac_kernel<<<n_out_blocks, 128, 0, s>>>(reinterpret_cast<const float4 *>(X), n_xwin,
                                                   reinterpret_cast<const float4 *>(H), n_part, reinterpret_cast<float4 *>(Y));
    return cudaGetLastError();
}

cudaError_t launch_output_blocks(const float2 *Y, int n_out_blocks, float *out, int out_len, float scale, int block,
                                 cudaStream_t s) {
    if (block != kB) return cudaErrorInvalidValue;
    if (n_out_blocks <= 0) return cudaSuccess;
    output_blocks_kernel<<<blocks_for(n_out_blocks, kSub), 64 * kSub, 0, s>>>(reinterpret_cast<const f2 *>(Y), n_out_blocks, out,
                                                                              out_len, scale);
    return cudaGetLastError();
}

cudaError_t launch_stream_step(const StreamConv &c, const float *d_in, float *d_out, cudaStream_t s, int *launches) {
    if (c.block != kB) return cudaErrorInvalidValue;
    stream_input_kernel<<<blocks_for(c.n_streams, kSub), 64 * kSub, 0, s>>>(d_in, c.prev, reinterpret_cast<f2 *>(c.fdl),
                                                                           c.n_streams, c.n_part, c.head);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    stream_cmac_kernel<false><<<dim3(c.n_split, c.n_streams), dim3(128, kCmacTy), 0, s>>>(
        reinterpret_cast<const float4 *>(c.fdl), reinterpret_cast<const float4 *>(c.H), reinterpret_cast<float4 *>(c.partial),
        c.n_part, c.part_per_split, c.head, nullptr);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    stream_output_kernel<<<blocks_for(c.n_streams, kSub), 64 * kSub, 0, s>>>(reinterpret_cast<const f2 *>(c.partial), c.n_split,
                                                                            d_out, c.n_streams);
    e = cudaGetLastError();
    if (e == cudaSuccess && launches) *launches += 3;
    return e;
}

cudaError_t launch_stream_step_fade(const StreamConv &c, const float2 *H2, float2 *partial2, const int *d_fade, const int *d_list,
                                    int n_list, const float *d_in, float *d_out, cudaStream_t s, int *launches) {
    if (c.block != kB || n_list <= 0) return cudaErrorInvalidValue;
    stream_input_kernel<<<blocks_for(c.n_streams, kSub), 64 * kSub, 0, s>>>(d_in, c.prev, reinterpret_cast<f2 *>(c.fdl),
                                                                           c.n_streams, c.n_part, c.head);
    stream_cmac_kernel<false><<<dim3(c.n_split, c.n_streams), dim3(128, kCmacTy), 0, s>>>(
        reinterpret_cast<const float4 *>(c.fdl), reinterpret_cast<const float4 *>(c.H), reinterpret_cast<float4 *>(c.partial),
        c.n_part, c.part_per_split, c.head, nullptr);
    stream_cmac_kernel<true><<<dim3(c.n_split, n_list), dim3(128, kCmacTy), 0, s>>>(
        reinterpret_cast<const float4 *>(c.fdl), reinterpret_cast<const float4 *>(H2), reinterpret_cast<float4 *>(partial2),
        c.n_part, c.part_per_split, c.head, d_list);
    stream_output_fade_kernel<<<blocks_for(c.n_streams, kSub), 64 * kSub, 0, s>>>(
        reinterpret_cast<const f2 *>(c.partial), reinterpret_cast<const f2 *>(partial2), d_fade, c.n_split, d_out, c.n_streams);
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && launches) *launches += 4;
    return e;
}

}  // namespace rar

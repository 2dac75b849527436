// band_synth.cu -- filter-bank synthesis of a broadband response from a banded slot, production kernel
// for slots whose band count is a multiple of 4 on the sample grid (time stride 1).  What it computes is stated in
// conv_kernels.cu (band_synth_kernel, the generic kernel that remains for the other shapes):
//        out[n] = sum_b (g_b * h_b)[n + 127],   h_b[n] = IR[n * bands + b] * 2^-40 * scale,
// per 256-sample segment as a 512-point circular convolution in the frequency domain, overlap-added into `out`.
//
// Layout of the work.  A warp owns one segment.  Its two halves (16 threads each) take four bands each of every group
// of eight and hold their transforms in registers (rar_synth16.cuh: 16 x 16 split, one transposition through shared
// memory per transform, real weights); the partner exchange of the split step is done by shuffles once per segment,
// and the halves' two windows are added by shuffles before the output step.  The warp stages the 64-bit histogram words of eight bands at a time
// through shared memory: every lane issues 16-byte loads, a warp-wide load reads 512 contiguous bytes (8 samples x 8
// bands), each word is converted once and stored band-major so that a transform's inputs are conflict-free 8-byte
// shared loads.  (A first version let each half warp run a whole segment and staged four bands at a time: it read
// half of every 64 bytes per pass and DRAM traffic was 1.5 x the histogram -- profiles/r02_ncu_band_synth16.txt.)
// Nothing in the main loop needs a block-wide barrier: everything a transform shares lives in one warp (__syncwarp).
// The transposition twiddles are staged in shared memory; the per-band weights (2 KB per band) come from global
// memory through L1 with coalesced 16-byte loads.
//
// HBM roofline: 8 bytes per (sample, band) read once, 4 bytes per sample added into the response (twice, by the two
// windows that cover it; the second add hits L2).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <mutex>

#include "rar_internal.h"
#include "rar_synth16.cuh"

namespace rar {
namespace {

constexpr int kWarps = 4;                // segments per CTA
constexpr int kRow = 260;                // floats per staged band row: 256 samples + 4 (rows 0/2/4/6 start 8 banks apart)
constexpr int kBatch = 8;                // 16-byte loads a lane keeps in flight while staging
constexpr int kStage = 8;                // bands staged at a time (four per half warp)
constexpr int kTRow = 18;                // f2 per transposition row: 16 + 2 (144-byte rows: conflict-free 16-byte reads)
constexpr float kQ = 9.094947017729282e-13f;  // 2^-40

struct SynthSmem {
    float stage[kWarps][kStage][kRow];
    f2 tr[kWarps][2][16][kTRow];
    float4 tw[128];  // the transposition twiddles, table layout (synth_tab)
};

__device__ __forceinline__ longlong2 ld_stream16(const long long *p) {
    longlong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}

// the 16 elements of thread t's row of a table (synth_tab layout)
template <class Ld>
__device__ __forceinline__ void load_row16(const float4 *__restrict__ tab, int t, f2 (&v)[16], Ld ld) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float4 a = ld(tab + j * 16 + t);
        v[2 * j] = f2{a.x, a.y};
        v[2 * j + 1] = f2{a.z, a.w};
    }
}

// transposition: this thread's y[k1] goes to row k1, column t; it then takes row t.  `mask`: the 16 lanes of this half
// warp (a half may run this while the other half has no bands left, so the warp-wide mask would wait for lanes that
// are elsewhere).
__device__ __forceinline__ void transpose16(f2 (*tr)[kTRow], int t, f2 (&y)[16], unsigned mask) {
#pragma unroll
    for (int k = 0; k < 16; k++) tr[k][t] = y[k];
    __syncwarp(mask);
    const float4 *q = reinterpret_cast<const float4 *>(&tr[t][0]);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float4 a = q[j];
        y[2 * j] = f2{a.x, a.y};
        y[2 * j + 1] = f2{a.z, a.w};
    }
    __syncwarp(mask);  // the rows are rewritten by the next transform
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// bytes: a multiple of 16, p 16-byte aligned
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes));
}

// ONE: at most eight bands, i.e. one staging group -- the sums are not live while the group is staged, so a lane
// keeps sixteen loads in flight instead of eight.
//
// Persistent warps: a warp takes (slot, segment) pairs from a global counter.  The pair after the current one is
// claimed before the current one is processed, and its first group of words is prefetched into L2 while the current
// segment's transforms run (a warp alternates between a load phase and some 4 000 instructions of arithmetic, and
// there are only four warps per scheduler to cover each other's load phases: ncu showed 30 % of the stall samples
// on the first use of a staged word).
template <bool ONE>
__global__ void __launch_bounds__(32 * kWarps, 4)
band_synth16_kernel(const __grid_constant__ BandSynthBatch batch, int bins, int bands, const float4 *__restrict__ T, int out_len, int n_seg,
                    int n_work, unsigned long long *__restrict__ counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SynthSmem &s = *reinterpret_cast<SynthSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 15, half = lane >> 4;
    const unsigned half_mask = 0xffffu << (16 * half);
    if (threadIdx.x < 128) s.tw[threadIdx.x] = __ldg(T + threadIdx.x);
    __syncthreads();
    const auto lds = [](const float4 *q) { return *q; };
    const auto ldg = [](const float4 *q) { return __ldg(q); };
    // staging: lane -> (sample lane/4 + 8 i, quarter lane%4 of the eight bands), i = 0..31; the words of a quarter are
    // 16 bytes, consecutive i are 8 samples (8 * bands words) apart
    const int smp0 = lane >> 2, q = lane & 3;
    const long long src_step = 8LL * bands;
    float *const dst0 = &s.stage[warp][2 * q][smp0];
    const int total_warps = gridDim.x * kWarps;

    for (int work = blockIdx.x * kWarps + warp; work < n_work;) {
        int next = 0;
        if (lane == 0) next = total_warps + (int)atomicAdd(counter, 1ull);  // used after the first group is staged
        const int item = work / n_seg, p = work - item * n_seg;
        const BandSynthItem it = batch.items[item];
        const long long *__restrict__ hist = it.hist;
        const long long first = (long long)p * 256;
        const int n_valid = (int)min((long long)256, (long long)bins - first);  // samples of this segment that exist
        const long long *src0 = hist + (first + smp0) * bands + 2 * q;
        const float qs = kQ * it.scale;  // 2^-40 * scale is exact, so (w * 2^-40) * scale == w * qs bit for bit

        f2 U[16], W[16];
        if (!ONE) {
#pragma unroll
            for (int k = 0; k < 16; k++) U[k] = W[k] = f2{0.f, 0.f};
        }

        for (int c0 = 0; c0 < (ONE ? 1 : bands); c0 += kStage) {
            // stage bands [c0, c0+8) of the segment: 256 samples x 64 bytes = 1024 16-byte loads, 32 per lane
            const long long *src = src0 + c0;
            constexpr int kB = ONE ? 2 * kBatch : kBatch;
            if (n_valid == 256 && c0 + kStage <= bands) {  // (warp-uniform) everything exists: no predicates
#pragma unroll 1
                for (int i0 = 0; i0 < 32; i0 += kB) {
                    longlong2 w[kB];
#pragma unroll
                    for (int j = 0; j < kB; j++, src += src_step) w[j] = ld_stream16(src);
#pragma unroll
                    for (int j = 0; j < kB; j++) {
                        float *dst = dst0 + 8 * (i0 + j);
                        dst[0] = (float)w[j].x * qs;
                        dst[kRow] = (float)w[j].y * qs;
                    }
                }
            } else {
                const int lim = c0 + 2 * q < bands ? (n_valid - smp0 + 7) >> 3 : 0;  // this lane's loads i < lim exist
#pragma unroll 1
                for (int i0 = 0; i0 < 32; i0 += kBatch) {
                    longlong2 w[kBatch];
#pragma unroll
                    for (int j = 0; j < kBatch; j++)
                        w[j] = i0 + j < lim ? ld_stream16(src + (long long)(i0 + j) * src_step) : make_longlong2(0, 0);
#pragma unroll
                    for (int j = 0; j < kBatch; j++) {
                        float *dst = dst0 + 8 * (i0 + j);
                        dst[0] = (float)w[j].x * qs;
                        dst[kRow] = (float)w[j].y * qs;
                    }
                }
            }
            __syncwarp();
            if (c0 == 0) next = __shfl_sync(0xffffffffu, next, 0);
            if (ONE && bands == kStage) {
                // L2 prefetch of the next pair's words: with eight bands a segment is 16 KB of contiguous memory, one
                // bulk prefetch issued by one lane
                if (next < n_work && lane == 0) {
                    const int nitem = next / n_seg, np = next - nitem * n_seg;
                    const long long nfirst = (long long)np * 256;
                    const int rows = (int)min((long long)256, (long long)bins - nfirst);
                    prefetch_l2_bulk(batch.items[nitem].hist + nfirst * kStage, (unsigned)rows * 64u);
                }
            } else {
                // ... in general: the next group of this segment, or group 0 of the next pair, row by row
                const bool more = c0 + kStage < bands;
                const int nitem = more ? item : next / n_seg, np = more ? p : next - nitem * n_seg;
                if (more || next < n_work) {
                    const long long *nh = batch.items[nitem].hist;
                    const long long nfirst = (long long)np * 256;
                    const int rows = (int)min((long long)256, (long long)bins - nfirst);
                    const int row_bytes = 8 * min(kStage, bands - (more ? c0 + kStage : 0));
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int r = lane + 32 * i;
                        if (r < rows) {
                            const char *a = reinterpret_cast<const char *>(nh + (nfirst + r) * bands + (more ? c0 + kStage : 0));
                            prefetch_l2(a);
                            if (row_bytes > 32 && ((reinterpret_cast<uintptr_t>(a) ^ reinterpret_cast<uintptr_t>(a + row_bytes - 1)) >> 7)) prefetch_l2(a + row_bytes - 1);
                        }
                    }
                }
            }
            if (ONE) {
#pragma unroll
                for (int k = 0; k < 16; k++) U[k] = W[k] = f2{0.f, 0.f};
            }
            if (c0 + 4 * half < bands) {  // (a band count of 4 mod 8 leaves the upper half without bands in the last group)
#pragma unroll 1
                for (int b = 0; b < 4; b++) {
                    f2 y[16];
                    const f2 *row = reinterpret_cast<const f2 *>(&s.stage[warp][4 * half + b][0]);
#pragma unroll
                    for (int r = 0; r < 8; r++) y[r] = row[t + 16 * r];
                    fft16<false, true>(y);
                    {
                        f2 tw[16];
                        load_row16(s.tw, t, tw, lds);
#pragma unroll
                        for (int k = 1; k < 16; k++) y[k] = cmul(y[k], tw[k]);
                    }
                    transpose16(s.tr[warp][half], t, y, half_mask);
                    fft16<false, false>(y);
                    f2 ab[16];
                    load_row16(T + 256 + (size_t)(c0 + 4 * half + b) * 128, t, ab, ldg);
                    synth_accumulate(U, W, y, ab);
                }
            }
            __syncwarp();  // the staging rows are rewritten by the next group
        }

        // Everything from here on is linear in (U, W): each half warp takes its own partial sums (its four bands of every
        // group) through the merge and the inverse transform, and the two windows are added at the very end.
        // Partner sums at M-k: lane (16 - t) & 15 of the same half warp, register 15 - k2 (own register (16 - k2) & 15
        // for t = 0).
        f2 Zp[16];
        {
            const int src_lane = (lane & 16) | ((16 - t) & 15);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 w2 = __ldg(T + 128 + j * 16 + t);  // elements 2j, 2j+1 of this thread's row of w2
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int k2 = 2 * j + h;
                    const f2 e = csub(W[15 - k2], U[15 - k2]);
                    f2 ep = f2{__shfl_sync(0xffffffffu, e.x, src_lane), __shfl_sync(0xffffffffu, e.y, src_lane)};
                    if (t == 0) ep = csub(W[(16 - k2) & 15], U[(16 - k2) & 15]);
                    Zp[k2] = synth_merge1(t == 0 && k2 == 0, U[k2], W[k2], ep, h ? f2{w2.z, w2.w} : f2{w2.x, w2.y});
                }
            }
        }
        fft16<true, false>(Zp);
        {
            f2 tw[16];
            load_row16(s.tw, t, tw, lds);
#pragma unroll
            for (int k = 1; k < 16; k++) Zp[k] = cmul(Zp[k], conj2(tw[k]));
        }
        transpose16(s.tr[warp][half], t, Zp, half_mask);
        fft16<true, false>(Zp);
        // Zp[n1] = 256 x (window[2n], window[2n+1]), n = t + 16 n1; window index m >= 384 is time m - 512.  o is even
        // and `out` 8-byte aligned (launcher), so a pair goes out as one float2 atomic (element-wise atomic, sm_90+).
#pragma unroll
        for (int k = 0; k < 16; k++) {  // this half's window + the other half's
            Zp[k].x += __shfl_xor_sync(0xffffffffu, Zp[k].x, 16);
            Zp[k].y += __shfl_xor_sync(0xffffffffu, Zp[k].y, 16);
        }
        float *__restrict__ out = it.out;
        const float inv = 1.0f / 256.0f;
        const int base = (int)first + 2 * t;  // sample counts fit in 31 bits (check_convolvable)
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            if ((n1 >> 3) != half) continue;
            const int o = base + (n1 < 12 ? 32 * n1 : 32 * n1 - 512);
            const float vx = Zp[n1].x * inv, vy = Zp[n1].y * inv;
            if (o >= 0 && o + 1 < out_len) {
                atomicAdd(reinterpret_cast<float2 *>(out + o), make_float2(vx, vy));
            } else {
                if (o >= 0 && o < out_len) atomicAdd(out + o, vx);
                if (o + 1 >= 0 && o + 1 < out_len) atomicAdd(out + o + 1, vy);
            }
        }
        work = next;
    }
}

// ---- partition spectra with the same register transforms ------------------------------------------------------------
// H[p] = rfft512([ir[256 p .. 256 p + 256), 0 x 256]) as a packed half spectrum (rar_fft.cuh), one partition per half
// warp: the 8-byte loads of a half warp are 128 contiguous bytes, and so are its stores (bins t + 16 k2 over t).
// Replaces the four-pass shared-memory ir_spectra_kernel of conv_kernels.cu on the paths that load responses
// (rar_conv_set_ir*, the per-slot spectra cache, the banded step after the synthesis).
__device__ float4 g_syn_tw[256];  // synth_tables(nullptr, 0, .): tw (128 float4), w2 (128 float4)

__global__ void __launch_bounds__(128) ir_spectra16_kernel(const float *__restrict__ ir, long long ir_stride, int ir_len, f2 *__restrict__ H,
                                                           long long h_stride, int n_part) {
    __shared__ __align__(16) f2 tr[8][16][kTRow];
    __shared__ float4 tw_s[128];
    const int lane = threadIdx.x & 31, t = lane & 15, g = threadIdx.x >> 4;
    const unsigned half_mask = 0xffffu << (lane & 16);
    tw_s[threadIdx.x] = g_syn_tw[threadIdx.x];
    __syncthreads();
    if (blockIdx.x * 8 + (g & ~1) >= n_part) return;  // (whole warp)
    const int p = blockIdx.x * 8 + g;
    const bool live = p < n_part;
    ir += (long long)blockIdx.y * ir_stride;
    H += (long long)blockIdx.y * h_stride;
    const long long first = (long long)p * 256;
    f2 y[16];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const long long n = first + 2 * (t + 16 * r);
        if (live && n + 1 < ir_len) {
            const float2 v = __ldcs(reinterpret_cast<const float2 *>(ir + n));
            y[r] = f2{v.x, v.y};
        } else {
            y[r] = f2{live && n < ir_len ? ir[n] : 0.0f, 0.0f};
        }
    }
    fft16<false, true>(y);
    {
        f2 tw[16];
        load_row16(tw_s, t, tw, [](const float4 *q) { return *q; });
#pragma unroll
        for (int k = 1; k < 16; k++) y[k] = cmul(y[k], tw[k]);
    }
    transpose16(tr[g], t, y, half_mask);
    fft16<false, false>(y);  // y[k2] = Z[t + 16 k2]
    const int src_lane = (lane & 16) | ((16 - t) & 15);
    f2 *__restrict__ dst = H + (size_t)p * 256 + t;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float4 w2 = g_syn_tw[128 + j * 16 + t];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int k2 = 2 * j + h;
            const f2 z = y[15 - k2];
            f2 zm = f2{__shfl_sync(0xffffffffu, z.x, src_lane), __shfl_sync(0xffffffffu, z.y, src_lane)};
            if (t == 0) zm = y[(16 - k2) & 15];
            f2 P = synth_split(y[k2], zm, h ? f2{w2.z, w2.w} : f2{w2.x, w2.y});
            if (t == 0 && k2 == 0) P = f2{y[0].x + y[0].y, y[0].x - y[0].y};
            if (live) __stcs(reinterpret_cast<float2 *>(dst + 16 * k2), make_float2(P.x, P.y));
        }
    }
}

}  // namespace

bool band_synth16_applicable(int bands, int stride) { return bands > 0 && bands % 4 == 0 && stride == 1; }

cudaError_t launch_band_synth16(const BandSynthBatch &batch, int n_items, int bins, int bands, const float2 *T, int out_len,
                                unsigned long long *counter, int sm_count, cudaStream_t s) {
    static std::mutex mu;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::unique_lock<std::mutex> lock(mu);  // contexts of several devices may be driven from several threads
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(band_synth16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SynthSmem));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(band_synth16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SynthSmem));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    lock.unlock();
    for (int k = 0; k < n_items; k++)
        if ((reinterpret_cast<uintptr_t>(batch.items[k].out) & 7u) || (reinterpret_cast<uintptr_t>(batch.items[k].hist) & 15u))
            return cudaErrorMisalignedAddress;
    const int n_seg = (bins + 255) / 256;
    const long long n_work = (long long)n_seg * n_items;
    if (n_work > 0x7fffffffLL / 2) return cudaErrorInvalidValue;
    const int grid = (int)std::min<long long>((n_work + kWarps - 1) / kWarps, 4LL * (sm_count > 0 ? sm_count : 148));
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (bands <= kStage)
        band_synth16_kernel<true><<<grid, 32 * kWarps, sizeof(SynthSmem), s>>>(batch, bins, bands, reinterpret_cast<const float4 *>(T), out_len, n_seg,
                                                                                (int)n_work, counter);
    else
        band_synth16_kernel<false><<<grid, 32 * kWarps, sizeof(SynthSmem), s>>>(batch, bins, bands, reinterpret_cast<const float4 *>(T), out_len, n_seg,
                                                                                 (int)n_work, counter);
    return cudaGetLastError();
}

}  // namespace rar

namespace rar {

void synth_init_tables(cudaStream_t stream) {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return;
    f2 tab[512];
    synth_tables(nullptr, 0, tab);
    cudaMemcpyToSymbolAsync(g_syn_tw, tab, sizeof tab, 0, cudaMemcpyHostToDevice, stream);
    cudaStreamSynchronize(stream);  // the host array is on this stack frame
    if (dev >= 0 && dev < 64) done[dev] = true;
}

bool ir_spectra16_applicable(const float *ir_f, long long ir_stride) {
    return (reinterpret_cast<uintptr_t>(ir_f) & 7u) == 0 && (ir_stride & 1) == 0;
}

cudaError_t launch_ir_spectra16(const float *ir_f, long long ir_stride, int ir_len, float2 *H, long long h_stride, int n_part, int n_items,
                                cudaStream_t s) {
    if (n_part <= 0 || n_items <= 0) return cudaSuccess;
    ir_spectra16_kernel<<<dim3((n_part + 7) / 8, n_items), 128, 0, s>>>(ir_f, ir_stride, ir_len, reinterpret_cast<f2 *>(H), h_stride, n_part);
    return cudaGetLastError();
}

}  // namespace rar

// band_synth.cu -- filter-bank synthesis of a broadband response from a banded slot, production kernel
// for slots whose band count is a multiple of 4 on the sample grid (time stride 1).  What it computes is stated in
// conv_kernels.cu (band_synth_kernel, the generic kernel that remains for the other shapes):
//        out[n] = sum_b (g_b * h_b)[n + 127],   h_b[n] = IR[n * bands + b] * 2^-40 * scale,
// per 256-sample segment as a 512-point circular convolution in the frequency domain, overlap-added into `out`.
//
// Layout of the work.  16 threads (half a warp) own one segment and hold every transform in registers
// (rar_synth16.cuh: 16 x 16 split, one transposition through shared memory per transform, real weights, the partner
// exchange by shuffles once per segment).  A warp stages the 64-bit histogram words of its two segments through
// shared memory four bands at a time: each lane issues 16-byte loads of (sample, 2 bands), a chunk of four bands is
// one full 32-byte sector per sample, the words are converted once and stored band-major so that a transform's
// inputs are conflict-free 8-byte shared loads.  Nothing in the main loop needs a block-wide barrier: the 16
// threads of a transform live in one warp (__syncwarp).  Twiddles and weights come from a small table in global
// memory (2 KB + 2 KB + 2 KB per band) through L1.
//
// HBM roofline: 8 bytes per (sample, band) read once, 4 bytes per sample added into the response (twice, by the two
// windows that cover it; the second add hits L2).
#include <cuda_runtime.h>

#include "rar_internal.h"
#include "rar_synth16.cuh"

namespace rar {
namespace {

constexpr int kGroups = 8;               // segments per CTA (two per warp)
constexpr int kRow = 264;                // floats per staged band row: 256 samples + 8 (rows 0/2 and 1/3 on different banks)
constexpr int kChunk = 4;                // bands staged at a time
constexpr int kTRow = 18;                // f2 per transposition row: 16 + 2 (144-byte rows: conflict-free 16-byte reads)
constexpr float kQ = 9.094947017729282e-13f;  // 2^-40

struct SynthSmem {
    float stage[kGroups][kChunk][kRow];
    f2 tr[kGroups][16][kTRow];
};

__device__ __forceinline__ void load_row16(const f2 *__restrict__ row, f2 (&v)[16]) {
    const float4 *q = reinterpret_cast<const float4 *>(row);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float4 a = __ldg(q + j);
        v[2 * j] = f2{a.x, a.y};
        v[2 * j + 1] = f2{a.z, a.w};
    }
}

// transposition: this thread's y[k1] goes to row k1, column t; it then takes row t
__device__ __forceinline__ void transpose16(f2 (*tr)[kTRow], int t, f2 (&y)[16]) {
#pragma unroll
    for (int k = 0; k < 16; k++) tr[k][t] = y[k];
    __syncwarp();
    const float4 *q = reinterpret_cast<const float4 *>(&tr[t][0]);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float4 a = q[j];
        y[2 * j] = f2{a.x, a.y};
        y[2 * j + 1] = f2{a.z, a.w};
    }
    __syncwarp();  // the rows are rewritten by the next transform
}

__global__ void __launch_bounds__(32 * kGroups / 2, 4)
band_synth16_kernel(const __grid_constant__ BandSynthBatch batch, int bins, int bands, const f2 *__restrict__ T, int out_len, int n_seg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SynthSmem &s = *reinterpret_cast<SynthSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 15, g = 2 * warp + (lane >> 4);
    const int p0 = blockIdx.x * kGroups + 2 * warp;  // the warp's first segment
    if (p0 >= n_seg) return;
    const int p = p0 + (lane >> 4);
    const BandSynthItem it = batch.items[blockIdx.y];
    const long long *__restrict__ hist = it.hist;
    const float scale = it.scale;
    const f2 *__restrict__ tw_row = T + t * 16, *__restrict__ w2_row = T + 256 + t * 16, *__restrict__ wt = T + 512 + t * 16;

    f2 U[16], W[16];
#pragma unroll
    for (int k = 0; k < 16; k++) U[k] = W[k] = f2{0.f, 0.f};

    for (int c0 = 0; c0 < bands; c0 += kChunk) {
        // stage bands [c0, c0+4) of the warp's two segments: 2 x 256 samples x 32 bytes = 1024 16-byte loads
#pragma unroll 1
        for (int i0 = 0; i0 < 32; i0 += 8) {
            longlong2 w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int idx = (i0 + j) * 32 + lane;  // (segment of the warp, sample, half of the chunk)
                const int sg = idx >> 9, smp = (idx >> 1) & 255, half = idx & 1;
                const long long gs = (long long)(p0 + sg) * 256 + smp;
                const bool on = p0 + sg < n_seg && gs < bins;
                w[j] = on ? __ldcs(reinterpret_cast<const longlong2 *>(hist + gs * bands + c0 + 2 * half)) : make_longlong2(0, 0);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int idx = (i0 + j) * 32 + lane;
                const int sg = idx >> 9, smp = (idx >> 1) & 255, half = idx & 1;
                float *dst = &s.stage[2 * warp + sg][2 * half][smp];
                dst[0] = ((float)w[j].x * kQ) * scale;
                dst[kRow] = ((float)w[j].y * kQ) * scale;
            }
        }
        __syncwarp();
#pragma unroll 1
        for (int b = 0; b < kChunk; b++) {
            f2 y[16];
            const f2 *row = reinterpret_cast<const f2 *>(&s.stage[g][b][0]);
#pragma unroll
            for (int r = 0; r < 8; r++) y[r] = row[t + 16 * r];
            fft16<false, true>(y);
            {
                f2 tw[16];
                load_row16(tw_row, tw);
#pragma unroll
                for (int k = 1; k < 16; k++) y[k] = cmul(y[k], tw[k]);
            }
            transpose16(s.tr[g], t, y);
            fft16<false, false>(y);
            f2 ab[16];
            load_row16(wt + (size_t)(c0 + b) * 256, ab);
            synth_accumulate(U, W, y, ab);
        }
        __syncwarp();  // the staging rows are rewritten by the next chunk
    }

    // partner sums at M-k: lane (16 - t) & 15 of the same half warp, register 15 - k2 (own register (16 - k2) & 15 for t = 0)
    f2 Zp[16];
    {
        f2 Up[16], Wp[16];
        const int src = (lane & 16) | ((16 - t) & 15);
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            const f2 u = U[15 - k2], w = W[15 - k2];
            f2 pu = f2{__shfl_sync(0xffffffffu, u.x, src), __shfl_sync(0xffffffffu, u.y, src)};
            f2 pw = f2{__shfl_sync(0xffffffffu, w.x, src), __shfl_sync(0xffffffffu, w.y, src)};
            if (t == 0) {
                pu = U[(16 - k2) & 15];
                pw = W[(16 - k2) & 15];
            }
            Up[k2] = pu;
            Wp[k2] = pw;
        }
        f2 w2[16];
        load_row16(w2_row, w2);
        synth_merge(t, U, W, Up, Wp, w2, Zp);
    }
    fft16<true, false>(Zp);
    {
        f2 tw[16];
        load_row16(tw_row, tw);
#pragma unroll
        for (int k = 1; k < 16; k++) Zp[k] = cmul(Zp[k], conj2(tw[k]));
    }
    transpose16(s.tr[g], t, Zp);
    fft16<true, false>(Zp);
    if (p < n_seg) {
        // Zp[n1] = 256 x (window[2n], window[2n+1]), n = t + 16 n1; window index m >= 384 is time m - 512
        float *__restrict__ out = it.out;
        const long long first = (long long)p * 256;
        const float inv = 1.0f / 256.0f;
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            const int m = 2 * (t + 16 * n1);
            const long long o = first + (n1 < 12 ? m : m - 512);
            if (o >= 0 && o < out_len) atomicAdd(out + o, Zp[n1].x * inv);
            if (o + 1 >= 0 && o + 1 < out_len) atomicAdd(out + o + 1, Zp[n1].y * inv);
        }
    }
}

}  // namespace

bool band_synth16_applicable(int bands, int stride) { return bands > 0 && bands % kChunk == 0 && stride == 1; }

cudaError_t launch_band_synth16(const BandSynthBatch &batch, int n_items, int bins, int bands, const float2 *T, int out_len,
                                cudaStream_t s) {
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(band_synth16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SynthSmem));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int n_seg = (bins + 255) / 256;
    const dim3 grid((n_seg + kGroups - 1) / kGroups, n_items);
    band_synth16_kernel<<<grid, 32 * kGroups / 2, sizeof(SynthSmem), s>>>(batch, bins, bands, reinterpret_cast<const f2 *>(T), out_len, n_seg);
    return cudaGetLastError();
}

}  // namespace rar

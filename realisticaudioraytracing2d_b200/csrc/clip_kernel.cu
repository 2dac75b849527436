// clip_kernel.cu -- RayTraceManager.cs:135-167 LoadSample for a batch of clips on the GPU: mono mix
// (channels summed in order, divided by the channel count) followed by the reference's linear resampling,
//     srcIdx = i * ratio;  idx0 = floor(srcIdx);  idx1 = min(idx0 + 1, samples - 1);
//     out[i] = Mathf.Lerp(mono[idx0], mono[idx1], srcIdx - idx0) = a + (b - a) * clamp01(t).
// The reference runs this in C# on the CPU, so its arithmetic is plain IEEE binary32 with no contraction; the kernel
// uses the explicitly rounded intrinsics (and the translation unit is built with --fmad=false), which makes the
// result bit-identical to oracle/rar_oracle.c:orc_load_sample.  HBM-bound: the algorithmic traffic is one read of
// the raw clip plus one write of the result.
#include <cuda_runtime.h>

#include "rar_internal.h"

namespace rar {

namespace {

// The raw channels of one source sample (CH 1/2: one vector load) and their mix.  sum / channels is an exact
// scaling for one and two channels (x/1 = x, x/2 = x*0.5 in every binade, subnormals included), so only the
// generic path pays for the IEEE division.
template <int CH> struct RawSample { float2 v; };

template <int CH>
__device__ __forceinline__ float2 raw_at(const float *__restrict__ clip, int idx) {
    if (CH == 1) return make_float2(__ldg(clip + idx), 0.0f);
    return __ldg(reinterpret_cast<const float2 *>(clip) + idx);
}

template <int CH>
__device__ __forceinline__ float mix(float2 v) {
    if (CH == 1) return __fadd_rn(0.0f, v.x);
    return __fmul_rn(__fadd_rn(__fadd_rn(0.0f, v.x), v.y), 0.5f);
}

template <int CH>
__device__ __forceinline__ float mono_at(const float *__restrict__ clip, int idx, int channels) {
    if (CH == 1 || CH == 2) return mix<CH>(raw_at<CH>(clip, idx));
    float sum = 0.0f;
    const float *p = clip + (size_t)idx * channels;
    for (int c = 0; c < channels; c++) sum = __fadd_rn(sum, __ldg(p + c));
    return __fdiv_rn(sum, (float)channels);
}

constexpr int kSpanCap = 4096;   // mono samples a block stages per tile (16 KB of shared memory)
constexpr int kTileMax = 2048;   // output samples per tile

// int -> float.  SMALL: the value is below 2^23, where 2^23 + i is exact, so two ALU instructions replace the
// conversion (which issues at a quarter of the rate); the result is the same float.
template <bool SMALL>
__device__ __forceinline__ float to_float(int i) {
    return SMALL ? __fsub_rn(__int_as_float(0x4B000000 | i), 8388608.0f) : __int2float_rn(i);
}

template <bool SMALL>
__device__ __forceinline__ int src_index(const ClipPrep &a, int i, float *src_out) {
    const float src = __fmul_rn(to_float<SMALL>(i), a.ratio);
    int idx0 = __float2int_rd(fminf(src, 2147483520.0f));  // floor
    if (idx0 > (int)a.samples - 1) idx0 = (int)a.samples - 1;  // the C# would throw; never reached for sane ratios
    *src_out = src;
    return idx0;
}

template <bool SMALL>
__device__ __forceinline__ float lerp_clamped(float x0, float x1, float src, int idx0) {
    float t = __fsub_rn(src, to_float<SMALL>(idx0));
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
    return __fadd_rn(x0, __fmul_rn(__fsub_rn(x1, x0), t));
}

// One block per tile of a.tile consecutive output samples of one clip: the mono mix of the source span the tile
// touches is computed once into shared memory with coalesced loads (each source sample is mixed once, not once per
// output that uses it; eight loads per thread are issued before the first is consumed), then every output
// interpolates from there.  Sample counts fit an int (AudioClip.samples and newLength are C# ints).
template <int CH, bool SMALL>
__global__ void __launch_bounds__(256) prepare_clips_kernel(const ClipPrep a) {
    __shared__ float s_mono[kSpanCap];
    const int new_len = (int)a.new_len, samples = (int)a.samples;
    const int tiles_per_clip = (new_len + a.tile - 1) / a.tile;
    const long long jobs = (long long)tiles_per_clip * a.n_clips;
    for (long long job = blockIdx.x; job < jobs; job += gridDim.x) {
        const int clip = (int)(job / tiles_per_clip), tile = (int)(job - (long long)clip * tiles_per_clip);
        const float *__restrict__ raw = a.raw + (size_t)clip * a.samples * a.channels;
        float *__restrict__ out = a.out + (size_t)clip * a.out_stride;
        const int i0 = tile * a.tile, i1 = min(i0 + a.tile, new_len);
        if (!a.resample) {
            for (int i = i0 + threadIdx.x; i < i1; i += 256) out[i] = mono_at<CH>(raw, i, a.channels);
            continue;
        }
        float f;
        const int s_lo = src_index<SMALL>(a, i0, &f);
        const int s_hi = min(src_index<SMALL>(a, i1 - 1, &f) + 1, samples - 1);
        const int n_src = s_hi - s_lo + 1;
        const bool staged = n_src <= kSpanCap;  // block-uniform; false only for extreme ratios
        if (staged) {
            if (CH == 1 || CH == 2) {
                const float *__restrict__ span = raw + (size_t)s_lo * CH;
                for (int base = 0; base < n_src; base += 8 * 256) {
                    float2 v[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int k = base + j * 256 + threadIdx.x;
                        v[j] = make_float2(0.0f, 0.0f);
                        if (k < n_src) v[j] = raw_at<CH>(span, k);
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int k = base + j * 256 + threadIdx.x;
                        if (k < n_src) s_mono[k] = mix<CH>(v[j]);
                    }
                }
            } else {
                for (int k = threadIdx.x; k < n_src; k += 256) s_mono[k] = mono_at<CH>(raw, s_lo + k, a.channels);
            }
        }
        __syncthreads();
        float *__restrict__ o = out + i0;
        const int n_out = i1 - i0;
        if (staged) {
#pragma unroll 8
            for (int j = threadIdx.x; j < n_out; j += 256) {
                float src;
                const int idx0 = src_index<SMALL>(a, i0 + j, &src);
                const int idx1 = min(idx0 + 1, samples - 1);
                o[j] = lerp_clamped<SMALL>(s_mono[idx0 - s_lo], s_mono[idx1 - s_lo], src, idx0);
            }
        } else {
            for (int j = threadIdx.x; j < n_out; j += 256) {
                float src;
                const int idx0 = src_index<SMALL>(a, i0 + j, &src);
                const int idx1 = min(idx0 + 1, samples - 1);
                o[j] = lerp_clamped<SMALL>(mono_at<CH>(raw, idx0, a.channels), mono_at<CH>(raw, idx1, a.channels), src, idx0);
            }
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t launch_prepare_clips(const ClipPrep &a_in, cudaStream_t s, int sm_count) {
    if (a_in.n_clips <= 0 || a_in.new_len <= 0) return cudaSuccess;
    ClipPrep a = a_in;
    // tile of outputs whose source span (tile * ratio + slack) fits the staging buffer
    double t = a.resample ? (double)(kSpanCap - 8) / (double)a.ratio : (double)kTileMax;
    a.tile = t >= kTileMax ? kTileMax : (t < 1.0 ? 1 : (int)t);
    const long long jobs = ((a.new_len + a.tile - 1) / a.tile) * a.n_clips;
    const long long cap = (long long)sm_count * 32;
    const unsigned grid = (unsigned)(jobs < cap ? jobs : cap);
    const bool small = a.new_len <= (1 << 23) && a.samples <= (1 << 23);
    if (a.channels == 1) {
        if (small) prepare_clips_kernel<1, true><<<grid, 256, 0, s>>>(a);
        else prepare_clips_kernel<1, false><<<grid, 256, 0, s>>>(a);
    } else if (a.channels == 2) {
        if (small) prepare_clips_kernel<2, true><<<grid, 256, 0, s>>>(a);
        else prepare_clips_kernel<2, false><<<grid, 256, 0, s>>>(a);
    } else {
        if (small) prepare_clips_kernel<0, true><<<grid, 256, 0, s>>>(a);
        else prepare_clips_kernel<0, false><<<grid, 256, 0, s>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace rar

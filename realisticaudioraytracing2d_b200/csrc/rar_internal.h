// rar_internal.h -- launch descriptors shared by the kernel translation units and the C-ABI layer.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rar2d.h"
#include "rar_ray.cuh"

namespace rar {

// Everything one launch of the trace kernel needs.  Passed by value as a __grid_constant__.
struct TraceLaunch {
    const f4 *geo;          // [n_walls] endpoint plane (TMA bulk-copy source, 16-byte records)
    const f4 *mat0;         // [n_walls] normal + absorption + scattering
    const f2 *mat1;         // [n_walls] transmission + ior
    const float *band_abs;  // [n_walls][bands] or nullptr
    int n_walls;
    int bands;        // bands the kernel instantiation carries: 1, or 8 (one chunk of a banded slot)
    int band_total;   // bands of the slot (row stride of hist and of band_abs)
    int band_offset;  // first band of this chunk
    int band_valid;   // bands of this chunk that exist (<= 8)
    RayConsts p;
    long long ray_begin, ray_end;
    int n_frames;  // >= 1: frames rng_state_offset .. +n_frames-1 of the same dispatch traced by one launch
    // block-cyclic sharding (rar_trace_interleaved): cyc_world > 1 maps launch index i to dispatch thread id
    // ((i >> shift) * world + rank) << shift | (i & mask); ids >= cyc_total do not exist.  [ray_begin, ray_end) is then
    // the launch's own index range, starting at 0.
    int cyc_world, cyc_rank, cyc_shift;
    long long cyc_total;
    unsigned long long *hist;  // [impulse_length][bands] Q23.40, nullptr in hit-list mode
    rar_ray_info *hits;        // hit-list mode outputs
    rar_hit_key *keys;
    long long hit_cap;
    unsigned long long *hit_count;
    unsigned long long *counters;  // 5 words (rar_counters) or nullptr
    unsigned long long *tile_counter;  // one word the launch code zeroes: CTAs claim their further ray tiles from it
    f4 *debug_rays;                // max(100, debug_ray_count) * (max_bounce_count+1) or nullptr
    int debug_ray_count;
    int debug_capacity;            // entries in debug_rays
    // batched listeners (BASELINE config 4): n_listeners > 0 selects the fused kernel; listener l deposits
    // into listener_hists[l]
    const f2 *listeners;
    unsigned long long *const *listener_hists;
    int n_listeners;
    // optional uniform grid (RAR_FLAG_USE_GRID); use_grid selects the grid instantiation
    GridView grid;
    int use_grid;
    int opaque;  // no wall has transmission > 0: kernels without the transmit/refract branch may be used
    const f4 *pair_a, *pair_b;  // the endpoint plane as two-wall records (rar_layout.h pair_planes), for the PACKED kernels
    int spec_ok; // opaque, and every operand range the range-checked-once arithmetic assumes holds (spec_ranges_ok)
};

struct DeviceFacts {
    int sm_count;
    int smem_optin;  // bytes of dynamic shared memory a block may opt in to
    int sm_clock_khz;
};

// trace_kernel.cu
cudaError_t launch_trace(const TraceLaunch &a, bool count_tests, const DeviceFacts &dev, cudaStream_t stream,
                         int *launches);
cudaError_t launch_fp32_peak(float *d_sink, int blocks, int threads, int iters, cudaStream_t stream);
cudaError_t launch_arithmetic_selftest(long long n, uint32_t seed, unsigned long long *d_mism, int blocks, cudaStream_t stream);

// conv_kernels.cu -- all spectra are "packed half spectra" of a real FFT of size 2*B: B complex values
// per block, bin 0 holding (DC, Nyquist) in (re, im).
struct ConvPlan {
    int block;       // B (256)
    int log2_block;  // 8
};

// hist (Q23.40 int64) -> float, optionally scaled; n values.
// hist[i] += sum over the peers' histograms (device pointers readable from the current device), i < n
struct PeerHists { const long long *p[16]; int n; };
cudaError_t launch_peer_reduce(long long *hist, PeerHists peers, long long n, cudaStream_t s);

// exchange_kernel.cu -- multi-process all-reduce of a histogram over CUDA-IPC-mapped peer memory
constexpr int kExMaxRanks = 16;
constexpr int kExMaxBlocks = 296;  // two per SM
constexpr size_t kExFlagWords = 2 * (size_t)kExMaxBlocks * kExMaxRanks;  // u32 words of the flag table
struct ExchangeLaunch {
    long long *hist;        // this rank's slot
    long long n_vecs;       // 16-byte vectors covering the slot's words
    long long slice_vecs;   // vectors per rank slice (two-shot), world * slice_vecs >= n_vecs
    int rank, world;
    unsigned epoch;         // number of this call, >= 1, the same on every rank
    int two_shot;
    unsigned *flags[kExMaxRanks];       // every rank's flag table (own entry: local address)
    long long *stage_in[kExMaxRanks];   // every rank's staging buffer of this call's parity
    long long *stage_out[kExMaxRanks];  // every rank's totals buffer of this call's parity
    unsigned *status;                   // own status word
    unsigned long long timeout_ns;
};
cudaError_t launch_exchange_allreduce(const ExchangeLaunch &x, cudaStream_t stream);

cudaError_t launch_fixed_to_float(const long long *hist, float *out, long long n, float scale, cudaStream_t s);
cudaError_t launch_float_to_fixed(const float *in, long long *hist, long long n, cudaStream_t s);

// Spectra of the IR partitions: H[p] = rfft([ir[p*B .. p*B+B), 0...0]) for p in [0, n_part).
// ir_f (float, already scaled) has ir_len valid samples.
cudaError_t launch_ir_spectra(const float *ir_f, int ir_len, float2 *H, int n_part, int block, cudaStream_t s);
// The same for n_items responses: response k at ir_f + k*ir_stride (floats), its spectra at H + k*h_stride (float2).
cudaError_t launch_ir_spectra_batch(const float *ir_f, long long ir_stride, int ir_len, float2 *H, long long h_stride, int n_part,
                                    int n_items, int block, cudaStream_t s);

// Filter bank of the banded model: 255-tap windowed-sinc band-pass filters, linear phase.
constexpr int kBandFilterTaps = 255;
constexpr int kBandFilterDelay = 127;
// Taps of band [lo, hi) (edges as fractions of the Nyquist frequency), g[kBandFilterTaps].
void band_filter_taps(double lo, double hi, float *g);
// out[n] += sum_b (g_b * h_b)[n + delay], h_b[n] = hist[(n/stride)*bands + b] * 2^-40 * scale when stride divides n; G =
// packed half spectra of the band filters [bands][256].  A batch of n_items slots of equal shape, descriptors in
// the launch arguments; every item's out (out_len floats) must have been zeroed.
struct BandSynthItem {
    const long long *hist;
    float *out;
    float scale;
    int pad;
};
constexpr int kBandSynthBatch = 32;
struct BandSynthBatch {
    BandSynthItem items[kBandSynthBatch];
};
cudaError_t launch_band_synth(const BandSynthBatch &batch, int n_items, int bins, int bands, int stride, const float2 *G, int out_len,
                              cudaStream_t s);
// The production kernel for bands % 4 == 0 on the sample grid (band_synth.cu): same result within the 1e-4 contract.
// T = the tables of rar_synth16.cuh synth_tables (twiddles and the zero-phase amplitudes of the band filters).
bool band_synth16_applicable(int bands, int stride);
// Partition spectra with the same register transforms (ir 8-byte aligned, even stride); synth_init_tables once per device.
void synth_init_tables(cudaStream_t stream);
bool ir_spectra16_applicable(const float *ir_f, long long ir_stride);
cudaError_t launch_ir_spectra16(const float *ir_f, long long ir_stride, int ir_len, float2 *H, long long h_stride, int n_part, int n_items,
                                cudaStream_t s);
// counter: 8 bytes of device memory the launch may use (work distribution; zeroed by the launcher on the stream).
cudaError_t launch_band_synth16(const BandSynthBatch &batch, int n_items, int bins, int bands, const float2 *T, int out_len,
                                unsigned long long *counter, int sm_count, cudaStream_t s);

// One-shot convolution (AudioConvolve semantics):
//  X[j] = rfft([x[(j-1)B .. jB), x[jB .. (j+1)B)]) with |x| <= 1e-4 zeroed, j in [0, n_xwin)
cudaError_t launch_input_spectra(const float *x, int x_len, float2 *X, int n_xwin, int block, cudaStream_t s);
//  Y[j] = sum_p X[j-p] * H[p], j in [0, n_out_blocks)
cudaError_t launch_block_cmac(const float2 *X, int n_xwin, const float2 *H, int n_part, float2 *Y, int n_out_blocks,
                              int block, cudaStream_t s);
//  out[jB + i] = irfft(Y[j])[B + i] * scale for jB+i < out_len; the rest of out (if any) is zeroed
cudaError_t launch_output_blocks(const float2 *Y, int n_out_blocks, float *out, int out_len, float scale, int block,
                                 cudaStream_t s);

// Streaming convolver (one block per stream per call).
struct StreamConv {
    int n_streams, block, n_part, n_split, part_per_split;
    float2 *H;        // [S][P][B]
    float2 *fdl;      // [S][P][B] ring of input-window spectra
    float2 *partial;  // [S][n_split][B]
    float *prev;      // [S][B] previous input block (first half of the overlap-save window)
    int head;         // ring slot the next window is written to
};
cudaError_t launch_stream_step(const StreamConv &c, const float *d_in, float *d_out, cudaStream_t s, int *launches);
// The same step while the n_list streams in d_list cross-fade to the spectra in H2 (fade[s] != 0 marks them).
cudaError_t launch_stream_step_fade(const StreamConv &c, const float2 *H2, float2 *partial2, const int *d_fade, const int *d_list,
                                    int n_list, const float *d_in, float *d_out, cudaStream_t s, int *launches);

// clip_kernel.cu -- LoadSample (mono mix + linear resample) for a batch of equally shaped clips
struct ClipPrep {
    const float *raw;      // [n_clips][samples][channels]
    float *out;            // [n_clips][out_stride], new_len valid values per clip
    long long samples, new_len, out_stride;
    int channels, n_clips;
    int resample;          // 0: clip frequency == sample rate (mono mix only)
    float ratio;           // (float)clip_frequency / sample_rate
    int tile;              // outputs per block tile (set by the launcher)
};
cudaError_t launch_prepare_clips(const ClipPrep &a, cudaStream_t s, int sm_count);

void conv_init_tables(cudaStream_t s);  // uploads the twiddle tables of the current device on stream s (idempotent per device)

// grid_kernel.cu -- device build of the uniform grid's lists (count, scan, fill, finish); cnt holds nx*ny words
struct GridFrame;
cudaError_t launch_grid_build(const f4 *geo, const f2 *end, int n, const GridFrame &fr, unsigned *cnt, unsigned *cell_start,
                              unsigned *items, f4 *item_geo, cudaStream_t s);

}  // namespace rar

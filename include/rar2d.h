/*
 * rar2d.h -- C-ABI of the B200-native hot path of RealisticAudioRaytracing2D.
 *
 * The reference has no FFI: its boundary between the C# components and the GPU kernels is Unity's
 * string-keyed compute API (ComputeShader.SetInt/SetFloat/SetVector/SetBuffer/Dispatch,
 * ComputeBuffer.SetData/GetData/SetCounterValue/CopyCount/Release, AsyncGPUReadback.Request).
 * Each entry point below replaces one group of those call sites; the call site is cited as
 * file:line relative to /root/reference/Assets/Script/.  A C# host binds these with [DllImport]
 * (see INTEGRATION.md and csharp/RarNative.cs); tests bind them with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host array and the library copies during
 *     the call (the copy semantics of ComputeBuffer.SetData/GetData); device memory belongs to the
 *     context;
 *   - every function returns an int status, RAR_OK (0) on success, negative on failure, never throws;
 *     rar_last_error() gives the message (the reference's convention is "null-guard and silently
 *     skip", RayTraceManager.cs:52,119,222 -- a status code is the closest C equivalent);
 *   - one caller thread per context (all reference GPU calls happen on the Unity main thread);
 *   - trace and convolve_begin are asynchronous with respect to the host: they enqueue on the
 *     context's CUDA stream and return; completion is polled (rar_poll), the analogue of
 *     `while (!req.done) yield return null` (RayTraceManager.cs:115);
 *   - there is no CPU fallback: every compute entry point fails with RAR_ERR_CUDA when no
 *     sm_100-class device is usable.
 */
#ifndef RAR2D_H
#define RAR2D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define RAR_API __declspec(dllexport)
#else
#define RAR_API __attribute__((visibility("default")))
#endif

#define RAR_VERSION 100 /* 0.1.0 */

enum {
    RAR_OK = 0,
    RAR_ERR_INVALID = -1,   /* bad argument */
    RAR_ERR_CUDA = -2,      /* CUDA runtime failure / no usable device */
    RAR_ERR_STATE = -3,     /* call out of order (no walls, slot not configured, ...) */
    RAR_ERR_NOMEM = -4,
    RAR_ERR_UNSUPPORTED = -5,
    RAR_ERR_PENDING = -6    /* ticket not complete yet (convolve_end before done) */
};

/* Helpers/SceneHelper.cs:15-22 `Segment` (LayoutKind.Sequential) == Raytrace2D.compute:19-22 `Wall`
 * with the nested AudioMat (Helpers/SceneHelper.cs:8-14 == Raytrace2D.compute:12-17).  40 bytes. */
typedef struct rar_segment {
    float start[2];
    float end[2];
    float normal[2];
    float absorption;
    float scattering;
    float transmission;
    float ior;
} rar_segment;

/* RayTraceManager.cs:43 / Raytrace2D.compute:24-28 `RayInfo`.  16 bytes. */
typedef struct rar_ray_info {
    float time_delay;
    float energy;
    float hit_point[2];
} rar_ray_info;

/* Which ray produced a rar_ray_info (parity tooling; the reference's append buffer is unordered). */
typedef struct rar_hit_key {
    uint32_t ray;
    uint16_t bounce;
    uint16_t kind; /* 0 direct listener crossing (Raytrace2D.compute:74-84), 1 next-event estimate (:101-119) */
} rar_hit_key;

#define RAR_FLAG_EXACT_RAY_COUNT 1u /* trace exactly ray_count rays; default reproduces the reference's
                                       unguarded dispatch of ceil(rayCount/64)*64 threads
                                       (Raytrace2D.compute:49-52, Helpers/ComputeHelper.cs:27-31) */
#define RAR_FLAG_COUNT_TESTS 2u     /* also count ray-segment tests the way the reference performs them
                                       (slower; for measurement and parity) */
#define RAR_FLAG_COUNT_EXECUTED 4u  /* with COUNT_TESTS: count only the tests the production kernel evaluates
                                       (it skips shadow rays whose estimate cannot clear the 1e-5 threshold of
                                       Raytrace2D.compute:111); used for the roofline's achieved figure */

/* The uniforms of Trace and ProcessHits: Raytrace2D.compute:5-10,36 set at RayTraceManager.cs:191-201
 * and :227-228; ImpulseLength from RayTraceManager.cs:174.  `bands`/`time_divisor` carry the banded
 * layout of RaytraceOcclusion2D.compute:241-248 (WindowSize); ray_begin/ray_end shard one dispatch
 * across GPUs by contiguous thread-id range. */
#define RAR_FLAG_USE_GRID 8u        /* look walls up through a uniform grid built over the scene instead of scanning
                                       all of them.  Results are identical bit for bit; the test
                                       counters then hold the tests actually evaluated.  Brute force is the
                                       default and the mode the tests/s metric is quoted in. */

typedef struct rar_trace_params {
    float source_pos[2];
    float listener_pos[2];
    float listener_radius;
    float speed_of_sound;
    float input_gain;
    int32_t max_bounce_count;
    uint32_t rng_state_offset; /* Time.frameCount, RayTraceManager.cs:197 */
    int32_t ray_count;         /* the `rayCount` uniform: denominator of the launch angle */
    int32_t debug_ray_count;
    int32_t sample_rate;
    int32_t impulse_length;    /* number of time bins */
    int32_t bands;             /* 1 = broadband; 2..128: slot layout IR[bin*bands + band] */
    float time_divisor;        /* 1: bin=(int)(t*SampleRate); W: bin=(int)(t*SampleRate/W) */
    uint32_t flags;            /* RAR_FLAG_* */
    int64_t ray_begin;         /* thread-id range [ray_begin, ray_end) traced by this call; */
    int64_t ray_end;           /* (0,0) = the whole dispatch */
} rar_trace_params;

typedef struct rar_counters {
    uint64_t ray_bounces;   /* iterations of the bounce loop, Raytrace2D.compute:66 */
    uint64_t nearest_tests; /* intersect() evaluations of the nearest-hit loop (:69-72) */
    uint64_t shadow_tests;  /* intersect() evaluations of checkVis (:40-47) with its early exit */
    uint64_t direct_hits;
    uint64_t nee_hits;
} rar_counters;

typedef struct rar_context rar_context;
typedef struct rar_convolver rar_convolver;
typedef struct rar_ring rar_ring;

/* ---- lifetime ------------------------------------------------------------------------------- */

RAR_API int rar_version(void);

/* Creates a context on CUDA device `device` with its own non-blocking stream.  Replaces the implicit
 * Unity graphics device.  Two IR slots (ping/pong, RayTraceManager.cs:36,212-218) exist from creation
 * and read as zero until configured by rar_ir_clear. */
RAR_API int rar_create(int device, rar_context **out);

/* RayTraceManager.cs:281 OnDestroy -> ComputeHelper.Release(...) (Helpers/ComputeHelper.cs:209-247).
 * NULL is accepted, like `buffer?.Release()`. */
RAR_API int rar_destroy(rar_context *ctx);

/* Message of the most recent failure on this context (or of rar_create when ctx is NULL). */
RAR_API const char *rar_last_error(const rar_context *ctx);

/* Plumbing for hosts that already own a CUDA stream (e.g. to order an NCCL all-reduce after the
 * trace): all later work of the context is enqueued on `cuda_stream` (a cudaStream_t); NULL restores
 * the context's own stream. */
RAR_API int rar_set_stream(rar_context *ctx, void *cuda_stream);

/* Blocks until everything enqueued on the context has finished. */
RAR_API int rar_sync(rar_context *ctx);

/* ---- geometry ------------------------------------------------------------------------------- */

/* RayTraceManager.cs:246-250 UpdateGeometry -> ComputeHelper.CreateStructuredBuffer(ref wallBuffer,
 * activeSegments) (Helpers/ComputeHelper.cs:114-125): copy-in of n x 40 B AoS segments, reallocating
 * when the count grows.  The library re-lays them out for the kernels (endpoint plane + material
 * planes).  n may be 0. */
RAR_API int rar_set_walls(rar_context *ctx, const rar_segment *segments, int32_t n);

/* Build extension for BASELINE config 3: per-wall absorption for `bands` frequency bands (2..128; the
 * experimental variant's WindowSize default is 128, RayTraceManagerComplex.cs:27), row-major [n][bands]; n must
 * equal the current wall count.  Slots of more than 8 bands are traced in chunks of 8 bands. */
RAR_API int rar_set_wall_band_absorption(rar_context *ctx, const float *absorption, int32_t n, int32_t bands);

/* Banded model, air absorption: band b of every arrival of a banded trace is additionally scaled by
 * exp(-alpha_per_m[b] * d), d = the arrival's whole path length in metres (source -> walls -> listener).  The
 * broadband energy -- and with it every threshold and branch of Raytrace2D.compute:66-155 -- is untouched, so ray
 * paths do not change.  exp is a fixed polynomial kernel of the arithmetic contract (histograms stay bit-exact
 * against the oracle).  `bands` must equal the band count of the traces that follow; NULL or bands == 0 switches the
 * attenuation off (the default).  The reference has no air model: its placeholder is the `muffle` factor of
 * RaytraceOcclusion2D.compute:125-126,247-248. */
RAR_API int rar_set_air_absorption(rar_context *ctx, const float *alpha_per_m, int32_t bands);

/* ---- impulse-response slots ------------------------------------------------------------------ */

/* RayTraceManager.cs:169-177 ResetIR -> ClearImpulse (Raytrace2D.compute:167-172), plus the buffer
 * (re)creation of GetActiveIRBuffer (:212-218): slot becomes impulse_length x bands zeros (64-bit
 * fixed point, Q23.40).  Asynchronous. */
RAR_API int rar_ir_clear(rar_context *ctx, int32_t slot, int32_t impulse_length, int32_t bands);

/* Float view of a slot: the un-normalised sum over the frames accumulated so far, the same quantity
 * the reference's float ImpulseResponse buffer holds (division by accumCount happens in the
 * convolution, AudioConvolve.compute:30).  n = impulse_length*bands values.  Blocking. */
RAR_API int rar_ir_read(rar_context *ctx, int32_t slot, float *out, int64_t n);

/* The same read without blocking the caller, the analogue of AsyncGPUReadback.Request on the IR buffer
 * (RayTraceManager.cs:114-121 does this for the convolution output): _begin enqueues the conversion and the copy
 * into pinned memory behind the work already on the stream and returns a ticket; rar_poll(ticket) tells whether it
 * has completed; _end waits if necessary, copies n values to `out` and releases the ticket.  Lets a host overlap
 * the next frame's upload and trace with this frame's readback: the conversion and the copy run on a second stream
 * of the context, behind everything enqueued before the call; a later call that touches the same slot is ordered
 * behind the read, work on other slots (ping/pong) is not. */
RAR_API int rar_ir_read_begin(rar_context *ctx, int32_t slot, int64_t n, int32_t *ticket);
RAR_API int rar_ir_read_end(rar_context *ctx, int32_t ticket, float *out, int64_t n);

/* The slot's exact contents, for bit-exact comparison.  Blocking. */
RAR_API int rar_ir_read_fixed(rar_context *ctx, int32_t slot, int64_t *out, int64_t n);

/* Overwrites a slot with quantised host data (tests and externally supplied IRs).  Blocking. */
RAR_API int rar_ir_write(rar_context *ctx, int32_t slot, const float *ir, int32_t impulse_length, int32_t bands);

/* Device address and word count of a slot's int64 histogram, so the host plumbing can hand it to a
 * collective (one ncclAllReduce(sum, int64) across the GPUs that traced disjoint ray ranges). */
RAR_API int rar_ir_device_ptr(rar_context *ctx, int32_t slot, void **device_ptr, int64_t *n_words);

/* Single-process multi-GPU hosts (a Unity process drives all GPUs itself): sums `slot` over n contexts, one per
 * device, and leaves the total in every context's slot -- the all-reduce of the ray-range sharding, done by a
 * kernel on the first context's device that reads the other devices' histograms directly over NVLink peer
 * memory, followed by peer copies of the total.  All slots must have the same configuration.  Ordered after the
 * work already enqueued on every context; asynchronous.  RAR_ERR_UNSUPPORTED when the devices cannot access
 * each other's memory.  (One-process-per-GPU hosts use rar_ir_device_ptr with their own collective instead.) */
RAR_API int rar_allreduce_slots(rar_context *const *ctxs, int32_t n, int32_t slot);

/* One-process-per-GPU hosts (torchrun-style): the same all-reduce as ONE kernel per rank over NVLink peer
 * memory, with no collective library on the data path.  Every rank owns an exchange region (flag table + staging
 * buffers) that the other processes map through CUDA IPC; the kernel stages the local histogram, meets its peers
 * at a flag barrier in peer memory, sums the ranks' staged histograms straight out of their HBM and leaves the
 * total in the slot (one-shot: every rank reads everything; two-shot: each rank reduces one slice and scatters
 * the totals).  Integer sums: the result is bit-identical to rar_allreduce_slots, to ncclAllReduce over
 * rar_ir_device_ptr, and to an unsharded trace.
 *
 *   rar_exchange_create   allocates the region for histograms of up to capacity_words words and writes the
 *                         RAR_EXCHANGE_HANDLE_BYTES-byte handle the host must deliver to every other rank
 *                         (any transport: torch.distributed all_gather, MPI, a pipe).  Blocking.
 *   rar_exchange_connect  `handles` = world handles in rank order (this rank's own included); maps the peers'
 *                         regions.  Every rank must have created its region first.  Blocking.
 *   rar_exchange_allreduce enqueues the kernel on the context's stream after the work already there; every
 *                         rank must call it the same number of times with the same slot configuration and mode
 *                         (RAR_EXCHANGE_AUTO picks one-shot up to 512 KiB).  world == 1 is a no-op.
 *   rar_exchange_status   waits for the stream and returns RAR_ERR_STATE if a peer failed to reach a barrier
 *                         within the time limit (5 s) -- the kernel gives up instead of spinning forever.
 *   rar_exchange_destroy  unmaps and frees; the host must make sure no peer is still inside a call (a process
 *                         barrier).  Also done by rar_destroy. */
#define RAR_EXCHANGE_HANDLE_BYTES 80
#define RAR_EXCHANGE_MAX_RANKS 16
enum { RAR_EXCHANGE_AUTO = 0, RAR_EXCHANGE_ONE_SHOT = 1, RAR_EXCHANGE_TWO_SHOT = 2 };
RAR_API int rar_exchange_create(rar_context *ctx, int64_t capacity_words, void *handle_out);
RAR_API int rar_exchange_connect(rar_context *ctx, int32_t rank, int32_t world, const void *handles);
RAR_API int rar_exchange_allreduce(rar_context *ctx, int32_t slot, int32_t mode);
RAR_API int rar_exchange_status(rar_context *ctx);
RAR_API int rar_exchange_destroy(rar_context *ctx);

/* ---- clip preparation -------------------------------------------------------------- */

/* RayTraceManager.cs:135-167 LoadSample for a batch of n_clips equally shaped clips: mono mix (channels summed in
 * order, divided by the channel count, :141-147) and, when clip_frequency != sample_rate, the reference's linear
 * resampling (:150-163): ratio = (float)clip_frequency / sample_rate, newLength = RoundToInt(samples / ratio)
 * (half to even), out[i] = Lerp(mono[floor(i*ratio)], mono[min(floor(i*ratio)+1, samples-1)], frac).  Plain IEEE
 * binary32 without contraction, as the C# computes it: results are bit-exact.
 *   raw  [n_clips][samples][channels] interleaved, as AudioClip.GetData returns it;
 *   out  [n_clips][out_stride]; rar_prepared_length(...) values are written per clip (out_stride >= that).
 * rar_prepare_clips takes host arrays and blocks; rar_prepare_clips_device takes device addresses (d_raw 8-byte
 * aligned), enqueues on the context's stream and returns. */
RAR_API int64_t rar_prepared_length(int64_t samples, int32_t clip_frequency, int32_t sample_rate);
RAR_API int rar_prepare_clips(rar_context *ctx, const float *raw, int64_t samples, int32_t channels,
                              int32_t clip_frequency, int32_t sample_rate, int32_t n_clips, float *out,
                              int64_t out_stride);
RAR_API int rar_prepare_clips_device(rar_context *ctx, const void *d_raw, int64_t samples, int32_t channels,
                                     int32_t clip_frequency, int32_t sample_rate, int32_t n_clips, void *d_out,
                                     int64_t out_stride);

/* ---- ray tracing ------------------------------------------------------------------------------ */

/* RayTraceManager.cs:179-210 RunSimulation (Trace dispatch :205) fused with :220-232
 * OnSimulationFinished (ProcessHits dispatch :231): traces the rays and adds every arrival straight
 * into `slot` -- no hit buffer, no hit-count round trip.  Asynchronous.  params->impulse_length and
 * params->bands must match the slot's configuration. */
RAR_API int rar_trace(rar_context *ctx, const rar_trace_params *params, int32_t slot);

/* Block-cyclic sharding of ONE dispatch over `world` GPUs: the thread ids of the dispatch are cut into contiguous
 * chunks of 2^chunk_log2 ids and this call traces chunks rank, rank + world, rank + 2 world, ... (params->ray_begin and
 * ray_end must be 0).  The union over the ranks is the whole dispatch, so the all-reduced histogram is bit-identical to
 * an unsharded rar_trace.  Why not one contiguous range per rank: adjacent rays see similar work, and a rank whose
 * angular sector faces the listener (or a dense part of the scene) runs longer -- measured 13 % between the eight
 * sectors of BASELINE config 2; interleaved chunks give every rank the same mix.  Asynchronous. */
RAR_API int rar_trace_interleaved(rar_context *ctx, const rar_trace_params *params, int32_t slot, int32_t rank, int32_t world,
                                  int32_t chunk_log2);

/* n_frames consecutive frames of the same dispatch (rng_state_offset, +1, ... +n_frames-1) accumulated into `slot`
 * by ONE launch: what n_frames calls of rar_trace with successive Time.frameCount values produce (the per-chunk
 * accumulation of RayTraceManager.cs:82,233, or the offline accumulation before RayTraceManagerComplex.BakeAudio),
 * without n_frames launch latencies -- a 15 000-ray frame is too small to fill 148 SMs on its own.  Asynchronous. */
RAR_API int rar_trace_frames(rar_context *ctx, const rar_trace_params *params, int32_t slot, int32_t n_frames);

/* BASELINE config 4 (batched auralisation): the same dispatch traced for n_listeners listener positions
 * (listeners_xy = x0,y0,x1,y1,...; params->listener_pos is ignored); listener l accumulates into slot
 * first_slot + l, each of which must be configured like `slot` of rar_trace.  The result in every slot is
 * identical to a single-listener rar_trace; for broadband slots the work is fused: a ray's path does not depend on
 * the listener (it is tested, never hit: Raytrace2D.compute:74-84,101-119), so each ray is traced once and only the
 * listener crossing / next-event tests run per listener.  With RAR_FLAG_COUNT_TESTS the counters then hold the tests
 * executed (nearest-hit tests once per ray, not once per listener).  Asynchronous after an initial upload.
 * Listeners shard across GPUs by contiguous range with no exchange. */
RAR_API int rar_trace_listeners(rar_context *ctx, const rar_trace_params *params, const float *listeners_xy,
                                int32_t n_listeners, int32_t first_slot);

/* Same trace, but the arrivals are returned instead of binned: the contents of the reference's
 * rayInfoBuffer (AppendStructuredBuffer<RayInfo>, Raytrace2D.compute:82,116), unordered, with the
 * producing ray/bounce in `keys` (may be NULL).  *count receives the number produced (may exceed
 * capacity; only min(count,capacity) are stored).  Blocking; not a timed path. */
RAR_API int rar_trace_hits(rar_context *ctx, const rar_trace_params *params, rar_ray_info *hits,
                           rar_hit_key *keys, int64_t capacity, int64_t *count);

/* Totals accumulated by traces run with RAR_FLAG_COUNT_TESTS since the last reset.  Blocking. */
RAR_API int rar_get_counters(rar_context *ctx, rar_counters *out, int32_t reset);

/* RayTraceManager.cs:183,204,207: the debugRays buffer of the most recent rar_trace, float4 per
 * vertex, max(100, debug_ray_count) * (max_bounce_count+1) entries.  Blocking. */
RAR_API int rar_get_debug_rays(rar_context *ctx, float *out_xyzw, int64_t n_float4);

/* ---- banded model: filter-bank synthesis ------------------------------------------------
 *
 * A banded slot (RaytraceOcclusion2D.compute:241-248: IR[bin*WindowSize + band], WindowSize set at
 * RayTraceManagerComplex.cs:27-28,75-77) holds one energy response per frequency band.  Every convolution entry
 * point below accepts such a slot and convolves with its broadband synthesis
 *     h[n] = sum_b (g_b * h_b)[n + 127],   h_b[n] = IR[(n / W) * bands + b] when W divides n, else 0,
 * W = the (integer) time_divisor the slot was traced with, g_b = the 255-tap linear-phase windowed-sinc band-pass
 * filter of band b (the filters of contiguous bands sum to a unit impulse, so equal bands synthesise to their common
 * response exactly).  The response then has impulse_length * W samples.  The reference never finished this stage
 * (its FFT/IFFT kernels, RaytraceOcclusion2D.compute:352-425, are not dispatched), so these semantics are this
 * build's; the oracle restates them in direct form.
 *
 * rar_set_band_edges: band b covers [edges_hz[b], edges_hz[b+1]); bands+1 ascending values from 0 to sample_rate/2.
 * NULL restores the default, `bands` equal-width bands -- the linear frequency index of the reference's layout.
 * rar_synthesize_ir: the synthesised response of a slot (un-normalised, like rar_ir_read), n floats.  Blocking. */
RAR_API int rar_set_band_edges(rar_context *ctx, const float *edges_hz, int32_t bands, int32_t sample_rate);
RAR_API int rar_synthesize_ir(rar_context *ctx, int32_t slot, float *out, int64_t n);

/* ---- convolution ------------------------------------------------------------------------------- */

/* RayTraceManager.cs:91-123 ProcessChunk and RayTraceManagerComplex.cs:170-227 BakeAudio: the
 * AudioConvolve kernel (AudioConvolve.compute:13-31) on `in` and the IR in `slot` (a banded slot
 * stands for its filter-bank synthesis, see above; ir_len is then impulse_length * time_divisor):
 *   out[n] = (1/accum_count) * sum_k in[k]*ir[n-k] over |in[k]| > 1e-4,  n in [0, in_len+ir_len),
 *   all zeros when accum_count <= 0.
 * Computed as a uniformly partitioned overlap-save FFT convolution (block 256).
 * rar_convolve is the blocking form (ComputeBuffer.GetData, RayTraceManagerComplex.cs:209). */
RAR_API int rar_convolve(rar_context *ctx, int32_t slot, const float *in, int32_t in_len, int32_t accum_count,
                         float *out, int32_t out_len);

/* Asynchronous form: SetData + Dispatch + AsyncGPUReadback.Request (RayTraceManager.cs:100-114).
 * `in` is copied before returning.  *ticket identifies the request. */
RAR_API int rar_convolve_begin(rar_context *ctx, int32_t slot, const float *in, int32_t in_len,
                               int32_t accum_count, int32_t *ticket);

/* `req.done` (RayTraceManager.cs:115): 1 = complete, 0 = still running, negative = failed. */
RAR_API int rar_poll(rar_context *ctx, int32_t ticket);

/* `req.GetData<float>().CopyTo(result)` (RayTraceManager.cs:121): waits if necessary, copies the
 * in_len+ir_len results and retires the ticket. */
RAR_API int rar_convolve_end(rar_context *ctx, int32_t ticket, float *out, int32_t out_len);

/* ---- batched streaming convolver (BASELINE config 5) --------------------------------------------
 * n_streams independent 48 kHz streams, each convolved with its own IR by uniformly partitioned
 * overlap-save: every call to process consumes `block` new samples per stream and produces `block`
 * output samples per stream.  State (frequency-domain delay lines, IR spectra) is device resident. */
RAR_API int rar_conv_create(rar_context *ctx, int32_t n_streams, int32_t block, int32_t max_ir_len,
                            rar_convolver **out);
RAR_API int rar_conv_destroy(rar_convolver *conv);
/* IR of one stream from host memory; values are multiplied by `scale` (1/accumCount). */
RAR_API int rar_conv_set_ir(rar_convolver *conv, int32_t stream, const float *ir, int32_t ir_len, float scale);
/* IR of one stream taken on the device from a traced slot (banded slots through their filter-bank synthesis),
 * scaled by 1/accum_count. */
RAR_API int rar_conv_set_ir_from_slot(rar_convolver *conv, int32_t stream, int32_t slot, int32_t accum_count);
/* The responses of n consecutive streams in one call: irs[k * ir_stride ...] for stream first_stream + k, or slots[k]
 * scaled by 1 / accum_counts[k] (the slots of one call must have the same shape; banded slots through their
 * filter-bank synthesis).  Asynchronous: the host arrays are copied during the call through alternating pinned
 * staging buffers, and nothing waits for the stream -- loading the 256 responses of BASELINE config 5 this way does
 * not stall the caller the way 256 rar_conv_set_ir calls do. */
RAR_API int rar_conv_set_irs(rar_convolver *conv, int32_t first_stream, int32_t n, const float *irs, int32_t ir_len,
                             int64_t ir_stride, float scale);
RAR_API int rar_conv_set_irs_from_slots(rar_convolver *conv, int32_t first_stream, int32_t n, const int32_t *slots,
                                        const int32_t *accum_counts);
/* Time-varying impulse responses (the ping/pong IR of the streaming path, RayTraceManager.cs:64-123,
 * inside the partitioned convolver instead of one full convolution per chunk).  Like rar_conv_set_ir /
 * rar_conv_set_ir_from_slot, but the new response takes effect with a cross-fade over the next block that
 * rar_conv_process handles: out[n] = y_old[n] + w[n] (y_new[n] - y_old[n]), w[n] = (n + 1) / block, where y_old and
 * y_new are that block's outputs under the old and the new response with the same input history; later blocks use
 * the new response alone.  Only the streams being updated pay for the second multiply-accumulate pass.  A second
 * update before the block is processed replaces the pending response; rar_conv_set_ir cancels it.  Asynchronous. */
RAR_API int rar_conv_update_ir(rar_convolver *conv, int32_t stream, const float *ir, int32_t ir_len, float scale);
RAR_API int rar_conv_update_ir_from_slot(rar_convolver *conv, int32_t stream, int32_t slot, int32_t accum_count);

/* Zeroes the delay lines (start of a stream, AudioManager.StartStreaming). */
RAR_API int rar_conv_reset(rar_convolver *conv);
/* One block for every stream: in/out are host arrays [n_streams][block].  Blocking; copies included. */
RAR_API int rar_conv_process(rar_convolver *conv, const float *in, float *out);
/* Same with device-resident in/out (cudaMalloc'ed by the caller); asynchronous on the context stream. */
RAR_API int rar_conv_process_device(rar_convolver *conv, const float *d_in, float *d_out);
/* Algorithmic HBM bytes one process call moves (delay-line + IR spectra reads, spectrum write, I/O). */
RAR_API int64_t rar_conv_bytes_per_block(const rar_convolver *conv);

/* ---- playback ring ------------------------------------------------------------------
 * AudioManager.cs as a native structure the C# AudioManager can bind: a float ring in pinned host memory, lock-free
 * for ONE producer thread (Unity main thread: PushSamples) and ONE consumer thread (audio thread: OnAudioFilterRead).
 * The reference guards both with lock(bufferLock) (AudioManager.cs:48,59); here the consumer is wait-free
 * (exchange-with-zero per sample) and the producer adds with a compare-exchange per sample, so the audio callback can
 * never wait for a 76 800-sample push.  Needs no GPU (falls back to ordinary memory when CUDA is unavailable).
 *
 *   rar_ring_create   AudioManager.StartStreaming (:26-36): bufferSize = CeilToInt(sampleRate * (reverbDuration + 1)),
 *                     silent, read head 0, streaming.
 *   rar_ring_reset    StartStreaming again on the same ring (not while the consumer is inside rar_ring_drain).
 *   rar_ring_stop     StopStreaming (:38-43): push and drain become no-ops, like `if (!isStreaming) return`.
 *   rar_ring_push     PushSamples (:45-54): ring[(sample_offset % size + i) % size] += samples[i]; sample_offset >= 0.
 *   rar_ring_drain    OnAudioFilterRead (:56-69): for each of data_length / channels frames, the sample at the read
 *                     head goes to every channel of the frame and is zeroed; the rest of `data` is left untouched.
 *   rar_ring_frames_drained  frames handed out since creation / reset (monotone): lets a producer pace itself.
 *   rar_conv_process_to_ring one block of the streaming convolver (like rar_conv_process, `in` = host [n_streams][block])
 *                     whose output goes straight into rings: stream s is pushed to rings[s] at sample_offset
 *                     (rings[s] == NULL: that stream's output is dropped).  Blocking; producer-side call. */
RAR_API int rar_ring_create(int32_t output_sample_rate, float reverb_duration, rar_ring **out);
RAR_API int rar_ring_destroy(rar_ring *ring);
RAR_API int rar_ring_reset(rar_ring *ring);
RAR_API int rar_ring_stop(rar_ring *ring);
RAR_API int32_t rar_ring_size(const rar_ring *ring);
RAR_API int32_t rar_ring_is_pinned(const rar_ring *ring);
RAR_API int64_t rar_ring_frames_drained(const rar_ring *ring);
RAR_API int rar_ring_push(rar_ring *ring, const float *samples, int32_t n, int64_t sample_offset);
RAR_API int rar_ring_drain(rar_ring *ring, float *data, int32_t data_length, int32_t channels);
RAR_API int rar_conv_process_to_ring(rar_convolver *conv, const float *in, rar_ring *const *rings, int64_t sample_offset);

/* ---- measurement helpers ------------------------------------------------------------------------ */

/* Device facts for the bench line: SM count, max SM clock (kHz), shared memory per block opt-in. */
RAR_API int rar_device_info(rar_context *ctx, int32_t *sm_count, int32_t *sm_clock_khz, int32_t *smem_optin_bytes);

/* Measures the FP32 FMA issue peak of the device with a register-only FFMA loop (lane-ops/s, one
 * fused multiply-add = 1 lane-op): the denominator of the ray stage's roofline. */
RAR_API int rar_measure_fp32_peak(rar_context *ctx, double *lane_ops_per_s);

/* Device self-test of the ray stage's arithmetic contract: the kernels for opaque, coordinate-bounded scenes evaluate
 * 1/x, sqrt(x) and a/b as the fast paths of the correctly rounded operations with ONE range test per group of
 * operations instead of one guard per operation (csrc/rar_math.cuh).  This compares those forms with the correctly
 * rounded device intrinsics, bit for bit, on n_samples pseudo-random operand sets spread over the whole admitted
 * exponent range; mismatches[0..4] receive the number of differing results (reciprocal, square root, division,
 * division by a launch-invariant divisor, range test).  All zeros is the only passing result.  Blocking. */
RAR_API int rar_selftest_arithmetic(rar_context *ctx, int64_t n_samples, uint32_t seed, uint64_t *mismatches);

/* Test hook for the uniform grid of RAR_FLAG_USE_GRID: builds it for the current walls if necessary (on the device,
 * asynchronously, as a grid trace would) and returns its dimensions, the total length of its cell lists and a digest
 * of them, so that a test can hold the device-built grid against the host builder's.  Blocking. */
RAR_API int rar_debug_grid(rar_context *ctx, int32_t *nx, int32_t *ny, int64_t *n_items, uint64_t *digest);

/* Number of kernels this library has launched on the context since creation (bench: gpu_launches). */
RAR_API int64_t rar_launch_count(const rar_context *ctx);

#ifdef __cplusplus
}
#endif
#endif /* RAR2D_H */

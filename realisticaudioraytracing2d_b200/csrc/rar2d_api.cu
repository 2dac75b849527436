// rar2d_api.cu -- the C-ABI of include/rar2d.h: contexts, wall upload, IR slots, trace and convolution
// entry points.  Host-side bookkeeping only; all compute is in trace_kernel.cu and conv_kernels.cu.
// There is no CPU path here: every compute entry point needs a CUDA device and fails otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "rar_internal.h"
#include "rar_layout.h"
#include "rar_synth16.cuh"

using namespace rar;

namespace {

thread_local std::string g_create_error;

constexpr int kMaxSlots = 1 << 16;
constexpr int kBlock = 256;

struct Slot {
    long long *d_hist = nullptr;
    long long cap_words = 0;
    int impulse_length = 0;
    int bands = 0;
    int time_stride = 1;  // samples per time bin (the time_divisor of the traces that filled the slot); 0: not an integer
    bool configured = false;
    bool aliased = false;  // rar_ir_device_ptr handed the histogram out: cached spectra are never trusted again
    // cached spectra of the slot's IR partitions (one-shot convolution)
    float2 *d_H = nullptr;
    int H_cap = 0;  // partitions allocated
    int H_parts = 0;
    bool H_valid = false;
    // rar_ir_read_begin converts and copies on the context's read stream; whoever touches the histogram next on the
    // main stream waits for it (get_slot)
    cudaEvent_t read_done = nullptr;
    bool read_pending = false;
};

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

template <class T>
struct PinnedBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMallocHost((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct Ticket {
    bool active = false;
    bool failed = false;
    int out_len = 0;
    cudaEvent_t done = nullptr;
    cudaEvent_t ready = nullptr;  // ir_read: recorded on the main stream, awaited by the read stream
    PinnedBuf<float> h_in, h_out;
    DevBuf<float> d_x, d_out;
    DevBuf<float2> d_X, d_Y;
};

// Exchange region of the multi-process peer-memory all-reduce (exchange_kernel.cu):
//   [flag table][status + padding][in 0][in 1][out 0][out 1], each buffer stage_bytes long.
struct Exchange {
    char *region = nullptr;
    size_t region_bytes = 0, stage_bytes = 0;
    long long cap_words = 0;
    bool connected = false;
    int rank = 0, world = 1;
    unsigned epoch = 0;
    void *opened[kExMaxRanks] = {};  // bases returned by cudaIpcOpenMemHandle (to close)
    char *peer[kExMaxRanks] = {};    // every rank's region as addressable from this process
};
constexpr size_t kExHeaderBytes = kExFlagWords * sizeof(unsigned) + 256;

struct ExchangeHandle {  // RAR_EXCHANGE_HANDLE_BYTES
    cudaIpcMemHandle_t mem;
    unsigned long long offset;     // of the region inside the IPC allocation
    unsigned long long cap_words;
};
static_assert(sizeof(ExchangeHandle) == RAR_EXCHANGE_HANDLE_BYTES, "exchange handle layout");
static_assert(kExMaxRanks == RAR_EXCHANGE_MAX_RANKS, "rank limit");

}  // namespace

struct rar_context {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t read_stream = nullptr;  // conversion + device-to-host copy of rar_ir_read_begin: overlaps the next trace
    cudaStream_t stream = nullptr;
    DeviceFacts dev{};
    std::string err;
    long long launches = 0;

    // walls
    DevBuf<f4> d_planes;  // geo | mat0 | (mat1, end): one allocation, one upload per rar_set_walls
    struct PlaneF4 { f4 *p = nullptr; void release() { p = nullptr; } } d_geo, d_mat0, d_pair_a, d_pair_b;   // views into d_planes
    struct PlaneF2 { f2 *p = nullptr; void release() { p = nullptr; } } d_mat1;
    DevBuf<float> d_band_abs;
    int n_walls = -1;  // -1: never set
    int band_rows = 0, band_count = 0;
    std::vector<float> air;  // air absorption per band, 1/m (empty: none); rar_set_air_absorption
    PlaneF2 d_end;                     // wall end points (the grid builder needs them unrounded)
    PinnedBuf<f4> h_planes;            // pinned staging of the geo | mat0 | mat1 | end planes of one upload
    cudaEvent_t walls_uploaded = nullptr;  // the staging buffer may be rewritten once this has completed
    PinnedBuf<float> h_bands;          // pinned staging of the band-absorption table
    cudaEvent_t bands_uploaded = nullptr;
    PinnedBuf<unsigned char> h_listeners;  // pinned staging of rar_trace_listeners' positions and slot addresses
    cudaEvent_t listeners_uploaded = nullptr;
    std::vector<rar_segment> h_walls;  // kept for the lazy grid build
    GridFrame grid_frame_;             // frame of the current grid (nx == 0: none could be built)
    GridHost h_grid;                   // host-built lists (fallback for scenes whose registration bound is enormous)
    DevBuf<uint32_t> d_grid_start, d_grid_items, d_grid_cnt;
    DevBuf<f4> d_grid_geo;
    long long grid_items_bound = 0;
    bool grid_valid = false;
    bool walls_are_opaque = false;  // no wall has transmission > 0
    bool walls_are_bounded = false; // every coordinate finite and within 2^30 (rar_layout.h walls_bounded)

    std::vector<Slot> slots;
    // filter bank of the banded model: edges as fractions of Nyquist (empty: equal-width bands), spectra per band count
    std::vector<float> band_edges;
    DevBuf<float2> d_band_G;
    int band_G_bands = 0;  // band count d_band_G was built for (0: stale)
    DevBuf<float> d_synth;  // synthesised broadband response (scratch)
    PinnedBuf<float> h_band_taps;
    DevBuf<float> d_band_taps;
    PinnedBuf<float2> h_band_T;  // tables of the production synthesis kernel (rar_synth16.cuh synth_tables)
    DevBuf<float2> d_band_T;
    DevBuf<unsigned long long> d_counters;  // 5 counters + 1 hit count
    DevBuf<f4> d_debug;
    int debug_entries = 0;
    DevBuf<float> d_irf;  // float view of a slot (scratch)
    DevBuf<float> d_clip_raw, d_clip_out;  // rar_prepare_clips staging
    DevBuf<f2> d_listeners;                        // batched-listener launch arguments
    DevBuf<unsigned long long *> d_listener_hists;
    std::vector<Ticket *> tickets;
    std::vector<rar_convolver *> convolvers;
    Exchange ex;
};

struct rar_convolver {
    rar_context *ctx = nullptr;
    StreamConv c{};
    int max_ir_len = 0;
    DevBuf<float2> H, fdl, partial;
    DevBuf<float> prev, d_in, d_out, d_irf;
    PinnedBuf<float> h_ir;
    // batched response loads (rar_conv_set_irs*): device scratch for a group of responses, two pinned staging halves
    DevBuf<float> d_batch;
    PinnedBuf<float> h_batch, h_out;
    cudaEvent_t batch_done[2] = {nullptr, nullptr};  // the staging half may be rewritten once its copy has completed
    // cross-faded impulse-response updates (allocated on first use)
    DevBuf<float2> H2, partial2;   // new spectra of the fading streams, their partial sums
    DevBuf<int> d_fade, d_list;    // [S] flags, compact list of fading streams
    std::vector<int> fade;         // host copy of the flags
};

namespace {

int fail(rar_context *ctx, int code, const char *fmt, const char *detail = "") {
    char buf[512];
    snprintf(buf, sizeof buf, fmt, detail);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}

#define RAR_CUDA(ctx, call)                                                              \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            std::string m__ = std::string(#call) + ": " + cudaGetErrorString(e__);       \
            cudaGetLastError();                                                          \
            return fail((ctx), e__ == cudaErrorMemoryAllocation ? RAR_ERR_NOMEM : RAR_ERR_CUDA, "%s", m__.c_str()); \
        }                                                                                \
    } while (0)

#define RAR_ENTER(ctx)                                                        \
    do {                                                                      \
        if (!(ctx)) return fail(nullptr, RAR_ERR_INVALID, "null context");    \
        cudaError_t e0__ = cudaSetDevice((ctx)->device);                      \
        if (e0__ != cudaSuccess) return fail((ctx), RAR_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e0__)); \
    } while (0)

Slot *get_slot(rar_context *ctx, int slot, bool create) {
    if (slot < 0 || slot >= kMaxSlots) return nullptr;
    if ((size_t)slot >= ctx->slots.size()) {
        if (!create) return nullptr;
        ctx->slots.resize(slot + 1);
    }
    Slot *S = &ctx->slots[slot];
    if (S->read_pending) {  // an asynchronous read of this slot may still be in flight on the read stream
        cudaStreamWaitEvent(ctx->stream, S->read_done, 0);
        S->read_pending = false;
    }
    return S;
}

int check_trace_params(rar_context *ctx, const rar_trace_params *p) {
    if (!p) return fail(ctx, RAR_ERR_INVALID, "null params");
    if (ctx->n_walls < 0) return fail(ctx, RAR_ERR_STATE, "rar_set_walls has not been called");
    if (p->ray_count <= 0) return fail(ctx, RAR_ERR_INVALID, "ray_count must be positive");
    if (p->max_bounce_count < 0) return fail(ctx, RAR_ERR_INVALID, "max_bounce_count must be >= 0");
    if (p->bands < 1 || p->bands > 128) return fail(ctx, RAR_ERR_UNSUPPORTED, "bands must be 1..128");
    if (p->bands > 1 && (ctx->band_count != p->bands || ctx->band_rows != ctx->n_walls))
        return fail(ctx, RAR_ERR_STATE, "banded trace needs rar_set_wall_band_absorption for the current walls");
    if (p->impulse_length < 0 || p->sample_rate <= 0) return fail(ctx, RAR_ERR_INVALID, "bad impulse_length/sample_rate");
    if (!(p->time_divisor > 0.0f)) return fail(ctx, RAR_ERR_INVALID, "time_divisor must be positive");
    if (p->ray_begin < 0 || p->ray_end < p->ray_begin || p->ray_end > 0xffffffffLL)
        return fail(ctx, RAR_ERR_INVALID, "bad ray range");
    return RAR_OK;
}

// Builds (once per wall upload) and attaches the uniform grid when the call asks for it.  The lists are built on the
// device by kernels enqueued on the context's stream (grid_kernel.cu); the host computes the frame and an upper
// bound of the list length in one pass over the walls and waits for nothing.
constexpr long long kGridDeviceBuildMaxItems = 64LL << 20;

int attach_grid(rar_context *ctx, const rar_trace_params *p, TraceLaunch &a) {
    a.use_grid = 0;
    if (!(p->flags & RAR_FLAG_USE_GRID) || ctx->n_walls <= 0) return RAR_OK;
    if (!ctx->grid_valid) {
        long long bound = 0;
        ctx->grid_frame_ = grid_frame(ctx->h_walls.data(), ctx->n_walls, &bound);
        const GridFrame &fr = ctx->grid_frame_;
        if (fr.nx > 0) {
            const size_t n_cells = (size_t)fr.nx * fr.ny;
            RAR_CUDA(ctx, ctx->d_grid_start.reserve(n_cells + 1));
            if (bound <= kGridDeviceBuildMaxItems) {
                RAR_CUDA(ctx, ctx->d_grid_cnt.reserve(n_cells));
                RAR_CUDA(ctx, ctx->d_grid_items.reserve((size_t)bound + 1));
                RAR_CUDA(ctx, ctx->d_grid_geo.reserve((size_t)bound + 1));
                RAR_CUDA(ctx, launch_grid_build(ctx->d_geo.p, ctx->d_end.p, ctx->n_walls, fr, ctx->d_grid_cnt.p, ctx->d_grid_start.p,
                                                ctx->d_grid_items.p, ctx->d_grid_geo.p, ctx->stream));
                ctx->launches += 4;
            } else {
                // a scene of thousands of walls that each span the whole extent: build on the host (blocking)
                build_grid(ctx->h_walls.data(), ctx->n_walls, ctx->h_grid);
                const GridHost &g = ctx->h_grid;
                RAR_CUDA(ctx, ctx->d_grid_items.reserve(g.items.size() + 1));
                RAR_CUDA(ctx, ctx->d_grid_geo.reserve(g.items.size() + 1));
                RAR_CUDA(ctx, cudaMemcpyAsync(ctx->d_grid_start.p, g.cell_start.data(), g.cell_start.size() * sizeof(uint32_t),
                                              cudaMemcpyHostToDevice, ctx->stream));
                if (!g.items.empty()) {
                    RAR_CUDA(ctx, cudaMemcpyAsync(ctx->d_grid_items.p, g.items.data(), g.items.size() * sizeof(uint32_t),
                                                  cudaMemcpyHostToDevice, ctx->stream));
                    RAR_CUDA(ctx, cudaMemcpyAsync(ctx->d_grid_geo.p, g.item_geo.data(), g.items.size() * sizeof(f4),
                                                  cudaMemcpyHostToDevice, ctx->stream));
                }
                RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            }
        }
        ctx->grid_items_bound = bound;
        ctx->grid_valid = true;
    }
    const GridFrame &g = ctx->grid_frame_;
    if (g.nx <= 0) return RAR_OK;  // no grid could be built (non-finite coordinates): brute force
    a.grid.x0 = g.x0;
    a.grid.y0 = g.y0;
    a.grid.cw = g.cw;
    a.grid.ch = g.ch;
    a.grid.inv_cw = 1.0f / g.cw;
    a.grid.inv_ch = 1.0f / g.ch;
    a.grid.nx = g.nx;
    a.grid.ny = g.ny;
    a.grid.cell_start = ctx->d_grid_start.p;
    a.grid.items = ctx->d_grid_items.p;
    a.grid.item_geo = ctx->d_grid_geo.p;
    a.use_grid = 1;
    return RAR_OK;
}

void fill_launch(rar_context *ctx, const rar_trace_params *p, TraceLaunch &a) {
    std::memset(&a, 0, sizeof a);
    a.geo = ctx->d_geo.p;
    a.pair_a = ctx->d_pair_a.p;
    a.pair_b = ctx->d_pair_b.p;
    a.mat0 = ctx->d_mat0.p;
    a.mat1 = ctx->d_mat1.p;
    a.band_abs = p->bands > 1 ? ctx->d_band_abs.p : nullptr;
    a.n_walls = ctx->n_walls;
    a.bands = p->bands > 1 ? 8 : 1;  // banded slots are traced in chunks of 8 bands
    a.band_total = p->bands;
    a.band_offset = 0;
    a.band_valid = p->bands > 8 ? 8 : p->bands;
    a.p = ray_consts(*p);
    a.opaque = ctx->walls_are_opaque ? 1 : 0;
    a.tile_counter = ctx->d_counters.p + 6;  // words 0-4: test counters, 5: hit count, 6: tile counter
    a.spec_ok = spec_ranges_ok(*p, ctx->walls_are_bounded, ctx->walls_are_opaque) ? 1 : 0;
    ray_range(*p, a.ray_begin, a.ray_end);
}

Ticket *get_ticket(rar_context *ctx, int id) {
    if (id < 0 || (size_t)id >= ctx->tickets.size()) return nullptr;
    return ctx->tickets[id];
}

void free_ticket(Ticket *t) {
    if (!t) return;
    if (t->done) cudaEventDestroy(t->done);
    if (t->ready) cudaEventDestroy(t->ready);
    t->h_in.release();
    t->h_out.release();
    t->d_x.release();
    t->d_out.release();
    t->d_X.release();
    t->d_Y.release();
    delete t;
}

// An inactive ticket of the context (or a new one); negative status on failure.
int acquire_ticket(rar_context *ctx) {
    for (size_t i = 0; i < ctx->tickets.size(); i++)
        if (!ctx->tickets[i]->active) return (int)i;
    Ticket *t = new (std::nothrow) Ticket();
    if (!t) return fail(ctx, RAR_ERR_NOMEM, "out of host memory");
    if (cudaEventCreateWithFlags(&t->done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&t->ready, cudaEventDisableTiming) != cudaSuccess) {
        if (t->done) cudaEventDestroy(t->done);
        delete t;
        return fail(ctx, RAR_ERR_CUDA, "cudaEventCreate failed");
    }
    ctx->tickets.push_back(t);
    return (int)ctx->tickets.size() - 1;
}

// Spectra of the band filters for `bands` bands (built on first use and after rar_set_band_edges).
int ensure_band_filters(rar_context *ctx, int bands) {
    if (ctx->band_G_bands == bands) return RAR_OK;
    if (!ctx->band_edges.empty() && (int)ctx->band_edges.size() != bands + 1)
        return fail(ctx, RAR_ERR_STATE, "rar_set_band_edges was called for a different band count than this slot has");
    const size_t taps = (size_t)bands * kBlock;  // each filter zero-padded to one partition
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the staging buffers may still feed an earlier upload
    RAR_CUDA(ctx, ctx->h_band_taps.reserve(taps));
    RAR_CUDA(ctx, ctx->d_band_taps.reserve(taps));
    RAR_CUDA(ctx, ctx->d_band_G.reserve(taps));
    std::memset(ctx->h_band_taps.p, 0, taps * sizeof(float));
    for (int b = 0; b < bands; b++) {
        const double lo = ctx->band_edges.empty() ? (double)b / bands : ctx->band_edges[b];
        const double hi = ctx->band_edges.empty() ? (double)(b + 1) / bands : ctx->band_edges[b + 1];
        band_filter_taps(lo, hi, ctx->h_band_taps.p + (size_t)b * kBlock);
    }
    RAR_CUDA(ctx, cudaMemcpyAsync(ctx->d_band_taps.p, ctx->h_band_taps.p, taps * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    // one partition per band: H[b] = rfft512([g_b, 0 ...])
    RAR_CUDA(ctx, launch_ir_spectra(ctx->d_band_taps.p, (int)taps, ctx->d_band_G.p, bands, kBlock, ctx->stream));
    ctx->launches++;
    if (band_synth16_applicable(bands, 1)) {
        const size_t len = synth_table_len(bands);
        RAR_CUDA(ctx, ctx->h_band_T.reserve(len));
        RAR_CUDA(ctx, ctx->d_band_T.reserve(len));
        synth_tables(ctx->h_band_taps.p, bands, reinterpret_cast<f2 *>(ctx->h_band_T.p));
        RAR_CUDA(ctx, cudaMemcpyAsync(ctx->d_band_T.p, ctx->h_band_T.p, len * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->band_G_bands = bands;
    return RAR_OK;
}

// Synthesis of a batch of banded slots of one shape: the register-transform kernel where it applies
// (RAR_NO_FAST_SYNTH=1 keeps the generic shared-memory kernel, for A/B measurements and tests).
cudaError_t synth_batch(rar_context *ctx, const BandSynthBatch &batch, int n_items, int bins, int bands, int stride, int out_len) {
    static const bool no_fast = [] { const char *e = getenv("RAR_NO_FAST_SYNTH"); return e && e[0] == '1'; }();
    if (!no_fast && band_synth16_applicable(bands, stride))
        return launch_band_synth16(batch, n_items, bins, bands, ctx->d_band_T.p, out_len, ctx->d_counters.p + 7, ctx->dev.sm_count,
                                   ctx->stream);  // (d_counters words 0-4: test counters, 5: hit count, 6: ray tiles, 7: synthesis work)
    return launch_band_synth(batch, n_items, bins, bands, stride, ctx->d_band_G.p, out_len, ctx->stream);
}

// Samples of the broadband response a slot stands for: bins x samples per bin.
long long slot_ir_samples(const Slot &S) { return (long long)S.impulse_length * (S.time_stride > 0 ? S.time_stride : 1); }

int check_convolvable(rar_context *ctx, const Slot &S) {
    if (S.bands > 1 && S.time_stride <= 0)
        return fail(ctx, RAR_ERR_UNSUPPORTED, "the slot was traced with a non-integer time_divisor: no sample grid to synthesise on");
    if (slot_ir_samples(S) > 0x7fffffffLL - 2 * kBlock) return fail(ctx, RAR_ERR_UNSUPPORTED, "impulse response too long");
    return RAR_OK;
}

// The slot's response as floats on the device, scaled: the histogram itself for a broadband slot, the filter-bank
// synthesis for a banded one (out must hold slot_ir_samples(S) floats).
int slot_response(rar_context *ctx, Slot &S, float scale, float *d_out) {
    const int n = (int)slot_ir_samples(S);
    if (S.bands == 1 && S.time_stride == 1) {
        RAR_CUDA(ctx, launch_fixed_to_float(S.d_hist, d_out, n, scale, ctx->stream));
        ctx->launches++;
        return RAR_OK;
    }
    int rc = ensure_band_filters(ctx, S.bands);
    if (rc != RAR_OK) return rc;
    BandSynthBatch batch;
    std::memset(&batch, 0, sizeof batch);
    batch.items[0] = BandSynthItem{S.d_hist, d_out, scale, 0};
    RAR_CUDA(ctx, cudaMemsetAsync(d_out, 0, (size_t)n * sizeof(float), ctx->stream));
    RAR_CUDA(ctx, synth_batch(ctx, batch, 1, S.impulse_length, S.bands, S.time_stride, n));
    ctx->launches++;
    return RAR_OK;
}

// Makes the cached partition spectra of a slot current.
int ensure_slot_spectra(rar_context *ctx, Slot &S) {
    const int n_samples = (int)slot_ir_samples(S);
    const int n_part = (n_samples + kBlock - 1) / kBlock;
    if (S.H_valid && !S.aliased && S.H_parts == n_part) return RAR_OK;
    if (n_part > S.H_cap) {
        if (S.d_H) cudaFree(S.d_H);
        S.d_H = nullptr;
        S.H_cap = 0;
        RAR_CUDA(ctx, cudaMalloc((void **)&S.d_H, (size_t)(n_part + 8) * kBlock * sizeof(float2)));
        S.H_cap = n_part + 8;
    }
    RAR_CUDA(ctx, ctx->d_irf.reserve((size_t)n_samples + 1));
    int rc = slot_response(ctx, S, 1.0f, ctx->d_irf.p);
    if (rc != RAR_OK) return rc;
    RAR_CUDA(ctx, launch_ir_spectra(ctx->d_irf.p, n_samples, S.d_H, n_part, kBlock, ctx->stream));
    ctx->launches += 1;
    S.H_parts = n_part;
    S.H_valid = true;
    return RAR_OK;
}

}  // namespace

// ---- lifetime -------------------------------------------------------------------------------------

extern "C" {

int rar_version(void) { return RAR_VERSION; }

int rar_create(int device, rar_context **out) {
    if (!out) return fail(nullptr, RAR_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(nullptr, RAR_ERR_CUDA, "no CUDA device: %s (there is no CPU fallback)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n_dev) return fail(nullptr, RAR_ERR_INVALID, "device index out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, RAR_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, RAR_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(nullptr, RAR_ERR_UNSUPPORTED, "%s", "device is not sm_100-class; this library ships sm_100a code only");
    rar_context *ctx = new (std::nothrow) rar_context();
    if (!ctx) return fail(nullptr, RAR_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    ctx->dev.sm_count = prop.multiProcessorCount;
    ctx->dev.smem_optin = (int)prop.sharedMemPerBlockOptin;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    ctx->dev.sm_clock_khz = khz;
    e = cudaStreamCreateWithFlags(&ctx->read_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        if (ctx->read_stream) cudaStreamDestroy(ctx->read_stream);
        delete ctx;
        return fail(nullptr, RAR_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    ctx->stream = ctx->own_stream;
    ctx->slots.resize(2);  // ping / pong, RayTraceManager.cs:36
    conv_init_tables(ctx->own_stream);
    synth_init_tables(ctx->own_stream);
    e = ctx->d_counters.reserve(8);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_counters.p, 0, 8 * sizeof(unsigned long long), ctx->own_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->own_stream);  // creation-time initialisation is complete on return
    if (e != cudaSuccess) {
        cudaStreamDestroy(ctx->own_stream);
        cudaStreamDestroy(ctx->read_stream);
        delete ctx;
        return fail(nullptr, RAR_ERR_CUDA, "counter allocation: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return RAR_OK;
}

int rar_conv_destroy(rar_convolver *conv);

int rar_destroy(rar_context *ctx) {
    if (!ctx) return RAR_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->read_stream) cudaStreamSynchronize(ctx->read_stream);
    rar_exchange_destroy(ctx);
    while (!ctx->convolvers.empty()) rar_conv_destroy(ctx->convolvers.back());
    for (Ticket *t : ctx->tickets) free_ticket(t);
    for (Slot &s : ctx->slots) {
        if (s.d_hist) cudaFree(s.d_hist);
        if (s.d_H) cudaFree(s.d_H);
        if (s.read_done) cudaEventDestroy(s.read_done);
    }
    ctx->h_planes.release();
    ctx->h_bands.release();
    ctx->h_listeners.release();
    ctx->d_end.release();
    ctx->d_grid_cnt.release();
    if (ctx->walls_uploaded) cudaEventDestroy(ctx->walls_uploaded);
    if (ctx->bands_uploaded) cudaEventDestroy(ctx->bands_uploaded);
    if (ctx->listeners_uploaded) cudaEventDestroy(ctx->listeners_uploaded);
    ctx->d_planes.release();
    ctx->d_geo.release();
    ctx->d_mat0.release();
    ctx->d_mat1.release();
    ctx->d_band_abs.release();
    ctx->d_counters.release();
    ctx->d_debug.release();
    ctx->d_irf.release();
    ctx->d_clip_raw.release();
    ctx->d_clip_out.release();
    ctx->d_listeners.release();
    ctx->d_listener_hists.release();
    ctx->d_band_G.release();
    ctx->d_synth.release();
    ctx->h_band_taps.release();
    ctx->d_band_taps.release();
    ctx->h_band_T.release();
    ctx->d_band_T.release();
    ctx->d_grid_start.release();
    ctx->d_grid_items.release();
    ctx->d_grid_geo.release();
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->read_stream) cudaStreamDestroy(ctx->read_stream);
    delete ctx;
    return RAR_OK;
}

const char *rar_last_error(const rar_context *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rar_set_stream(rar_context *ctx, void *cuda_stream) {
    RAR_ENTER(ctx);
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return RAR_OK;
}

int rar_sync(rar_context *ctx) {
    RAR_ENTER(ctx);
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->read_stream));  // (its work was enqueued behind the main stream's)
    return RAR_OK;
}

// ---- geometry ---------------------------------------------------------------------------------------

int rar_set_walls(rar_context *ctx, const rar_segment *segments, int32_t n) {
    RAR_ENTER(ctx);
    if (n < 0 || (n > 0 && !segments)) return fail(ctx, RAR_ERR_INVALID, "bad wall array");
    static_assert(sizeof(rar_segment) == 40, "Segment must be 40 bytes (Helpers/SceneHelper.cs:15-22)");
    static_assert(sizeof(rar_ray_info) == 16, "RayInfo must be 16 bytes (RayTraceManager.cs:43)");
    const size_t pad = ((size_t)n + 2 + 1) & ~(size_t)1;  // mat1 plane is bulk-copied in 16-byte units; even => planes stay 16-byte aligned
    // Pinned staging: the planes go up as asynchronous copies and the call returns without waiting for
    // them; the next upload waits (normally not at all) for this one before it rewrites the staging buffer.
    if (ctx->walls_uploaded) RAR_CUDA(ctx, cudaEventSynchronize(ctx->walls_uploaded));
    else RAR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->walls_uploaded, cudaEventDisableTiming));
    if (pad * 4 > ctx->h_planes.cap) RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // reallocation frees the old buffer
    RAR_CUDA(ctx, ctx->h_planes.reserve(pad * 4));   // geo | mat0 | (mat1, end): two 8-byte planes share the third f4 plane | (pair_a, pair_b)
    f4 *h_geo = ctx->h_planes.p, *h_mat0 = h_geo + pad;
    f2 *h_mat1 = reinterpret_cast<f2 *>(h_mat0 + pad);
    f2 *h_end = h_mat1 + pad;
    f4 *h_pair_a = h_geo + 3 * pad, *h_pair_b = h_pair_a + pad / 2;   // pad / 2 >= (n + 1) / 2 records each
    std::memset(h_geo, 0, pad * 4 * sizeof(f4));
    split_walls(segments, n, h_geo, h_mat0, h_mat1);
    for (int w = 0; w < n; w++) h_end[w] = f2{segments[w].end[0], segments[w].end[1]};
    pair_planes(h_geo, n, h_pair_a, h_pair_b);
    RAR_CUDA(ctx, ctx->d_planes.reserve(pad * 4));  // (growth frees the old planes: cudaFree waits for the device)
    ctx->d_geo.p = ctx->d_planes.p;
    ctx->d_mat0.p = ctx->d_geo.p + pad;
    ctx->d_mat1.p = reinterpret_cast<f2 *>(ctx->d_mat0.p + pad);
    ctx->d_end.p = ctx->d_mat1.p + pad;
    ctx->d_pair_a.p = ctx->d_geo.p + 3 * pad;
    ctx->d_pair_b.p = ctx->d_pair_a.p + pad / 2;
    // The previous planes may still be read by an enqueued trace; stream order makes the copy safe.  The device
    // planes mirror the staging layout, so the whole scene is ONE host-to-device copy.
    RAR_CUDA(ctx, cudaMemcpyAsync(ctx->d_planes.p, h_geo, pad * 4 * sizeof(f4), cudaMemcpyHostToDevice, ctx->stream));
    RAR_CUDA(ctx, cudaEventRecord(ctx->walls_uploaded, ctx->stream));
    if (n != ctx->n_walls) {
        ctx->band_rows = 0;
        ctx->band_count = 0;
    }
    ctx->n_walls = n;
    ctx->h_walls.assign(segments, segments + n);
    ctx->grid_valid = false;
    ctx->walls_are_opaque = walls_opaque(segments, n);
    ctx->walls_are_bounded = walls_bounded(segments, n);
    return RAR_OK;
}

int rar_set_wall_band_absorption(rar_context *ctx, const float *absorption, int32_t n, int32_t bands) {
    RAR_ENTER(ctx);
    if (ctx->n_walls < 0) return fail(ctx, RAR_ERR_STATE, "rar_set_walls has not been called");
    if (n != ctx->n_walls) return fail(ctx, RAR_ERR_INVALID, "row count must equal the wall count");
    if (bands < 2 || bands > 128) return fail(ctx, RAR_ERR_UNSUPPORTED, "bands must be 2..128");
    if (n > 0 && !absorption) return fail(ctx, RAR_ERR_INVALID, "null absorption table");
    // Staged through pinned memory and copied on the context's stream: ordered behind any banded trace still in
    // flight (which reads the old table) and ahead of the next one, without blocking the caller.
    const size_t words = (size_t)n * bands;
    if (ctx->bands_uploaded) RAR_CUDA(ctx, cudaEventSynchronize(ctx->bands_uploaded));
    else RAR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->bands_uploaded, cudaEventDisableTiming));
    if (words + 16 > ctx->d_band_abs.cap || words > ctx->h_bands.cap) RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // growth frees buffers in use
    RAR_CUDA(ctx, ctx->d_band_abs.reserve(words + 16));  // a short last chunk of 8 reads past its row
    RAR_CUDA(ctx, ctx->h_bands.reserve(words + 1));
    if (n > 0) {
        std::memcpy(ctx->h_bands.p, absorption, words * sizeof(float));
        RAR_CUDA(ctx, cudaMemcpyAsync(ctx->d_band_abs.p, ctx->h_bands.p, words * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        RAR_CUDA(ctx, cudaEventRecord(ctx->bands_uploaded, ctx->stream));
    }
    ctx->band_rows = n;
    ctx->band_count = bands;
    return RAR_OK;
}

// ---- filter bank of the banded model ------------------------------------------------------------------

}  // extern "C"

// g[n] = w[n] (hi sinc(hi m) - lo sinc(lo m)), m = n - 127, w = Hann over the 255 taps (nonzero at both ends).
// Contiguous bands telescope: sum_b g_b[n] = w[n] sinc(m) = [n == 127].
void rar::band_filter_taps(double lo, double hi, float *g) {
    const double pi = 3.14159265358979323846;
    for (int n = 0; n < kBandFilterTaps; n++) {
        const int m = n - kBandFilterDelay;
        const double ideal = m == 0 ? hi - lo : (std::sin(pi * hi * m) - std::sin(pi * lo * m)) / (pi * m);
        const double w = 0.5 - 0.5 * std::cos(2.0 * pi * (n + 1) / (kBandFilterTaps + 1));
        g[n] = (float)(ideal * w);
    }
}

extern "C" {

int rar_set_band_edges(rar_context *ctx, const float *edges_hz, int32_t bands, int32_t sample_rate) {
    RAR_ENTER(ctx);
    if (!edges_hz) {  // back to the default: equal-width bands
        ctx->band_edges.clear();
        ctx->band_G_bands = 0;
        return RAR_OK;
    }
    if (bands < 1 || bands > 128 || sample_rate <= 0) return fail(ctx, RAR_ERR_INVALID, "bands must be 1..128 and sample_rate positive");
    const double nyq = 0.5 * sample_rate;
    std::vector<float> e(bands + 1);
    for (int b = 0; b <= bands; b++) {
        e[b] = (float)(edges_hz[b] / nyq);
        if (!(e[b] >= 0.0f && e[b] <= 1.0f) || (b > 0 && !(e[b] > e[b - 1])))
            return fail(ctx, RAR_ERR_INVALID, "band edges must ascend strictly within [0, sample_rate / 2]");
    }
    if (e[0] != 0.0f || e[bands] != 1.0f)
        return fail(ctx, RAR_ERR_INVALID, "band edges must start at 0 and end at sample_rate / 2 (the bands tile the spectrum)");
    ctx->band_edges.swap(e);
    ctx->band_G_bands = 0;
    return RAR_OK;
}

int rar_synthesize_ir(rar_context *ctx, int32_t slot, float *out, int64_t n) {
    RAR_ENTER(ctx);
    Slot *S = get_slot(ctx, slot, false);
    if (!out || n < 0) return fail(ctx, RAR_ERR_INVALID, "bad output array");
    if (!S || !S->configured) return fail(ctx, RAR_ERR_STATE, "slot is not configured (call rar_ir_clear first)");
    int rc = check_convolvable(ctx, *S);
    if (rc != RAR_OK) return rc;
    const long long have = std::min<long long>(n, slot_ir_samples(*S));
    if (slot_ir_samples(*S) > 0) {
        RAR_CUDA(ctx, ctx->d_synth.reserve((size_t)slot_ir_samples(*S) + 1));
        rc = slot_response(ctx, *S, 1.0f, ctx->d_synth.p);
        if (rc != RAR_OK) return rc;
        if (have > 0)
            RAR_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_synth.p, (size_t)have * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n > have) std::memset(out + have, 0, (size_t)(n - have) * sizeof(float));
    return RAR_OK;
}

int rar_set_air_absorption(rar_context *ctx, const float *alpha_per_m, int32_t bands) {
    RAR_ENTER(ctx);
    if (!alpha_per_m || bands == 0) {
        ctx->air.clear();
        return RAR_OK;
    }
    if (bands < 2 || bands > 128) return fail(ctx, RAR_ERR_UNSUPPORTED, "bands must be 2..128");
    for (int b = 0; b < bands; b++)
        if (!(alpha_per_m[b] >= 0.0f && alpha_per_m[b] <= 1e6f)) return fail(ctx, RAR_ERR_INVALID, "air absorption must be finite and >= 0");
    ctx->air.assign(alpha_per_m, alpha_per_m + bands);
    return RAR_OK;
}

// ---- IR slots ---------------------------------------------------------------------------------------

int rar_ir_clear(rar_context *ctx, int32_t slot, int32_t impulse_length, int32_t bands) {
    RAR_ENTER(ctx);
    Slot *S = get_slot(ctx, slot, true);
    if (!S) return fail(ctx, RAR_ERR_INVALID, "slot index out of range");
    if (impulse_length < 0 || bands < 1) return fail(ctx, RAR_ERR_INVALID, "bad impulse_length/bands");
    const long long words = (long long)impulse_length * bands;
    if (words > S->cap_words) {
        if (S->d_hist && S->aliased)
            return fail(ctx, RAR_ERR_STATE, "the slot must grow, but rar_ir_device_ptr handed out its device address, which would "
                                            "dangle: configure the slot at its largest size first, or use another slot");
        if (S->d_hist) {
            RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(S->d_hist);
            S->d_hist = nullptr;
            S->cap_words = 0;
        }
        RAR_CUDA(ctx, cudaMalloc((void **)&S->d_hist, (size_t)(words + 16) * sizeof(long long)));
        S->cap_words = words + 16;
    }
    if (words > 0) RAR_CUDA(ctx, cudaMemsetAsync(S->d_hist, 0, (size_t)words * sizeof(long long), ctx->stream));
    S->impulse_length = impulse_length;
    S->bands = bands;
    S->time_stride = 1;
    S->configured = true;
    S->H_valid = false;
    return RAR_OK;
}

int rar_ir_read_fixed(rar_context *ctx, int32_t slot, int64_t *out, int64_t n) {
    RAR_ENTER(ctx);
    Slot *S = get_slot(ctx, slot, false);
    if (!out || n < 0) return fail(ctx, RAR_ERR_INVALID, "bad output array");
    const long long words = (S && S->configured) ? (long long)S->impulse_length * S->bands : 0;
    // An unconfigured slot reads as zeros ("contents defined at creation").
    const long long have = n < words ? n : words;
    if (have > 0)
        RAR_CUDA(ctx, cudaMemcpyAsync(out, S->d_hist, (size_t)have * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n > have) std::memset(out + have, 0, (size_t)(n - have) * sizeof(int64_t));
    return RAR_OK;
}

int rar_ir_read(rar_context *ctx, int32_t slot, float *out, int64_t n) {
    RAR_ENTER(ctx);
    Slot *S = get_slot(ctx, slot, false);
    if (!out || n < 0) return fail(ctx, RAR_ERR_INVALID, "bad output array");
    const long long words = (S && S->configured) ? (long long)S->impulse_length * S->bands : 0;
    const long long have = n < words ? n : words;
    if (have > 0) {
        RAR_CUDA(ctx, ctx->d_irf.reserve((size_t)have));
        RAR_CUDA(ctx, launch_fixed_to_float(S->d_hist, ctx->d_irf.p, have, 1.0f, ctx->stream));
        ctx->launches++;
        RAR_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_irf.p, (size_t)have * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n > have) std::memset(out + have, 0, (size_t)(n - have) * sizeof(float));
    return RAR_OK;
}

int rar_ir_read_begin(rar_context *ctx, int32_t slot, int64_t n, int32_t *ticket) {
    RAR_ENTER(ctx);
    if (!ticket) return fail(ctx, RAR_ERR_INVALID, "null ticket");
    *ticket = -1;
    if (n < 0 || n > 0x7fffffffLL) return fail(ctx, RAR_ERR_INVALID, "bad length");
    Slot *S = get_slot(ctx, slot, false);
    const long long words = (S && S->configured) ? (long long)S->impulse_length * S->bands : 0;
    const long long have = n < words ? n : words;
    const int id = acquire_ticket(ctx);
    if (id < 0) return id;
    Ticket &T = *ctx->tickets[id];
    T.out_len = (int)n;
    T.failed = false;
    RAR_CUDA(ctx, T.h_out.reserve((size_t)n + 1));
    RAR_CUDA(ctx, T.d_out.reserve((size_t)n + 1));
    if (n > have) std::memset(T.h_out.p + have, 0, (size_t)(n - have) * sizeof(float));  // an unconfigured slot reads as zeros
    // The read runs on its own stream behind everything enqueued so far, so that the next frame's work on ANOTHER slot
    // (the ping/pong of RayTraceManager.cs:212-218) does not queue behind the conversion and the copy.
    RAR_CUDA(ctx, cudaEventRecord(T.ready, ctx->stream));
    RAR_CUDA(ctx, cudaStreamWaitEvent(ctx->read_stream, T.ready, 0));
    if (have > 0) {
        RAR_CUDA(ctx, launch_fixed_to_float(S->d_hist, T.d_out.p, have, 1.0f, ctx->read_stream));
        ctx->launches++;
        RAR_CUDA(ctx, cudaMemcpyAsync(T.h_out.p, T.d_out.p, (size_t)have * sizeof(float), cudaMemcpyDeviceToHost, ctx->read_stream));
        if (!S->read_done) RAR_CUDA(ctx, cudaEventCreateWithFlags(&S->read_done, cudaEventDisableTiming));
        RAR_CUDA(ctx, cudaEventRecord(S->read_done, ctx->read_stream));
        S->read_pending = true;
    }
    RAR_CUDA(ctx, cudaEventRecord(T.done, ctx->read_stream));
    T.active = true;
    *ticket = id;
    return RAR_OK;
}

int rar_ir_read_end(rar_context *ctx, int32_t ticket, float *out, int64_t n) {
    if (n < 0 || n > 0x7fffffffLL) return fail(ctx, RAR_ERR_INVALID, "bad length");
    return rar_convolve_end(ctx, ticket, out, (int32_t)n);  // same ticket mechanics: wait, copy out, release
}

int rar_ir_write(rar_context *ctx, int32_t slot, const float *ir, int32_t impulse_length, int32_t bands) {
    RAR_ENTER(ctx);
    if (!ir && impulse_length > 0) return fail(ctx, RAR_ERR_INVALID, "null ir");
    int rc = rar_ir_clear(ctx, slot, impulse_length, bands);
    if (rc != RAR_OK) return rc;
    Slot *S = get_slot(ctx, slot, false);
    const long long words = (long long)impulse_length * bands;
    if (words > 0) {
        RAR_CUDA(ctx, ctx->d_irf.reserve((size_t)words));
        RAR_CUDA(ctx, cudaMemcpyAsync(ctx->d_irf.p, ir, (size_t)words * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        RAR_CUDA(ctx, launch_float_to_fixed(ctx->d_irf.p, S->d_hist, words, ctx->stream));
        ctx->launches++;
        RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return RAR_OK;
}

int rar_ir_device_ptr(rar_context *ctx, int32_t slot, void **device_ptr, int64_t *n_words) {
    RAR_ENTER(ctx);
    Slot *S = get_slot(ctx, slot, false);
    if (!S || !S->configured) return fail(ctx, RAR_ERR_STATE, "slot is not configured (call rar_ir_clear first)");
    if (device_ptr) *device_ptr = S->d_hist;
    if (n_words) *n_words = (int64_t)S->impulse_length * S->bands;
    S->H_valid = false;  // the caller may modify the histogram (all-reduce) at any later time:
    S->aliased = true;   // cached spectra of this slot are never trusted again
    return RAR_OK;
}

int rar_allreduce_slots(rar_context *const *ctxs, int32_t n, int32_t slot) {
    if (!ctxs || n <= 0 || !ctxs[0]) return fail(nullptr, RAR_ERR_INVALID, "bad context array");
    rar_context *root = ctxs[0];
    if (n > 16) return fail(root, RAR_ERR_UNSUPPORTED, "at most 16 contexts");
    Slot *S0 = get_slot(root, slot, false);
    if (!S0 || !S0->configured) return fail(root, RAR_ERR_STATE, "slot is not configured on context 0");
    const long long words = (long long)S0->impulse_length * S0->bands;
    if (n == 1) return RAR_OK;
    PeerHists peers;
    peers.n = 0;
    for (int i = 1; i < n; i++) {
        rar_context *c = ctxs[i];
        if (!c) return fail(root, RAR_ERR_INVALID, "null context in the array");
        Slot *S = get_slot(c, slot, false);
        if (!S || !S->configured || S->impulse_length != S0->impulse_length || S->bands != S0->bands)
            return fail(root, RAR_ERR_INVALID, "slot configuration differs between contexts");
        if (c->device != root->device) {
            int can_rp = 0, can_pr = 0;
            cudaDeviceCanAccessPeer(&can_rp, root->device, c->device);
            cudaDeviceCanAccessPeer(&can_pr, c->device, root->device);
            if (!can_rp || !can_pr) return fail(root, RAR_ERR_UNSUPPORTED, "devices cannot access each other's memory");
        }
        peers.p[peers.n++] = S->d_hist;
        S->H_valid = false;
    }
    S0->H_valid = false;
    // peer access root <-> others (idempotent)
    for (int i = 1; i < n; i++) {
        if (ctxs[i]->device == root->device) continue;
        cudaSetDevice(root->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[i]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(root, RAR_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
        cudaSetDevice(ctxs[i]->device);
        e = cudaDeviceEnablePeerAccess(root->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(root, RAR_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
    }
    // the root waits for every context's enqueued work
    std::vector<cudaEvent_t> evs(n, nullptr);
    cudaError_t e = cudaSuccess;
    for (int i = 1; i < n && e == cudaSuccess; i++) {
        cudaSetDevice(ctxs[i]->device);
        e = cudaEventCreateWithFlags(&evs[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(evs[i], ctxs[i]->stream);
    }
    cudaSetDevice(root->device);
    for (int i = 1; i < n && e == cudaSuccess; i++) e = cudaStreamWaitEvent(root->stream, evs[i], 0);
    if (e == cudaSuccess) e = launch_peer_reduce(S0->d_hist, peers, words, root->stream);
    if (e == cudaSuccess) root->launches++;
    // broadcast the total and make every context wait for its copy
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&evs[0], cudaEventDisableTiming);
    for (int i = 1; i < n && e == cudaSuccess; i++) {
        Slot *S = get_slot(ctxs[i], slot, false);
        e = cudaMemcpyPeerAsync(S->d_hist, ctxs[i]->device, S0->d_hist, root->device, (size_t)words * sizeof(long long), root->stream);
    }
    if (e == cudaSuccess) e = cudaEventRecord(evs[0], root->stream);
    for (int i = 1; i < n && e == cudaSuccess; i++) {
        cudaSetDevice(ctxs[i]->device);
        e = cudaStreamWaitEvent(ctxs[i]->stream, evs[0], 0);
    }
    for (int i = 0; i < n; i++)
        if (evs[i]) cudaEventDestroy(evs[i]);  // destruction is deferred until the event has completed
    cudaSetDevice(root->device);
    RAR_CUDA(root, e);
    return RAR_OK;
}

// ---- multi-process peer-memory all-reduce -------------------------------------------------------------

int rar_exchange_destroy(rar_context *ctx) {
    RAR_ENTER(ctx);
    Exchange &X = ctx->ex;
    if (!X.region) return RAR_OK;
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < kExMaxRanks; r++)
        if (X.opened[r]) cudaIpcCloseMemHandle(X.opened[r]);
    cudaFree(X.region);
    cudaGetLastError();
    X = Exchange();
    return RAR_OK;
}

int rar_exchange_create(rar_context *ctx, int64_t capacity_words, void *handle_out) {
    RAR_ENTER(ctx);
    if (capacity_words <= 0 || !handle_out) return fail(ctx, RAR_ERR_INVALID, "bad capacity/handle");
    int rc = rar_exchange_destroy(ctx);
    if (rc != RAR_OK) return rc;
    Exchange &X = ctx->ex;
    X.stage_bytes = (((size_t)capacity_words + 2) * sizeof(long long) + 255) & ~(size_t)255;
    size_t bytes = kExHeaderBytes + 4 * X.stage_bytes;
    bytes = (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);  // whole 2 MiB pages: an allocation of its own
    RAR_CUDA(ctx, cudaMalloc((void **)&X.region, bytes));
    X.region_bytes = bytes;
    X.cap_words = capacity_words;
    cudaError_t e = cudaMemsetAsync(X.region, 0, bytes, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    ExchangeHandle h;
    std::memset(&h, 0, sizeof h);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h.mem, X.region);
    if (e != cudaSuccess) {
        cudaFree(X.region);
        X = Exchange();
        RAR_CUDA(ctx, e);
    }
    // The handle names the allocation the driver carved the region from; ship the region's offset in it.  Without the
    // offset a peer would address the wrong bytes (silent corruption), so failing to learn it is an error.
    typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    unsigned long long base = 0;
    size_t span = 0;
    const bool have_range = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) == cudaSuccess && fn &&
                            ((range_fn)fn)(&base, &span, (unsigned long long)(uintptr_t)X.region) == 0 && base;
    cudaGetLastError();
    if (!have_range) {
        cudaFree(X.region);
        X = Exchange();
        return fail(ctx, RAR_ERR_UNSUPPORTED, "cannot query the address range of the exchange region (cuMemGetAddressRange): "
                                              "its offset inside the IPC allocation is unknown");
    }
    h.offset = (unsigned long long)(uintptr_t)X.region - base;
    h.cap_words = (unsigned long long)capacity_words;
    std::memcpy(handle_out, &h, sizeof h);
    return RAR_OK;
}

int rar_exchange_connect(rar_context *ctx, int32_t rank, int32_t world, const void *handles) {
    RAR_ENTER(ctx);
    Exchange &X = ctx->ex;
    if (!X.region) return fail(ctx, RAR_ERR_STATE, "rar_exchange_create has not been called");
    if (X.connected) return fail(ctx, RAR_ERR_STATE, "exchange is already connected (destroy it first)");
    if (world < 1 || world > kExMaxRanks || rank < 0 || rank >= world || !handles)
        return fail(ctx, RAR_ERR_INVALID, "bad rank/world/handles (at most 16 ranks)");
    const ExchangeHandle *hs = static_cast<const ExchangeHandle *>(handles);
    for (int r = 0; r < world; r++) {
        ExchangeHandle h;
        std::memcpy(&h, hs + r, sizeof h);
        if ((long long)h.cap_words != X.cap_words) {
            for (int k = 0; k < r; k++)  // leave nothing mapped behind a failed connect
                if (X.opened[k]) {
                    cudaIpcCloseMemHandle(X.opened[k]);
                    X.opened[k] = nullptr;
                }
            return fail(ctx, RAR_ERR_INVALID, "exchange capacity differs between ranks");
        }
        if (r == rank) {
            X.peer[r] = X.region;
            continue;
        }
        void *base = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&base, h.mem, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int k = 0; k < r; k++)
                if (X.opened[k]) {
                    cudaIpcCloseMemHandle(X.opened[k]);
                    X.opened[k] = nullptr;
                }
            std::string m = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
            return fail(ctx, e == cudaErrorPeerAccessUnsupported ? RAR_ERR_UNSUPPORTED : RAR_ERR_CUDA, "%s", m.c_str());
        }
        X.opened[r] = base;
        X.peer[r] = static_cast<char *>(base) + h.offset;
    }
    X.rank = rank;
    X.world = world;
    X.epoch = 0;
    X.connected = true;
    return RAR_OK;
}

int rar_exchange_allreduce(rar_context *ctx, int32_t slot, int32_t mode) {
    RAR_ENTER(ctx);
    Exchange &X = ctx->ex;
    if (!X.connected) return fail(ctx, RAR_ERR_STATE, "exchange is not connected");
    Slot *S = get_slot(ctx, slot, false);
    if (!S || !S->configured) return fail(ctx, RAR_ERR_STATE, "slot is not configured (call rar_ir_clear first)");
    if (mode < RAR_EXCHANGE_AUTO || mode > RAR_EXCHANGE_TWO_SHOT) return fail(ctx, RAR_ERR_INVALID, "bad exchange mode");
    const long long words = (long long)S->impulse_length * S->bands;
    if (words > X.cap_words) return fail(ctx, RAR_ERR_INVALID, "slot is larger than the exchange capacity");
    if (X.world == 1 || words == 0) return RAR_OK;
    ExchangeLaunch a;
    std::memset(&a, 0, sizeof a);
    a.hist = S->d_hist;
    a.n_vecs = (words + 1) / 2;
    a.slice_vecs = (a.n_vecs + X.world - 1) / X.world;
    a.rank = X.rank;
    a.world = X.world;
    a.epoch = ++X.epoch;
    a.two_shot = mode == RAR_EXCHANGE_TWO_SHOT || (mode == RAR_EXCHANGE_AUTO && words * 8 > (512 << 10));
    const size_t parity = a.epoch & 1u;
    for (int r = 0; r < X.world; r++) {
        a.flags[r] = reinterpret_cast<unsigned *>(X.peer[r]);
        a.stage_in[r] = reinterpret_cast<long long *>(X.peer[r] + kExHeaderBytes + parity * X.stage_bytes);
        a.stage_out[r] = reinterpret_cast<long long *>(X.peer[r] + kExHeaderBytes + (2 + parity) * X.stage_bytes);
    }
    a.status = reinterpret_cast<unsigned *>(X.region + kExFlagWords * sizeof(unsigned));
    a.timeout_ns = 5000000000ull;
    RAR_CUDA(ctx, launch_exchange_allreduce(a, ctx->stream));
    ctx->launches++;
    S->H_valid = false;
    return RAR_OK;
}

int rar_exchange_status(rar_context *ctx) {
    RAR_ENTER(ctx);
    Exchange &X = ctx->ex;
    if (!X.region) return fail(ctx, RAR_ERR_STATE, "rar_exchange_create has not been called");
    unsigned st = 0;
    RAR_CUDA(ctx, cudaMemcpyAsync(&st, X.region + kExFlagWords * sizeof(unsigned), sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (st) return fail(ctx, RAR_ERR_STATE, "a peer did not reach the exchange barrier within the time limit");
    return RAR_OK;
}

// ---- clip preparation -----------------------------------------------------------------------------------

int64_t rar_prepared_length(int64_t samples, int32_t clip_frequency, int32_t sample_rate) {
    if (samples <= 0 || clip_frequency <= 0 || sample_rate <= 0) return 0;
    if (clip_frequency == sample_rate) return samples;
    const float ratio = (float)clip_frequency / (float)sample_rate;  // RayTraceManager.cs:152
    volatile float q = (float)samples / ratio;                       // :153, rounded to binary32 before RoundToInt
    return (int64_t)std::nearbyintf(q);                              // Mathf.RoundToInt: half to even
}

static int prepare_clips_common(rar_context *ctx, const float *d_raw, int64_t samples, int32_t channels, int32_t clip_frequency,
                                int32_t sample_rate, int32_t n_clips, float *d_out, int64_t out_stride) {
    ClipPrep a;
    a.raw = d_raw;
    a.out = d_out;
    a.samples = samples;
    a.new_len = rar_prepared_length(samples, clip_frequency, sample_rate);
    a.out_stride = out_stride;
    a.channels = channels;
    a.n_clips = n_clips;
    a.resample = clip_frequency != sample_rate;
    a.ratio = (float)clip_frequency / (float)sample_rate;
    RAR_CUDA(ctx, launch_prepare_clips(a, ctx->stream, ctx->dev.sm_count));
    if (a.n_clips > 0 && a.new_len > 0) ctx->launches++;
    return RAR_OK;
}

static int check_clip_args(rar_context *ctx, const void *raw, int64_t samples, int32_t channels, int32_t clip_frequency,
                           int32_t sample_rate, int32_t n_clips, const void *out, int64_t out_stride) {
    if (samples < 0 || n_clips < 0 || channels < 1 || clip_frequency <= 0 || sample_rate <= 0)
        return fail(ctx, RAR_ERR_INVALID, "bad clip shape");
    if (samples > 0x7fffffffLL) return fail(ctx, RAR_ERR_UNSUPPORTED, "clips are limited to 2^31-1 samples (AudioClip.samples is an int)");
    const int64_t n = rar_prepared_length(samples, clip_frequency, sample_rate);
    if (n > 0x7fffffffLL) return fail(ctx, RAR_ERR_UNSUPPORTED, "the prepared length exceeds 2^31-1 samples (newLength is an int)");
    if (out_stride < n) return fail(ctx, RAR_ERR_INVALID, "out_stride is smaller than the prepared length");
    if (n_clips > 0 && ((samples > 0 && !raw) || (n > 0 && !out))) return fail(ctx, RAR_ERR_INVALID, "null clip array");
    return RAR_OK;
}

int rar_prepare_clips_device(rar_context *ctx, const void *d_raw, int64_t samples, int32_t channels, int32_t clip_frequency,
                             int32_t sample_rate, int32_t n_clips, void *d_out, int64_t out_stride) {
    RAR_ENTER(ctx);
    int rc = check_clip_args(ctx, d_raw, samples, channels, clip_frequency, sample_rate, n_clips, d_out, out_stride);
    if (rc != RAR_OK) return rc;
    if (((uintptr_t)d_raw & 7u) != 0 || ((uintptr_t)d_out & 3u) != 0)
        return fail(ctx, RAR_ERR_INVALID, "device clip arrays must be 8-byte (raw) and 4-byte (out) aligned");
    return prepare_clips_common(ctx, static_cast<const float *>(d_raw), samples, channels, clip_frequency, sample_rate, n_clips,
                                static_cast<float *>(d_out), out_stride);
}

int rar_prepare_clips(rar_context *ctx, const float *raw, int64_t samples, int32_t channels, int32_t clip_frequency,
                      int32_t sample_rate, int32_t n_clips, float *out, int64_t out_stride) {
    RAR_ENTER(ctx);
    int rc = check_clip_args(ctx, raw, samples, channels, clip_frequency, sample_rate, n_clips, out, out_stride);
    if (rc != RAR_OK) return rc;
    const int64_t n = rar_prepared_length(samples, clip_frequency, sample_rate);
    if (n_clips == 0 || n == 0) return RAR_OK;
    const size_t in_words = (size_t)n_clips * samples * channels, out_words = (size_t)n_clips * n;
    RAR_CUDA(ctx, ctx->d_clip_raw.reserve(in_words + 4));
    RAR_CUDA(ctx, ctx->d_clip_out.reserve(out_words));
    RAR_CUDA(ctx, cudaMemcpyAsync(ctx->d_clip_raw.p, raw, in_words * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    rc = prepare_clips_common(ctx, ctx->d_clip_raw.p, samples, channels, clip_frequency, sample_rate, n_clips, ctx->d_clip_out.p, n);
    if (rc != RAR_OK) return rc;
    RAR_CUDA(ctx, cudaMemcpy2DAsync(out, (size_t)out_stride * sizeof(float), ctx->d_clip_out.p, (size_t)n * sizeof(float),
                                    (size_t)n * sizeof(float), (size_t)n_clips, cudaMemcpyDeviceToHost, ctx->stream));
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RAR_OK;
}

// ---- trace ------------------------------------------------------------------------------------------

struct Interleave {
    int world = 1, rank = 0, shift = 0;
};
static int trace_frames_impl(rar_context *ctx, const rar_trace_params *params, int32_t slot, int32_t n_frames, Interleave il = Interleave());

int rar_trace(rar_context *ctx, const rar_trace_params *params, int32_t slot) { return trace_frames_impl(ctx, params, slot, 1); }

int rar_trace_interleaved(rar_context *ctx, const rar_trace_params *params, int32_t slot, int32_t rank, int32_t world, int32_t chunk_log2) {
    if (world < 1 || rank < 0 || rank >= world || chunk_log2 < 5 || chunk_log2 > 30)
        return fail(ctx, RAR_ERR_INVALID, "need 0 <= rank < world and chunks of 2^5 .. 2^30 thread ids");
    if (params && (params->ray_begin != 0 || params->ray_end != 0))
        return fail(ctx, RAR_ERR_INVALID, "rar_trace_interleaved shards the whole dispatch: ray_begin / ray_end must be 0");
    Interleave il;
    il.world = world;
    il.rank = rank;
    il.shift = chunk_log2;
    return trace_frames_impl(ctx, params, slot, 1, il);
}

int rar_trace_frames(rar_context *ctx, const rar_trace_params *params, int32_t slot, int32_t n_frames) {
    if (n_frames < 0) return fail(ctx, RAR_ERR_INVALID, "n_frames must be >= 0");
    if (n_frames == 0) return RAR_OK;
    return trace_frames_impl(ctx, params, slot, n_frames);
}

static int trace_frames_impl(rar_context *ctx, const rar_trace_params *params, int32_t slot, int32_t n_frames, Interleave il) {
    RAR_ENTER(ctx);
    int rc = check_trace_params(ctx, params);
    if (rc != RAR_OK) return rc;
    Slot *S = get_slot(ctx, slot, false);
    if (!S || !S->configured) return fail(ctx, RAR_ERR_STATE, "slot is not configured (call rar_ir_clear first)");
    if (S->impulse_length != params->impulse_length || S->bands != params->bands)
        return fail(ctx, RAR_ERR_INVALID, "params impulse_length/bands do not match the slot");
    TraceLaunch a;
    fill_launch(ctx, params, a);
    rc = attach_grid(ctx, params, a);
    if (rc != RAR_OK) return rc;
    a.hist = reinterpret_cast<unsigned long long *>(S->d_hist);
    a.n_frames = n_frames;
    if (il.world > 1) {
        // this rank's chunks: c = rank, rank + world, ... below ceil(total / chunk); the launch indexes them densely
        const long long total = a.ray_end, chunk = 1LL << il.shift;
        const long long chunks = (total + chunk - 1) / chunk;
        const long long mine = chunks > il.rank ? (chunks - il.rank + il.world - 1) / il.world : 0;
        a.cyc_world = il.world;
        a.cyc_rank = il.rank;
        a.cyc_shift = il.shift;
        a.cyc_total = total;
        a.ray_begin = 0;
        a.ray_end = mine << il.shift;
    }
    const bool count = (params->flags & RAR_FLAG_COUNT_TESTS) != 0;
    a.counters = count ? ctx->d_counters.p : nullptr;
    if (params->debug_ray_count > 0) {
        const int rows = params->debug_ray_count > 100 ? params->debug_ray_count : 100;
        const long long entries = (long long)rows * (params->max_bounce_count + 1);
        // The kernel re-initialises the row of every ray it traces, so the usual whole-dispatch frame needs no separate
        // clear of the buffer; a launch that does not cover all rows (a ray-range shard) clears it the old way.
        if ((size_t)entries > ctx->d_debug.cap) RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // growth frees a buffer in use
        RAR_CUDA(ctx, ctx->d_debug.reserve((size_t)entries));
        if (a.ray_begin > 0 || a.ray_end < rows || a.cyc_world > 1)
            RAR_CUDA(ctx, cudaMemsetAsync(ctx->d_debug.p, 0, (size_t)entries * sizeof(f4), ctx->stream));
        ctx->debug_entries = (int)entries;
        a.debug_rays = ctx->d_debug.p;
        a.debug_ray_count = params->debug_ray_count;
        a.debug_capacity = (int)entries;
    }
    if (params->bands > 1 && !ctx->air.empty() && (int)ctx->air.size() != params->bands)
        return fail(ctx, RAR_ERR_STATE, "rar_set_air_absorption was called for a different band count than this trace has");
    // A ray's path depends on the broadband material only, so a slot of more than 8 bands is filled by tracing the
    // same rays once per chunk of 8 bands (tests are counted for the first chunk only).
    for (int b0 = 0; b0 < params->bands; b0 += 8) {
        a.band_offset = b0;
        a.band_valid = params->bands - b0 < 8 ? params->bands - b0 : 8;
        a.p.air_on = (params->bands > 1 && !ctx->air.empty()) ? 1 : 0;
        for (int k = 0; k < 8; k++) a.p.air[k] = (a.p.air_on && b0 + k < params->bands) ? ctx->air[b0 + k] : 0.0f;
        if (b0 > 0) {
            a.counters = nullptr;
            a.debug_rays = nullptr;
        }
        int launched = 0;
        RAR_CUDA(ctx, launch_trace(a, count && b0 == 0, ctx->dev, ctx->stream, &launched));
        ctx->launches += launched;
    }
    S->H_valid = false;
    {   // the sample grid the slot's bins stand on (RaytraceOcclusion2D.compute:241-243: bin = (int)(t*SampleRate/WindowSize))
        const float d = params->time_divisor;
        const int stride = (d >= 1.0f && d <= 65536.0f && d == std::floor(d)) ? (int)d : 0;
        S->time_stride = stride;
    }
    return RAR_OK;
}

int rar_trace_listeners(rar_context *ctx, const rar_trace_params *params, const float *listeners_xy, int32_t n_listeners,
                        int32_t first_slot) {
    RAR_ENTER(ctx);
    if (!params) return fail(ctx, RAR_ERR_INVALID, "null params");
    if (n_listeners < 0 || (n_listeners > 0 && !listeners_xy)) return fail(ctx, RAR_ERR_INVALID, "bad listener array");
    if (first_slot < 0 || (long long)first_slot + n_listeners > kMaxSlots) return fail(ctx, RAR_ERR_INVALID, "slot range out of bounds");
    if (n_listeners == 0) return RAR_OK;
    if (n_listeners == 1 || params->bands != 1) {
        // one listener, or banded slots: plain single-listener traces
        for (int l = 0; l < n_listeners; l++) {
            rar_trace_params p = *params;
            p.listener_pos[0] = listeners_xy[2 * l];
            p.listener_pos[1] = listeners_xy[2 * l + 1];
            int rc = rar_trace(ctx, &p, first_slot + l);
            if (rc != RAR_OK) return rc;
        }
        return RAR_OK;
    }
    // Fused kernel: every ray is traced once and tested against all listeners.
    int rc = check_trace_params(ctx, params);
    if (rc != RAR_OK) return rc;
    std::vector<unsigned long long *> hists(n_listeners);
    std::vector<f2> pos(n_listeners);
    for (int l = 0; l < n_listeners; l++) {
        Slot *S = get_slot(ctx, first_slot + l, false);
        if (!S || !S->configured) return fail(ctx, RAR_ERR_STATE, "a listener slot is not configured (call rar_ir_clear first)");
        if (S->impulse_length != params->impulse_length || S->bands != 1)
            return fail(ctx, RAR_ERR_INVALID, "params impulse_length/bands do not match a listener slot");
        hists[l] = reinterpret_cast<unsigned long long *>(S->d_hist);
        pos[l] = f2{listeners_xy[2 * l], listeners_xy[2 * l + 1]};
        S->H_valid = false;
    }
    // positions and slot addresses go up through pinned staging on the context's stream; the call does not wait for them
    const size_t pos_bytes = (size_t)n_listeners * sizeof(f2), hist_bytes = (size_t)n_listeners * sizeof(unsigned long long *);
    if (ctx->listeners_uploaded) RAR_CUDA(ctx, cudaEventSynchronize(ctx->listeners_uploaded));
    else RAR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->listeners_uploaded, cudaEventDisableTiming));
    if ((size_t)n_listeners > ctx->d_listeners.cap || (size_t)n_listeners > ctx->d_listener_hists.cap)
        RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // growth frees arrays an enqueued launch may still read
    RAR_CUDA(ctx, ctx->d_listeners.reserve((size_t)n_listeners));
    RAR_CUDA(ctx, ctx->d_listener_hists.reserve((size_t)n_listeners));
    RAR_CUDA(ctx, ctx->h_listeners.reserve(pos_bytes + hist_bytes));
    std::memcpy(ctx->h_listeners.p, hists.data(), hist_bytes);              // 8-byte items first: keeps both parts aligned
    std::memcpy(ctx->h_listeners.p + hist_bytes, pos.data(), pos_bytes);
    RAR_CUDA(ctx, cudaMemcpyAsync(ctx->d_listener_hists.p, ctx->h_listeners.p, hist_bytes, cudaMemcpyHostToDevice, ctx->stream));
    RAR_CUDA(ctx, cudaMemcpyAsync(ctx->d_listeners.p, ctx->h_listeners.p + hist_bytes, pos_bytes, cudaMemcpyHostToDevice, ctx->stream));
    RAR_CUDA(ctx, cudaEventRecord(ctx->listeners_uploaded, ctx->stream));
    TraceLaunch a;
    fill_launch(ctx, params, a);
    rc = attach_grid(ctx, params, a);
    if (rc != RAR_OK) return rc;
    a.listeners = ctx->d_listeners.p;
    a.listener_hists = ctx->d_listener_hists.p;
    a.n_listeners = n_listeners;
    const bool count = (params->flags & RAR_FLAG_COUNT_TESTS) != 0;
    a.counters = count ? ctx->d_counters.p : nullptr;
    int launched = 0;
    RAR_CUDA(ctx, launch_trace(a, count, ctx->dev, ctx->stream, &launched));
    ctx->launches += launched;
    return RAR_OK;
}

int rar_trace_hits(rar_context *ctx, const rar_trace_params *params, rar_ray_info *hits, rar_hit_key *keys,
                   int64_t capacity, int64_t *count) {
    RAR_ENTER(ctx);
    int rc = check_trace_params(ctx, params);
    if (rc != RAR_OK) return rc;
    if (capacity < 0 || (capacity > 0 && !hits)) return fail(ctx, RAR_ERR_INVALID, "bad hit array");
    DevBuf<rar_ray_info> d_hits;
    DevBuf<rar_hit_key> d_keys;
    RAR_CUDA(ctx, d_hits.reserve((size_t)capacity + 1));
    if (keys) RAR_CUDA(ctx, d_keys.reserve((size_t)capacity + 1));
    unsigned long long *d_count = ctx->d_counters.p + 5;
    RAR_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), ctx->stream));
    TraceLaunch a;
    fill_launch(ctx, params, a);
    a.hits = d_hits.p;
    a.keys = keys ? d_keys.p : nullptr;
    a.hit_cap = capacity;
    a.hit_count = d_count;
    const bool cnt = (params->flags & RAR_FLAG_COUNT_TESTS) != 0;
    a.counters = cnt ? ctx->d_counters.p : nullptr;
    int launched = 0;
    cudaError_t e = launch_trace(a, cnt, ctx->dev, ctx->stream, &launched);
    ctx->launches += launched;
    unsigned long long produced = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&produced, d_count, sizeof produced, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    const long long stored = (long long)produced < capacity ? (long long)produced : capacity;
    if (e == cudaSuccess && stored > 0) {
        e = cudaMemcpy(hits, d_hits.p, (size_t)stored * sizeof(rar_ray_info), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && keys) e = cudaMemcpy(keys, d_keys.p, (size_t)stored * sizeof(rar_hit_key), cudaMemcpyDeviceToHost);
    }
    d_hits.release();
    d_keys.release();
    RAR_CUDA(ctx, e);
    if (count) *count = (int64_t)produced;
    return RAR_OK;
}

int rar_get_counters(rar_context *ctx, rar_counters *out, int32_t reset) {
    RAR_ENTER(ctx);
    if (!out) return fail(ctx, RAR_ERR_INVALID, "null out");
    unsigned long long v[5];
    RAR_CUDA(ctx, cudaMemcpyAsync(v, ctx->d_counters.p, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
    if (reset) RAR_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, sizeof v, ctx->stream));
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out->ray_bounces = v[0];
    out->nearest_tests = v[1];
    out->shadow_tests = v[2];
    out->direct_hits = v[3];
    out->nee_hits = v[4];
    return RAR_OK;
}

int rar_get_debug_rays(rar_context *ctx, float *out_xyzw, int64_t n_float4) {
    RAR_ENTER(ctx);
    if (!out_xyzw || n_float4 < 0) return fail(ctx, RAR_ERR_INVALID, "bad output array");
    const long long have = n_float4 < ctx->debug_entries ? n_float4 : ctx->debug_entries;
    if (have > 0)
        RAR_CUDA(ctx, cudaMemcpyAsync(out_xyzw, ctx->d_debug.p, (size_t)have * sizeof(f4), cudaMemcpyDeviceToHost, ctx->stream));
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_float4 > have) std::memset(out_xyzw + have * 4, 0, (size_t)(n_float4 - have) * sizeof(f4));
    return RAR_OK;
}

// ---- one-shot convolution ---------------------------------------------------------------------------

int rar_convolve_begin(rar_context *ctx, int32_t slot, const float *in, int32_t in_len, int32_t accum_count,
                       int32_t *ticket) {
    RAR_ENTER(ctx);
    if (!ticket) return fail(ctx, RAR_ERR_INVALID, "null ticket");
    *ticket = -1;
    if (in_len < 0 || (in_len > 0 && !in)) return fail(ctx, RAR_ERR_INVALID, "bad input array");
    Slot *S = get_slot(ctx, slot, false);
    if (!S || !S->configured) return fail(ctx, RAR_ERR_STATE, "slot is not configured (call rar_ir_clear first)");
    {
        int rc = check_convolvable(ctx, *S);
        if (rc != RAR_OK) return rc;
    }
    const int ir_len = (int)slot_ir_samples(*S);  // a banded slot is convolved through its filter-bank synthesis
    const int out_len = in_len + ir_len;  // AudioConvolve.compute:15

    const int id = acquire_ticket(ctx);
    if (id < 0) return id;
    Ticket &T = *ctx->tickets[id];
    T.out_len = out_len;
    T.failed = false;
    RAR_CUDA(ctx, T.h_out.reserve((size_t)out_len + 1));
    RAR_CUDA(ctx, T.d_out.reserve((size_t)out_len + 1));

    const bool zero = accum_count <= 0 || in_len == 0 || ir_len == 0;  // AudioConvolve.compute:30
    if (zero) {
        if (out_len > 0) RAR_CUDA(ctx, cudaMemsetAsync(T.d_out.p, 0, (size_t)out_len * sizeof(float), ctx->stream));
    } else {
        int rc = ensure_slot_spectra(ctx, *S);
        if (rc != RAR_OK) return rc;
        const int n_part = S->H_parts;
        const int n_xwin = (in_len + kBlock - 1) / kBlock + 1;
        const int n_out_blocks = (out_len + kBlock - 1) / kBlock;
        RAR_CUDA(ctx, T.h_in.reserve((size_t)in_len));
        RAR_CUDA(ctx, T.d_x.reserve((size_t)in_len));
        RAR_CUDA(ctx, T.d_X.reserve((size_t)n_xwin * kBlock));
        RAR_CUDA(ctx, T.d_Y.reserve((size_t)n_out_blocks * kBlock));
        std::memcpy(T.h_in.p, in, (size_t)in_len * sizeof(float));
        RAR_CUDA(ctx, cudaMemcpyAsync(T.d_x.p, T.h_in.p, (size_t)in_len * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        RAR_CUDA(ctx, launch_input_spectra(T.d_x.p, in_len, T.d_X.p, n_xwin, kBlock, ctx->stream));
        RAR_CUDA(ctx, launch_block_cmac(T.d_X.p, n_xwin, S->d_H, n_part, T.d_Y.p, n_out_blocks, kBlock, ctx->stream));
        const float scale = (1.0f / (float)accum_count) / (float)kBlock;
        RAR_CUDA(ctx, launch_output_blocks(T.d_Y.p, n_out_blocks, T.d_out.p, out_len, scale, kBlock, ctx->stream));
        ctx->launches += 3;
    }
    if (out_len > 0)
        RAR_CUDA(ctx, cudaMemcpyAsync(T.h_out.p, T.d_out.p, (size_t)out_len * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    RAR_CUDA(ctx, cudaEventRecord(T.done, ctx->stream));
    T.active = true;
    *ticket = id;
    return RAR_OK;
}

int rar_poll(rar_context *ctx, int32_t ticket) {
    RAR_ENTER(ctx);
    Ticket *T = get_ticket(ctx, ticket);
    if (!T || !T->active) return fail(ctx, RAR_ERR_INVALID, "unknown ticket");
    cudaError_t e = cudaEventQuery(T->done);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) {
        cudaGetLastError();
        return 0;
    }
    T->failed = true;
    T->active = false;  // released: the mirrors drop a failed request without calling convolve_end (RayTraceManager.cs:119)
    cudaGetLastError();
    return fail(ctx, RAR_ERR_CUDA, "ticket failed: %s", cudaGetErrorString(e));
}

int rar_convolve_end(rar_context *ctx, int32_t ticket, float *out, int32_t out_len) {
    RAR_ENTER(ctx);
    Ticket *T = get_ticket(ctx, ticket);
    if (!T || !T->active) return fail(ctx, RAR_ERR_INVALID, "unknown ticket");
    if (out_len < 0 || (out_len > 0 && !out)) return fail(ctx, RAR_ERR_INVALID, "bad output array");
    cudaError_t e = cudaEventSynchronize(T->done);
    T->active = false;
    RAR_CUDA(ctx, e);
    const int n = out_len < T->out_len ? out_len : T->out_len;
    if (n > 0) std::memcpy(out, T->h_out.p, (size_t)n * sizeof(float));
    if (out_len > n) std::memset(out + n, 0, (size_t)(out_len - n) * sizeof(float));
    return RAR_OK;
}

int rar_convolve(rar_context *ctx, int32_t slot, const float *in, int32_t in_len, int32_t accum_count, float *out,
                 int32_t out_len) {
    int32_t ticket = -1;
    int rc = rar_convolve_begin(ctx, slot, in, in_len, accum_count, &ticket);
    if (rc != RAR_OK) return rc;
    return rar_convolve_end(ctx, ticket, out, out_len);
}

// ---- streaming convolver ----------------------------------------------------------------------------

int rar_conv_create(rar_context *ctx, int32_t n_streams, int32_t block, int32_t max_ir_len, rar_convolver **out) {
    RAR_ENTER(ctx);
    if (!out) return fail(ctx, RAR_ERR_INVALID, "null out pointer");
    *out = nullptr;
    if (n_streams <= 0 || max_ir_len <= 0) return fail(ctx, RAR_ERR_INVALID, "n_streams and max_ir_len must be positive");
    if (block != kBlock) return fail(ctx, RAR_ERR_UNSUPPORTED, "block must be 256");
    rar_convolver *cv = new (std::nothrow) rar_convolver();
    if (!cv) return fail(ctx, RAR_ERR_NOMEM, "out of host memory");
    cv->ctx = ctx;
    cv->max_ir_len = max_ir_len;
    StreamConv &c = cv->c;
    c.n_streams = n_streams;
    c.block = block;
    c.n_part = (max_ir_len + block - 1) / block;
    // Split the partitions of a stream over several CTAs so that the grid is many waves deep.
    int want = (32 * ctx->dev.sm_count + n_streams - 1) / n_streams;
    int max_split = (c.n_part + 15) / 16;
    if (want > max_split) want = max_split;
    if (want < 1) want = 1;
    c.part_per_split = (c.n_part + want - 1) / want;
    c.n_split = (c.n_part + c.part_per_split - 1) / c.part_per_split;
    c.head = 0;
    const size_t spec = (size_t)n_streams * c.n_part * block;
    cudaError_t e = cv->H.reserve(spec);
    if (e == cudaSuccess) e = cv->fdl.reserve(spec);
    if (e == cudaSuccess) e = cv->partial.reserve((size_t)n_streams * c.n_split * block);
    if (e == cudaSuccess) e = cv->prev.reserve((size_t)n_streams * block);
    if (e == cudaSuccess) e = cv->d_in.reserve((size_t)n_streams * block);
    if (e == cudaSuccess) e = cv->d_out.reserve((size_t)n_streams * block);
    if (e == cudaSuccess) e = cv->d_irf.reserve((size_t)c.n_part * block);
    if (e == cudaSuccess) e = cv->h_ir.reserve((size_t)c.n_part * block);
    if (e == cudaSuccess) e = cudaMemsetAsync(cv->H.p, 0, spec * sizeof(float2), ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(cv->fdl.p, 0, spec * sizeof(float2), ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(cv->prev.p, 0, (size_t)n_streams * block * sizeof(float), ctx->stream);
    if (e != cudaSuccess) {
        cv->H.release(); cv->fdl.release(); cv->partial.release(); cv->prev.release();
        cv->d_in.release(); cv->d_out.release(); cv->d_irf.release(); cv->h_ir.release();
        delete cv;
        RAR_CUDA(ctx, e);
    }
    c.H = cv->H.p;
    c.fdl = cv->fdl.p;
    c.partial = cv->partial.p;
    c.prev = cv->prev.p;
    ctx->convolvers.push_back(cv);
    *out = cv;
    return RAR_OK;
}

int rar_conv_destroy(rar_convolver *cv) {
    if (!cv) return RAR_OK;
    rar_context *ctx = cv->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (size_t i = 0; i < ctx->convolvers.size(); i++)
        if (ctx->convolvers[i] == cv) {
            ctx->convolvers.erase(ctx->convolvers.begin() + i);
            break;
        }
    cv->H.release(); cv->fdl.release(); cv->partial.release(); cv->prev.release();
    cv->d_in.release(); cv->d_out.release(); cv->d_irf.release(); cv->h_ir.release();
    cv->H2.release(); cv->partial2.release(); cv->d_fade.release(); cv->d_list.release();
    cv->d_batch.release(); cv->h_batch.release(); cv->h_out.release();
    for (cudaEvent_t &e : cv->batch_done)
        if (e) cudaEventDestroy(e);
    delete cv;
    return RAR_OK;
}

// Where the spectra of `stream` go: the active table, or (fade) the table of pending cross-faded updates.
static int conv_target(rar_convolver *cv, int32_t stream, bool fade, float2 **H) {
    rar_context *ctx = cv->ctx;
    StreamConv &c = cv->c;
    if (cv->fade.empty()) cv->fade.assign(c.n_streams, 0);
    if (!fade) {
        cv->fade[stream] = 0;  // a hard set cancels a pending cross-fade of the stream
        *H = c.H + (size_t)stream * c.n_part * c.block;
        return RAR_OK;
    }
    RAR_CUDA(ctx, cv->H2.reserve((size_t)c.n_streams * c.n_part * c.block));
    RAR_CUDA(ctx, cv->partial2.reserve((size_t)c.n_streams * c.n_split * c.block));
    RAR_CUDA(ctx, cv->d_fade.reserve((size_t)c.n_streams));
    RAR_CUDA(ctx, cv->d_list.reserve((size_t)c.n_streams));
    cv->fade[stream] = 1;
    *H = cv->H2.p + (size_t)stream * c.n_part * c.block;
    return RAR_OK;
}

static int conv_set_ir_impl(rar_convolver *cv, int32_t stream, const float *ir, int32_t ir_len, float scale, bool fade) {
    if (!cv) return fail(nullptr, RAR_ERR_INVALID, "null convolver");
    rar_context *ctx = cv->ctx;
    RAR_ENTER(ctx);
    if (stream < 0 || stream >= cv->c.n_streams) return fail(ctx, RAR_ERR_INVALID, "stream index out of range");
    if (ir_len < 0 || ir_len > cv->max_ir_len || (ir_len > 0 && !ir)) return fail(ctx, RAR_ERR_INVALID, "bad ir array");
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // h_ir is reused between calls
    for (int i = 0; i < ir_len; i++) cv->h_ir.p[i] = ir[i] * scale;
    if (ir_len > 0)
        RAR_CUDA(ctx, cudaMemcpyAsync(cv->d_irf.p, cv->h_ir.p, (size_t)ir_len * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    float2 *H = nullptr;
    int rc = conv_target(cv, stream, fade, &H);
    if (rc != RAR_OK) return rc;
    RAR_CUDA(ctx, launch_ir_spectra(cv->d_irf.p, ir_len, H, cv->c.n_part, cv->c.block, ctx->stream));
    ctx->launches++;
    return RAR_OK;
}

static int conv_set_ir_from_slot_impl(rar_convolver *cv, int32_t stream, int32_t slot, int32_t accum_count, bool fade) {
    if (!cv) return fail(nullptr, RAR_ERR_INVALID, "null convolver");
    rar_context *ctx = cv->ctx;
    RAR_ENTER(ctx);
    if (stream < 0 || stream >= cv->c.n_streams) return fail(ctx, RAR_ERR_INVALID, "stream index out of range");
    Slot *S = get_slot(ctx, slot, false);
    if (!S || !S->configured) return fail(ctx, RAR_ERR_STATE, "slot is not configured");
    int rc = check_convolvable(ctx, *S);
    if (rc != RAR_OK) return rc;
    const long long ir_len = slot_ir_samples(*S);
    if (ir_len > cv->max_ir_len) return fail(ctx, RAR_ERR_INVALID, "slot IR is longer than max_ir_len");
    const float scale = accum_count > 0 ? 1.0f / (float)accum_count : 0.0f;
    float2 *H = nullptr;
    rc = conv_target(cv, stream, fade, &H);
    if (rc != RAR_OK) return rc;
    rc = slot_response(ctx, *S, scale, cv->d_irf.p);
    if (rc != RAR_OK) return rc;
    RAR_CUDA(ctx, launch_ir_spectra(cv->d_irf.p, (int)ir_len, H, cv->c.n_part, cv->c.block, ctx->stream));
    ctx->launches += 1;
    return RAR_OK;
}


// Responses of n consecutive streams in one call (config 5 loads 256 of them): staged through two pinned halves that
// alternate, so the host fills one while the other is on its way, with no stream synchronisation -- only an event
// wait when a half is needed again -- and one spectra launch per group instead of one per stream.
int rar_conv_set_irs(rar_convolver *cv, int32_t first_stream, int32_t n, const float *irs, int32_t ir_len, int64_t ir_stride,
                     float scale) {
    if (!cv) return fail(nullptr, RAR_ERR_INVALID, "null convolver");
    rar_context *ctx = cv->ctx;
    RAR_ENTER(ctx);
    StreamConv &c = cv->c;
    if (n < 0 || first_stream < 0 || (long long)first_stream + n > c.n_streams) return fail(ctx, RAR_ERR_INVALID, "stream range out of bounds");
    if (ir_len < 0 || ir_len > cv->max_ir_len || ir_stride < ir_len || (n > 0 && ir_len > 0 && !irs)) return fail(ctx, RAR_ERR_INVALID, "bad ir array");
    if (n == 0) return RAR_OK;
    if (cv->fade.empty()) cv->fade.assign(c.n_streams, 0);
    const size_t row = (size_t)c.n_part * c.block;                       // floats per staged response (zero padded)
    size_t group = (size_t)(16u << 20) / (row * sizeof(float));          // ~16 MB per staging half
    if (group < 1) group = 1;
    if (group > (size_t)n) group = (size_t)n;
    RAR_CUDA(ctx, cv->h_batch.reserve(2 * group * row));
    RAR_CUDA(ctx, cv->d_batch.reserve(2 * group * row));
    for (cudaEvent_t &e : cv->batch_done)
        if (!e) RAR_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int half = 0;
    for (int s0 = 0; s0 < n; s0 += (int)group, half ^= 1) {
        const int g = (int)std::min<size_t>(group, (size_t)(n - s0));
        float *h = cv->h_batch.p + (size_t)half * group * row, *d = cv->d_batch.p + (size_t)half * group * row;
        RAR_CUDA(ctx, cudaEventSynchronize(cv->batch_done[half]));       // completes at once for a never-recorded event
        for (int k = 0; k < g; k++) {
            const float *src = irs + (size_t)(s0 + k) * ir_stride;
            float *dst = h + (size_t)k * row;
            for (int i = 0; i < ir_len; i++) dst[i] = src[i] * scale;
            std::memset(dst + ir_len, 0, (row - ir_len) * sizeof(float));
            cv->fade[first_stream + s0 + k] = 0;                         // a hard set cancels a pending cross-fade
        }
        RAR_CUDA(ctx, cudaMemcpyAsync(d, h, (size_t)g * row * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        RAR_CUDA(ctx, launch_ir_spectra_batch(d, (long long)row, (int)row, c.H + (size_t)(first_stream + s0) * row, (long long)row, c.n_part,
                                              g, c.block, ctx->stream));
        RAR_CUDA(ctx, cudaEventRecord(cv->batch_done[half], ctx->stream));
        ctx->launches++;
    }
    return RAR_OK;
}

// The same from traced slots, all on the device: stream first_stream + k takes slot slots[k] scaled by
// 1 / accum_counts[k] (banded slots through their filter-bank synthesis, in groups of one launch).
int rar_conv_set_irs_from_slots(rar_convolver *cv, int32_t first_stream, int32_t n, const int32_t *slots, const int32_t *accum_counts) {
    if (!cv) return fail(nullptr, RAR_ERR_INVALID, "null convolver");
    rar_context *ctx = cv->ctx;
    RAR_ENTER(ctx);
    StreamConv &c = cv->c;
    if (n < 0 || first_stream < 0 || (long long)first_stream + n > c.n_streams) return fail(ctx, RAR_ERR_INVALID, "stream range out of bounds");
    if (n > 0 && (!slots || !accum_counts)) return fail(ctx, RAR_ERR_INVALID, "null slot / accum_count array");
    if (n == 0) return RAR_OK;
    if (cv->fade.empty()) cv->fade.assign(c.n_streams, 0);
    const Slot *S0 = get_slot(ctx, slots[0], false);
    if (!S0 || !S0->configured) return fail(ctx, RAR_ERR_STATE, "slot is not configured");
    for (int k = 0; k < n; k++) {
        Slot *S = get_slot(ctx, slots[k], false);
        if (!S || !S->configured) return fail(ctx, RAR_ERR_STATE, "slot is not configured");
        if (S->impulse_length != S0->impulse_length || S->bands != S0->bands || S->time_stride != S0->time_stride)
            return fail(ctx, RAR_ERR_INVALID, "the slots of one call must have the same shape");
    }
    int rc = check_convolvable(ctx, *S0);
    if (rc != RAR_OK) return rc;
    const long long ir_len = slot_ir_samples(*S0);
    if (ir_len > cv->max_ir_len) return fail(ctx, RAR_ERR_INVALID, "slot IR is longer than max_ir_len");
    const bool banded = !(S0->bands == 1 && S0->time_stride == 1);
    if (banded) {
        rc = ensure_band_filters(ctx, S0->bands);
        if (rc != RAR_OK) return rc;
    }
    const size_t row = (size_t)c.n_part * c.block;
    const int group = std::min(n, kBandSynthBatch);
    RAR_CUDA(ctx, cv->d_batch.reserve((size_t)group * row));             // stream order serialises the groups' use of it
    for (int s0 = 0; s0 < n; s0 += group) {
        const int g = std::min(group, n - s0);
        float *d = cv->d_batch.p;
        if (banded) {
            BandSynthBatch batch;
            std::memset(&batch, 0, sizeof batch);
            for (int k = 0; k < g; k++) {
                const int acc = accum_counts[s0 + k];
                batch.items[k] = BandSynthItem{get_slot(ctx, slots[s0 + k], false)->d_hist, d + (size_t)k * row,
                                               acc > 0 ? 1.0f / (float)acc : 0.0f, 0};
            }
            RAR_CUDA(ctx, cudaMemsetAsync(d, 0, (size_t)g * row * sizeof(float), ctx->stream));
            RAR_CUDA(ctx, synth_batch(ctx, batch, g, S0->impulse_length, S0->bands, S0->time_stride, (int)ir_len));
            ctx->launches++;
        } else {
            for (int k = 0; k < g; k++) {
                const int acc = accum_counts[s0 + k];
                RAR_CUDA(ctx, launch_fixed_to_float(get_slot(ctx, slots[s0 + k], false)->d_hist, d + (size_t)k * row, ir_len,
                                                    acc > 0 ? 1.0f / (float)acc : 0.0f, ctx->stream));
                ctx->launches++;
            }
        }
        RAR_CUDA(ctx, launch_ir_spectra_batch(d, (long long)row, (int)ir_len, c.H + (size_t)(first_stream + s0) * row, (long long)row,
                                              c.n_part, g, c.block, ctx->stream));
        ctx->launches++;
        for (int k = 0; k < g; k++) cv->fade[first_stream + s0 + k] = 0;
    }
    return RAR_OK;
}

int rar_conv_set_ir(rar_convolver *cv, int32_t stream, const float *ir, int32_t ir_len, float scale) {
    return conv_set_ir_impl(cv, stream, ir, ir_len, scale, false);
}
int rar_conv_set_ir_from_slot(rar_convolver *cv, int32_t stream, int32_t slot, int32_t accum_count) {
    return conv_set_ir_from_slot_impl(cv, stream, slot, accum_count, false);
}
int rar_conv_update_ir(rar_convolver *cv, int32_t stream, const float *ir, int32_t ir_len, float scale) {
    return conv_set_ir_impl(cv, stream, ir, ir_len, scale, true);
}
int rar_conv_update_ir_from_slot(rar_convolver *cv, int32_t stream, int32_t slot, int32_t accum_count) {
    return conv_set_ir_from_slot_impl(cv, stream, slot, accum_count, true);
}

int rar_conv_reset(rar_convolver *cv) {
    if (!cv) return fail(nullptr, RAR_ERR_INVALID, "null convolver");
    rar_context *ctx = cv->ctx;
    RAR_ENTER(ctx);
    const size_t spec = (size_t)cv->c.n_streams * cv->c.n_part * cv->c.block;
    RAR_CUDA(ctx, cudaMemsetAsync(cv->fdl.p, 0, spec * sizeof(float2), ctx->stream));
    RAR_CUDA(ctx, cudaMemsetAsync(cv->prev.p, 0, (size_t)cv->c.n_streams * cv->c.block * sizeof(float), ctx->stream));
    cv->c.head = 0;
    return RAR_OK;
}

int rar_conv_process_device(rar_convolver *cv, const float *d_in, float *d_out) {
    if (!cv) return fail(nullptr, RAR_ERR_INVALID, "null convolver");
    rar_context *ctx = cv->ctx;
    RAR_ENTER(ctx);
    if (!d_in || !d_out) return fail(ctx, RAR_ERR_INVALID, "null device array");
    int launched = 0;
    std::vector<int> list;
    for (size_t st = 0; st < cv->fade.size(); st++)
        if (cv->fade[st]) list.push_back((int)st);
    if (list.empty()) {
        RAR_CUDA(ctx, launch_stream_step(cv->c, d_in, d_out, ctx->stream, &launched));
    } else {
        // this block cross-fades the listed streams from their current spectra to the pending ones, which then
        // become current
        StreamConv &c = cv->c;
        RAR_CUDA(ctx, cudaMemcpyAsync(cv->d_fade.p, cv->fade.data(), cv->fade.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        RAR_CUDA(ctx, cudaMemcpyAsync(cv->d_list.p, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        RAR_CUDA(ctx, launch_stream_step_fade(c, cv->H2.p, cv->partial2.p, cv->d_fade.p, cv->d_list.p, (int)list.size(), d_in, d_out,
                                              ctx->stream, &launched));
        const size_t row = (size_t)c.n_part * c.block;
        for (int st : list)
            RAR_CUDA(ctx, cudaMemcpyAsync(c.H + st * row, cv->H2.p + st * row, row * sizeof(float2), cudaMemcpyDeviceToDevice, ctx->stream));
        cv->fade.assign(cv->fade.size(), 0);
    }
    ctx->launches += launched;
    cv->c.head = (cv->c.head + 1) % cv->c.n_part;
    return RAR_OK;
}

int rar_conv_process(rar_convolver *cv, const float *in, float *out) {
    if (!cv) return fail(nullptr, RAR_ERR_INVALID, "null convolver");
    rar_context *ctx = cv->ctx;
    RAR_ENTER(ctx);
    if (!in || !out) return fail(ctx, RAR_ERR_INVALID, "null array");
    const size_t bytes = (size_t)cv->c.n_streams * cv->c.block * sizeof(float);
    RAR_CUDA(ctx, cudaMemcpyAsync(cv->d_in.p, in, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = rar_conv_process_device(cv, cv->d_in.p, cv->d_out.p);
    if (rc != RAR_OK) return rc;
    RAR_CUDA(ctx, cudaMemcpyAsync(out, cv->d_out.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RAR_OK;
}

int rar_conv_process_to_ring(rar_convolver *cv, const float *in, rar_ring *const *rings, int64_t sample_offset) {
    if (!cv) return fail(nullptr, RAR_ERR_INVALID, "null convolver");
    rar_context *ctx = cv->ctx;
    RAR_ENTER(ctx);
    if (!in || !rings || sample_offset < 0) return fail(ctx, RAR_ERR_INVALID, "null array or negative sample offset");
    const size_t words = (size_t)cv->c.n_streams * cv->c.block;
    RAR_CUDA(ctx, cv->h_out.reserve(words));   // pinned: the block lands next to the (pinned) rings without a bounce
    RAR_CUDA(ctx, cudaMemcpyAsync(cv->d_in.p, in, words * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    int rc = rar_conv_process_device(cv, cv->d_in.p, cv->d_out.p);
    if (rc != RAR_OK) return rc;
    RAR_CUDA(ctx, cudaMemcpyAsync(cv->h_out.p, cv->d_out.p, words * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int st = 0; st < cv->c.n_streams; st++) {
        if (!rings[st]) continue;
        rc = rar_ring_push(rings[st], cv->h_out.p + (size_t)st * cv->c.block, cv->c.block, sample_offset);
        if (rc != RAR_OK) return fail(ctx, rc, "rar_ring_push failed");
    }
    return RAR_OK;
}

int64_t rar_conv_bytes_per_block(const rar_convolver *cv) {
    if (!cv) return 0;
    const StreamConv &c = cv->c;
    const int64_t row = (int64_t)c.block * sizeof(float2);
    // per stream: P delay-line spectra + P IR spectra read, 1 spectrum written, partial sums written and
    // read back, block in + block out + previous-block read/write.
    const int64_t per_stream = 2 * (int64_t)c.n_part * row + row + 2 * (int64_t)c.n_split * row + 4 * (int64_t)c.block * 4;
    return per_stream * c.n_streams;
}

// ---- measurement helpers ----------------------------------------------------------------------------

int rar_device_info(rar_context *ctx, int32_t *sm_count, int32_t *sm_clock_khz, int32_t *smem_optin_bytes) {
    RAR_ENTER(ctx);
    if (sm_count) *sm_count = ctx->dev.sm_count;
    if (sm_clock_khz) *sm_clock_khz = ctx->dev.sm_clock_khz;
    if (smem_optin_bytes) *smem_optin_bytes = ctx->dev.smem_optin;
    return RAR_OK;
}

int rar_measure_fp32_peak(rar_context *ctx, double *lane_ops_per_s) {
    RAR_ENTER(ctx);
    if (!lane_ops_per_s) return fail(ctx, RAR_ERR_INVALID, "null out");
    DevBuf<float> sink;
    RAR_CUDA(ctx, sink.reserve(4));
    const int blocks = ctx->dev.sm_count * 8, threads = 256, iters = 4000;
    cudaEvent_t e0, e1;
    RAR_CUDA(ctx, cudaEventCreate(&e0));
    RAR_CUDA(ctx, cudaEventCreate(&e1));
    cudaError_t e = launch_fp32_peak(sink.p, blocks, threads, iters / 4, ctx->stream);  // warm-up
    double best_ms = 1e30;
    for (int rep = 0; rep < 3 && e == cudaSuccess; rep++) {
        cudaEventRecord(e0, ctx->stream);
        e = launch_fp32_peak(sink.p, blocks, threads, iters, ctx->stream);
        cudaEventRecord(e1, ctx->stream);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        float ms = 0;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (ms > 0 && ms < best_ms) best_ms = ms;
    }
    ctx->launches += 4;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    sink.release();
    RAR_CUDA(ctx, e);
    const double ops = (double)blocks * threads * (double)iters * 16.0 * 8.0;
    *lane_ops_per_s = ops / (best_ms * 1e-3);
    return RAR_OK;
}

int rar_debug_grid(rar_context *ctx, int32_t *nx, int32_t *ny, int64_t *n_items, uint64_t *digest) {
    RAR_ENTER(ctx);
    if (ctx->n_walls < 0) return fail(ctx, RAR_ERR_STATE, "rar_set_walls has not been called");
    rar_trace_params p;
    std::memset(&p, 0, sizeof p);
    p.flags = RAR_FLAG_USE_GRID;
    TraceLaunch a;
    std::memset(&a, 0, sizeof a);
    int rc = attach_grid(ctx, &p, a);
    if (rc != RAR_OK) return rc;
    const GridFrame &g = ctx->grid_frame_;
    if (nx) *nx = g.nx;
    if (ny) *ny = g.ny;
    if (n_items) *n_items = 0;
    if (digest) *digest = 0;
    if (!a.use_grid) return RAR_OK;
    const size_t n_cells = (size_t)g.nx * g.ny;
    std::vector<uint32_t> start(n_cells + 1);
    RAR_CUDA(ctx, cudaMemcpyAsync(start.data(), ctx->d_grid_start.p, (n_cells + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    RAR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<uint32_t> items(start[n_cells]);
    if (!items.empty()) RAR_CUDA(ctx, cudaMemcpy(items.data(), ctx->d_grid_items.p, items.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (n_items) *n_items = (int64_t)items.size();
    if (digest) *digest = grid_digest(start.data(), start.size(), items.data(), items.size());
    return RAR_OK;
}

int rar_selftest_arithmetic(rar_context *ctx, int64_t n_samples, uint32_t seed, uint64_t *mismatches) {
    RAR_ENTER(ctx);
    if (n_samples < 0 || !mismatches) return fail(ctx, RAR_ERR_INVALID, "bad selftest arguments");
    DevBuf<unsigned long long> d;
    RAR_CUDA(ctx, d.reserve(8));
    cudaError_t e = cudaMemsetAsync(d.p, 0, 8 * sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess && n_samples > 0) {
        e = launch_arithmetic_selftest(n_samples, seed, d.p, ctx->dev.sm_count * 8, ctx->stream);
        ctx->launches++;
    }
    unsigned long long h[5] = {0, 0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    d.release();
    RAR_CUDA(ctx, e);
    for (int k = 0; k < 5; k++) mismatches[k] = h[k];
    return RAR_OK;
}

int64_t rar_launch_count(const rar_context *ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

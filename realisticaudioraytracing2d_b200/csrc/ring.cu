// ring.cu -- the playback ring of the streaming path as a native, lock-free single-producer / single-consumer
// structure.  Host code only; it lives in the CUDA library because its storage is pinned host memory
// (so device-to-host copies can land next to it without a bounce) and because rar_conv_process_to_ring feeds it.
//
// Reference: Assets/Script/AudioManager.cs.  `PushSamples` (:45-54, Unity main thread) overlap-adds a convolved chunk
// at an absolute sample offset; `OnAudioFilterRead` (:56-69, audio thread) hands out the samples at the read head and
// zeroes them.  The reference serialises the two with `lock (bufferLock)`.  An audio callback must not block on a
// lock the main thread may hold for a 76 800-sample chunk, so here every ring element is an atomic word instead:
//   consumer:  s = exchange(element, 0)                       -- wait-free
//   producer:  element = element + sample via compare-exchange -- retries only if the consumer zeroed that very
//              element in between (at most once per lap), so lock-free
// Element for element this is what the lock gives (a read-and-zero and a += never interleave inside one element).
// What the lock adds -- a whole chunk being pushed before or after a whole callback -- only matters when a push
// lands ON the read head, i.e. when the chunk is late, and then both versions play samples a lap late.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/rar2d.h"

struct rar_ring {
    std::atomic<uint32_t> *cell = nullptr;  // float bits
    int32_t size = 0;
    bool pinned = false;
    std::atomic<int32_t> streaming{0};
    // consumer-owned; published so that a producer can pace itself
    int32_t read_head = 0;
    std::atomic<int64_t> frames_drained{0};
};

static_assert(sizeof(std::atomic<uint32_t>) == sizeof(float), "ring cells are plain 32-bit words");

namespace {
inline float bits_to_float(uint32_t b) {
    float f;
    std::memcpy(&f, &b, sizeof f);
    return f;
}
inline uint32_t float_to_bits(float f) {
    uint32_t b;
    std::memcpy(&b, &f, sizeof b);
    return b;
}
}  // namespace

extern "C" {

int rar_ring_create(int32_t output_sample_rate, float reverb_duration, rar_ring **out) {
    if (!out) return RAR_ERR_INVALID;
    *out = nullptr;
    if (output_sample_rate <= 0 || !(reverb_duration >= 0.0f)) return RAR_ERR_INVALID;
    // AudioManager.cs:30  bufferSize = Mathf.CeilToInt(sampleRate * (reverbDuration + 1f)), in binary32
    const volatile float span = reverb_duration + 1.0f;
    const volatile float prod = (float)output_sample_rate * span;
    const double size_d = std::ceil((double)prod);
    if (!(size_d >= 1.0 && size_d <= 1e9)) return RAR_ERR_INVALID;
    rar_ring *r = new (std::nothrow) rar_ring();
    if (!r) return RAR_ERR_NOMEM;
    r->size = (int32_t)size_d;
    void *mem = nullptr;
    const size_t bytes = (size_t)r->size * sizeof(uint32_t);
    if (cudaHostAlloc(&mem, bytes, cudaHostAllocPortable) == cudaSuccess) {
        r->pinned = true;
    } else {
        cudaGetLastError();  // no device / no driver: the ring is host logic and works from ordinary memory too
        mem = std::malloc(bytes);
        if (!mem) {
            delete r;
            return RAR_ERR_NOMEM;
        }
    }
    r->cell = static_cast<std::atomic<uint32_t> *>(mem);
    for (int32_t i = 0; i < r->size; i++) new (&r->cell[i]) std::atomic<uint32_t>(0u);
    r->streaming.store(1, std::memory_order_release);  // StartStreaming: isStreaming = true (:34)
    *out = r;
    return RAR_OK;
}

int rar_ring_destroy(rar_ring *r) {
    if (!r) return RAR_OK;
    if (r->cell) {
        if (r->pinned) cudaFreeHost(r->cell);
        else std::free(r->cell);
    }
    delete r;
    return RAR_OK;
}

int32_t rar_ring_size(const rar_ring *r) { return r ? r->size : 0; }
int32_t rar_ring_is_pinned(const rar_ring *r) { return r && r->pinned ? 1 : 0; }
int64_t rar_ring_frames_drained(const rar_ring *r) { return r ? r->frames_drained.load(std::memory_order_acquire) : 0; }

// AudioManager.StartStreaming on an existing ring (:26-36): silence, read head at 0, streaming.  Call it while the
// consumer is not inside rar_ring_drain (the reference swaps the whole array under the audio thread's feet, unlocked).
int rar_ring_reset(rar_ring *r) {
    if (!r) return RAR_ERR_INVALID;
    r->streaming.store(0, std::memory_order_release);
    for (int32_t i = 0; i < r->size; i++) r->cell[i].store(0u, std::memory_order_relaxed);
    r->read_head = 0;
    r->frames_drained.store(0, std::memory_order_release);
    r->streaming.store(1, std::memory_order_release);
    return RAR_OK;
}

// AudioManager.StopStreaming (:38-43): pushes and drains become no-ops.
int rar_ring_stop(rar_ring *r) {
    if (!r) return RAR_ERR_INVALID;
    r->streaming.store(0, std::memory_order_release);
    return RAR_OK;
}

// AudioManager.PushSamples (:45-54).  Producer side: one thread at a time.
int rar_ring_push(rar_ring *r, const float *samples, int32_t n, int64_t sample_offset) {
    if (!r || n < 0 || (n > 0 && !samples) || sample_offset < 0) return RAR_ERR_INVALID;
    if (!r->streaming.load(std::memory_order_acquire)) return RAR_OK;  // :47
    const int32_t size = r->size;
    int32_t pos = (int32_t)(sample_offset % size);  // :49
    for (int32_t i = 0; i < n; i++) {
        std::atomic<uint32_t> &c = r->cell[pos];
        uint32_t old = c.load(std::memory_order_relaxed);
        while (!c.compare_exchange_weak(old, float_to_bits(bits_to_float(old) + samples[i]), std::memory_order_relaxed,
                                        std::memory_order_relaxed)) {
        }
        if (++pos == size) pos = 0;  // (writePos + i) % bufferSize (:51): a chunk longer than the ring laps and accumulates
    }
    return RAR_OK;
}

// AudioManager.OnAudioFilterRead (:56-69).  Consumer side: one thread, never blocks, never allocates.
int rar_ring_drain(rar_ring *r, float *data, int32_t data_length, int32_t channels) {
    if (!r || data_length < 0 || channels < 1 || (data_length > 0 && !data)) return RAR_ERR_INVALID;
    if (!r->streaming.load(std::memory_order_acquire)) return RAR_OK;  // :58: data is left as it was
    const int32_t frames = data_length / channels, size = r->size;
    int32_t head = r->read_head;
    for (int32_t i = 0; i < frames; i++) {
        const float s = bits_to_float(r->cell[head].exchange(0u, std::memory_order_relaxed));  // :63-64
        if (++head == size) head = 0;                                                            // :65
        for (int32_t c = 0; c < channels; c++) data[(size_t)i * channels + c] = s;              // :66
    }
    r->read_head = head;
    r->frames_drained.fetch_add(frames, std::memory_order_release);
    return RAR_OK;
}

}  // extern "C"

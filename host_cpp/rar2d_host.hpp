// rar2d_host.hpp -- C++ host mirror of the reference's C# component surface for the hot path, over the
// C-ABI of include/rar2d.h.
//
// The reference's host side is C# (Unity MonoBehaviours).  No C# toolchain exists in the build image, so
// besides the uncompiled P/Invoke files under csharp/ this header gives a COMPILED host layer in C++ with
// the reference's names, fields and call order:
//   Segment / AudioMat                     Assets/Script/Helpers/SceneHelper.cs:8-22
//   SceneToData2D::GetSegmentsFromColliders Assets/Script/Helpers/SceneHelper.cs:29-110
//   AudioManager                           Assets/Script/AudioManager.cs:5-71
//   RayTraceManager                        Assets/Script/RayTraceManager.cs:8-281
// Unity's frame loop is replaced by the caller invoking Start()/Update()/FixedUpdate().  Header-only; link
// with librar2d.so.  tests/host_cpp_driver.cpp exercises it on the GPU box.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/rar2d.h"

namespace rar2d_host {

struct Vector2 { float x = 0, y = 0; };

struct AudioMaterial {  // AudioMaterial.cs:6-20
    float absorption = 0.1f, scattering = 0.5f, transmission = 0.0f, ior = 1.0f;
};

using Segment = rar_segment;  // 40 bytes, LayoutKind.Sequential

struct Transform {  // world position, rotation about z as quaternion (0,0,z,w), lossy scale
    Vector2 position;
    float qz = 0.0f, qw = 1.0f;
    Vector2 lossyScale{1.0f, 1.0f};
    Vector2 TransformPoint(Vector2 p) const {
        const float r00 = 1.0f - 2.0f * (qz * qz), r01 = -(2.0f * (qz * qw)), r10 = 2.0f * (qz * qw), r11 = r00;
        const float lx = p.x * lossyScale.x, ly = p.y * lossyScale.y;
        volatile float x = r00 * lx + r01 * ly, y = r10 * lx + r11 * ly;  // volatile: no contraction into the add
        return Vector2{x + position.x, y + position.y};
    }
};

struct GameObject {  // a collider + its AcousticSurface.material
    enum Kind { None, Box, Circle, Polygon } kind = None;
    bool enabled = true;
    Transform transform;
    Vector2 size{1.0f, 1.0f}, offset{0.0f, 0.0f};  // BoxCollider2D / CircleCollider2D.offset
    float radius = 0.5f;                            // CircleCollider2D
    std::vector<std::vector<Vector2>> paths;        // PolygonCollider2D
    bool hasSurface = true;
    AudioMaterial material;
};

struct SceneToData2D {
    static constexpr int CIRCLE_RESOLUTION = 32;  // Helpers/SceneHelper.cs:26

    static std::vector<Segment> GetSegmentsFromColliders(const std::vector<GameObject> &objects) {
        std::vector<Segment> all;
        for (const GameObject &obj : objects) {
            if (obj.kind == GameObject::None || !obj.enabled) continue;  // :34
            if (!obj.hasSurface) throw std::runtime_error("GameObject has no AcousticSurface.material");  // :102-103
            if (obj.kind == GameObject::Polygon) {
                for (const auto &path : obj.paths) AddLoopToSegments(obj.transform, path, all, obj.material);
            } else if (obj.kind == GameObject::Box) {
                const float hx = obj.size.x * 0.5f, hy = obj.size.y * 0.5f, ox = obj.offset.x, oy = obj.offset.y;
                AddLoopToSegments(obj.transform, {{ox - hx, oy - hy}, {ox + hx, oy - hy}, {ox + hx, oy + hy}, {ox - hx, oy + hy}}, all,
                                  obj.material);
            } else {
                std::vector<Vector2> pts;
                for (int i = 0; i < CIRCLE_RESOLUTION; i++) {
                    const float angle = ((float)i / (float)CIRCLE_RESOLUTION) * 3.14159274f * 2.0f;
                    pts.push_back({obj.offset.x + (float)std::cos((double)angle) * obj.radius,
                                   obj.offset.y + (float)std::sin((double)angle) * obj.radius});
                }
                AddLoopToSegments(obj.transform, pts, all, obj.material);
            }
        }
        return all;
    }

    static void AddLoopToSegments(const Transform &t, const std::vector<Vector2> &local, std::vector<Segment> &out,
                                  const AudioMaterial &m) {
        const float winding = (t.lossyScale.x * t.lossyScale.y) >= 0.0f ? 1.0f : -1.0f;  // Mathf.Sign (:81)
        for (size_t i = 0; i < local.size(); i++) {
            const Vector2 a = t.TransformPoint(local[i]), b = t.TransformPoint(local[(i + 1) % local.size()]);
            float dx = b.x - a.x, dy = b.y - a.y;
            volatile float sq = dx * dx;
            volatile float sq2 = dy * dy;
            const float mag = std::sqrt(sq + sq2);
            if (mag > 1e-5f) { dx /= mag; dy /= mag; } else { dx = 0.0f; dy = 0.0f; }  // Vector2.normalized
            Segment s;
            s.start[0] = a.x; s.start[1] = a.y; s.end[0] = b.x; s.end[1] = b.y;
            s.normal[0] = dy * winding; s.normal[1] = -dx * winding;  // (:93)
            s.absorption = m.absorption; s.scattering = m.scattering; s.transmission = m.transmission; s.ior = m.ior;
            out.push_back(s);
        }
    }
};

class AudioManager {  // AudioManager.cs
public:
    float chunkDuration = 0.1f;
    explicit AudioManager(int outputSampleRate = 48000) : sampleRate(outputSampleRate) {}
    bool IsStreaming() const { return isStreaming; }
    void StartStreaming(float reverbDuration) {  // :26-36
        if (isStreaming) StopStreaming();
        bufferSize = (int)std::ceil((float)sampleRate * (reverbDuration + 1.0f));
        ringBuffer.assign(bufferSize, 0.0f);
        readHead = 0;
        isStreaming = true;
    }
    void StopStreaming() { isStreaming = false; }  // :38-43
    void PushSamples(const std::vector<float> &samples, int sampleOffset) {  // :45-54
        if (!isStreaming || ringBuffer.empty()) return;
        std::lock_guard<std::mutex> lock(bufferLock);
        const int writePos = sampleOffset % bufferSize;
        for (size_t i = 0; i < samples.size(); i++) ringBuffer[(writePos + i) % bufferSize] += samples[i];
    }
    void OnAudioFilterRead(float *data, int length, int channels) {  // :56-69
        if (!isStreaming || ringBuffer.empty()) return;
        std::lock_guard<std::mutex> lock(bufferLock);
        for (int i = 0; i < length / channels; i++) {
            const float s = ringBuffer[readHead];
            ringBuffer[readHead] = 0.0f;
            readHead = (readHead + 1) % bufferSize;
            for (int c = 0; c < channels; c++) data[i * channels + c] = s;
        }
    }
    std::vector<float> ringBuffer;

private:
    int readHead = 0, sampleRate, bufferSize = 0;
    std::mutex bufferLock;
    bool isStreaming = false;
};

struct AudioClip {
    std::vector<float> data;  // interleaved
    int channels = 1, frequency = 48000;
    int samples() const { return (int)data.size() / channels; }
};

class RayTraceManager {  // RayTraceManager.cs
public:
    // [Header("Simulation")] :12-16
    int rayCount = 1000, maxBounces = 5;
    float speedOfSound = 343.0f;
    bool dynamicObstacles = false;
    // [Header("Audio")] :18-24
    const AudioClip *inputClip = nullptr;
    AudioManager *audioManager = nullptr;
    int sampleRate = 48000;
    float inputGain = 1.0f, reverbDuration = 2.0f;
    bool loop = true;
    // [Header("Scene")] :26-29
    const Transform *source = nullptr, *listener = nullptr;
    float listenerRadius = 0.5f;
    std::vector<GameObject> obstacleObjects;
    // [Header("Debug")] :31-34
    int debugRayCount = 100;
    // not in the reference
    int gridThreshold = 64;       // wall count from which traces use RAR_FLAG_USE_GRID (identical results)
    float fixedDeltaTime = 0.02f; // Time.fixedDeltaTime
    int frameCount = 0;           // Time.frameCount
    int accumFrames = 0;

    explicit RayTraceManager(int device = 0) {
        if (rar_create(device, &ctx) != RAR_OK) throw std::runtime_error(std::string("rar_create: ") + rar_last_error(nullptr));
    }
    ~RayTraceManager() { OnDestroy(); }
    RayTraceManager(const RayTraceManager &) = delete;
    RayTraceManager &operator=(const RayTraceManager &) = delete;

    void Start() { UpdateGeometry(); }  // :45-48

    void Update() {  // :50-62
        frameCount++;
        if (!source || !listener) return;
        RunSimulation();
        PollChunks();  // Unity resumes `yield return null` coroutines once per frame
    }

    void FixedUpdate() {  // :64-89
        if (!audioManager || !audioManager->IsStreaming()) return;
        if (dynamicObstacles) UpdateGeometry();
        samplesSinceLastChunk += (int)std::nearbyint(fixedDeltaTime * (float)sampleRate);
        if (samplesSinceLastChunk < chunkSamples) return;
        if (nextStreamingOffset >= (int)fullInputSamples.size()) {
            if (loop) nextStreamingOffset = 0; else audioManager->StopStreaming();
        }
        if (!audioManager->IsStreaming()) return;
        ProcessChunk(nextStreamingOffset, chunkSamples, std::max(1, accumFrames), GetActiveIRBuffer());
        activeIRIndex = 1 - activeIRIndex;
        nextStreamingOffset += chunkSamples;
        ResetIR();
        samplesSinceLastChunk -= chunkSamples;
    }

    void StartStreaming() {  // :125-133
        nextStreamingOffset = 0;
        samplesSinceLastChunk = 0;
        chunkSamples = (int)std::nearbyint((float)sampleRate * audioManager->chunkDuration);
        fullInputSamples = LoadSample(*inputClip);
        ResetIR();
        audioManager->StartStreaming(reverbDuration);
    }

    std::vector<float> LoadSample(const AudioClip &clip) const {  // :135-167
        const int n = clip.samples(), ch = clip.channels;
        std::vector<float> mono(n);
        for (int i = 0; i < n; i++) {
            float sum = 0;
            for (int c = 0; c < ch; c++) sum += clip.data[(size_t)i * ch + c];
            mono[i] = sum / ch;
        }
        if (clip.frequency == sampleRate) return mono;
        const float ratio = (float)clip.frequency / (float)sampleRate;
        const int newLength = (int)std::nearbyint((float)n / ratio);
        std::vector<float> out(newLength);
        for (int i = 0; i < newLength; i++) {
            const float srcIdx = (float)i * ratio;
            const int i0 = (int)std::floor(srcIdx), i1 = std::min(i0 + 1, n - 1);
            const float t = std::min(1.0f, std::max(0.0f, srcIdx - (float)i0));
            out[i] = mono[i0] + (mono[i1] - mono[i0]) * t;  // Mathf.Lerp
        }
        return out;
    }

    void ResetIR() {  // :169-177
        accumFrames = 0;
        const int slot = GetActiveIRBuffer();
        Check(rar_ir_clear(ctx, slot, IrLength(), 1), "rar_ir_clear");
        slotLength[slot] = IrLength();
    }

    void RunSimulation() {  // :179-210 fused with :220-233
        if (!wallsUploaded) UpdateGeometry();
        rar_trace_params p;
        std::memset(&p, 0, sizeof p);
        p.source_pos[0] = source->position.x; p.source_pos[1] = source->position.y;
        p.listener_pos[0] = listener->position.x; p.listener_pos[1] = listener->position.y;
        p.listener_radius = listenerRadius; p.speed_of_sound = speedOfSound; p.input_gain = inputGain;
        p.max_bounce_count = maxBounces; p.rng_state_offset = (uint32_t)frameCount; p.ray_count = rayCount;
        p.debug_ray_count = debugRayCount; p.sample_rate = sampleRate; p.impulse_length = IrLength();
        p.bands = 1; p.time_divisor = 1.0f;
        p.flags = (int)activeSegments.size() >= gridThreshold ? RAR_FLAG_USE_GRID : 0u;
        Check(rar_trace(ctx, &p, GetActiveIRBuffer()), "rar_trace");
        accumFrames++;
    }

    int GetActiveIRBuffer() {  // :212-218
        for (int s = 0; s < 2; s++)
            if (slotLength[s] != IrLength()) {
                Check(rar_ir_clear(ctx, s, IrLength(), 1), "rar_ir_clear");
                slotLength[s] = IrLength();
            }
        return activeIRIndex;
    }

    void UpdateGeometry() {  // :246-250
        activeSegments = SceneToData2D::GetSegmentsFromColliders(obstacleObjects);
        Check(rar_set_walls(ctx, activeSegments.data(), (int32_t)activeSegments.size()), "rar_set_walls");
        wallsUploaded = true;
    }

    std::vector<float> ReadActiveIR() {
        std::vector<float> ir(IrLength());
        Check(rar_ir_read(ctx, GetActiveIRBuffer(), ir.data(), (int64_t)ir.size()), "rar_ir_read");
        return ir;
    }
    std::vector<int64_t> ReadActiveIRFixed() {
        std::vector<int64_t> q(IrLength());
        Check(rar_ir_read_fixed(ctx, GetActiveIRBuffer(), q.data(), (int64_t)q.size()), "rar_ir_read_fixed");
        return q;
    }

    // RayTraceManagerComplex.BakeAudio (:170-227) + PlayResult (:228-245): whole clip, peak-normalised.
    std::vector<float> BakeAudio(const AudioClip &clip) {
        AudioClip same = clip;
        const int saved = sampleRate;
        same.frequency = saved;  // BakeAudio does not resample (it asserts sampleRate == clip.frequency, :61)
        std::vector<float> mono = LoadSample(same);
        std::vector<float> out(mono.size() + IrLength());
        Check(rar_convolve(ctx, GetActiveIRBuffer(), mono.data(), (int32_t)mono.size(), std::max(1, accumFrames), out.data(),
                           (int32_t)out.size()), "rar_convolve");
        float maxVol = 0;
        for (float v : out) maxVol = std::max(maxVol, std::fabs(v));
        if (maxVol > 0.0001f) for (float &v : out) v *= 1.0f / maxVol;
        return out;
    }

    bool ChunksPending() const { return !pending.empty(); }

    void OnDestroy() {  // :281
        rar_destroy(ctx);
        ctx = nullptr;
    }

    std::vector<Segment> activeSegments;

private:
    struct Pending { int ticket, sampleOffset, outputLen; };

    int IrLength() const { return (int)((float)sampleRate * reverbDuration); }

    void ProcessChunk(int sampleOffset, int chunkLen, int accumCount, int slot) {  // :91-123, first half
        const int inputLen = std::min(chunkLen, (int)fullInputSamples.size() - sampleOffset);
        if (inputLen <= 0) return;
        int ticket = -1;
        if (rar_convolve_begin(ctx, slot, fullInputSamples.data() + sampleOffset, inputLen, accumCount, &ticket) != RAR_OK) return;
        pending.push_back({ticket, sampleOffset, inputLen + slotLength[slot]});
    }

    void PollChunks() {  // :115-122: while (!req.done) yield return null; ... PushSamples
        for (size_t i = 0; i < pending.size();) {
            const int st = rar_poll(ctx, pending[i].ticket);
            if (st == 0) { i++; continue; }
            if (st > 0) {
                std::vector<float> result(pending[i].outputLen);
                if (rar_convolve_end(ctx, pending[i].ticket, result.data(), (int32_t)result.size()) == RAR_OK)
                    audioManager->PushSamples(result, pending[i].sampleOffset);
            }
            pending.erase(pending.begin() + i);  // req.hasError -> dropped (:119)
        }
    }

    void Check(int rc, const char *what) {
        if (rc < 0) throw std::runtime_error(std::string(what) + ": " + rar_last_error(ctx));
    }

    rar_context *ctx = nullptr;
    std::vector<float> fullInputSamples;
    std::vector<Pending> pending;
    int activeIRIndex = 0, samplesSinceLastChunk = 0, chunkSamples = 0, nextStreamingOffset = 0;
    int slotLength[2] = {-1, -1};
    bool wallsUploaded = false;
};

}  // namespace rar2d_host

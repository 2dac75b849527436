// host_cpp_driver.cpp -- drives the C++ host mirror (host_cpp/rar2d_host.hpp) the way Unity drives the
// reference: SmollRoom scene, Start, three Update frames, then a streamed clip and a baked clip.  Prints
// numbers that tests/test_gpu_host_cpp.py compares with the oracle.
#include <cstdio>
#include <cstdlib>

#include "../host_cpp/rar2d_host.hpp"

using namespace rar2d_host;

static GameObject box(float px, float py, float qz, float qw, float sx, float sy, AudioMaterial m) {
    GameObject g;
    g.kind = GameObject::Box;
    g.transform.position = {px, py};
    g.transform.qz = qz; g.transform.qw = qw;
    g.transform.lossyScale = {sx, sy};
    g.material = m;
    return g;
}

int main(int argc, char **argv) {
    const char *out_path = argc > 1 ? argv[1] : nullptr;
    try {
        const AudioMaterial border{0.507f, 0.5f, 0.271f, 0.01f}, mat{0.148f, 1.0f, 1.0f, 0.6f};
        RayTraceManager m(0);
        m.rayCount = 15000; m.maxBounces = 5; m.reverbDuration = 1.5f; m.gridThreshold = 1 << 30;
        Transform src, lis;
        src.position = {-18.0f, 9.0f};
        lis.position = {0.0f, -3.68f};
        m.source = &src; m.listener = &lis;
        m.obstacleObjects = {box(0, 10, 0, 1, 100, 1, border), box(0.01f, -5, 0, 1, 100, 1, border),
                             box(-20, 0, 0.7071068f, 0.7071068f, 20, 1, border), box(20, 0, 0.7071068f, 0.7071068f, 20, 1, border),
                             box(-11.8f, 7.18f, 0.47792548f, 0.8784004f, 100, 1, mat)};
        m.Start();
        m.ResetIR();
        for (int f = 0; f < 3; f++) m.Update();
        std::vector<int64_t> q = m.ReadActiveIRFixed();
        long long sum = 0, nz = 0;
        for (int64_t v : q) { sum += v; nz += v != 0; }
        std::printf("segments %zu frames %d accum %d nonzero %lld sum %lld\n", m.activeSegments.size(), m.frameCount, m.accumFrames, nz, sum);
        if (out_path) {
            FILE *f = std::fopen(out_path, "wb");
            std::fwrite(m.activeSegments.data(), sizeof(Segment), m.activeSegments.size(), f);
            std::fwrite(q.data(), sizeof(int64_t), q.size(), f);
            // bake a deterministic clip with the 3-frame IR
            AudioClip clip;
            clip.channels = 2;
            clip.data.resize(2 * 6000);
            unsigned s = 12345u;
            for (size_t i = 0; i < 6000; i++) {
                s = s * 1664525u + 1013904223u;
                const float v = ((s >> 8) & 0xffff) / 65536.0f - 0.5f;
                clip.data[2 * i] = v; clip.data[2 * i + 1] = v;
            }
            std::vector<float> baked = m.BakeAudio(clip);
            std::fwrite(baked.data(), sizeof(float), baked.size(), f);
            std::fclose(f);
            std::printf("baked %zu\n", baked.size());
        }
        // streaming: 0.1 s chunks into the AudioManager ring
        AudioManager am(48000);
        AudioClip clip;
        clip.data.assign(9600, 0.25f);
        m.audioManager = &am; m.inputClip = &clip; m.loop = false;
        m.StartStreaming();
        for (int step = 0; step < 12; step++) { m.Update(); m.FixedUpdate(); }
        for (int i = 0; i < 200 && m.ChunksPending(); i++) m.Update();
        double energy = 0;
        for (float v : am.ringBuffer) energy += (double)v * v;
        std::printf("streaming pending %d ring_energy %.6e\n", (int)m.ChunksPending(), energy);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "host_cpp_driver: %s\n", e.what());
        return 1;
    }
}

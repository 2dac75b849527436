// AudioManager_rar2d.cs (install as Assets/Script/AudioManager.cs) -- the playback sink over librar2d's native ring.
//
// Same component surface as the reference's AudioManager (chunkDuration, IsStreaming, StartStreaming, StopStreaming,
// PushSamples, OnAudioFilterRead), different body: the float[] ring guarded by lock(bufferLock) becomes rar_ring_*,
// a lock-free single-producer / single-consumer ring in pinned host memory (include/rar2d.h).  The audio thread's
// OnAudioFilterRead is wait-free (one exchange-with-zero per sample); the main thread's PushSamples of a
// chunk + reverb tail (76 800 samples in the bundled scenes) can no longer hold it up.
//
// NOT COMPILED IN THIS REPOSITORY'S CI (no dotnet/mono/Unity in the build image); the Python mirror
// realisticaudioraytracing2d_b200/host/audio_manager.py (NativeAudioManager) drives the same entry points under
// test, including a two-thread stress test against the literal mirror of the reference class (tests/test_ring.py).
using System;
using UnityEngine;
using Rar2D;

public class AudioManager : MonoBehaviour
{
    [Range(0.05f, 1.0f)] public float chunkDuration = 0.1f;

    IntPtr ring = IntPtr.Zero;
    int sampleRate;
    volatile bool isStreaming;

    public bool IsStreaming => isStreaming;

    void Awake()
    {
        sampleRate = AudioSettings.outputSampleRate;
        var src = gameObject.AddComponent<AudioSource>();
        src.playOnAwake = false;
        src.loop = true;
        var clip = AudioClip.Create("Silent", sampleRate, 1, sampleRate, false);
        clip.SetData(new float[sampleRate], 0);
        src.clip = clip;
    }

    public void StartStreaming(float reverbDuration)
    {
        if (isStreaming) StopStreaming();
        if (ring != IntPtr.Zero) { RarNative.rar_ring_destroy(ring); ring = IntPtr.Zero; }
        // bufferSize = CeilToInt(sampleRate * (reverbDuration + 1)) is computed by the library
        if (RarNative.rar_ring_create(sampleRate, reverbDuration, out ring) < 0) { ring = IntPtr.Zero; return; }
        isStreaming = true;
        GetComponent<AudioSource>().Play();
    }

    public void StopStreaming()
    {
        if (!isStreaming) return;
        isStreaming = false;
        if (ring != IntPtr.Zero) RarNative.rar_ring_stop(ring);
        GetComponent<AudioSource>()?.Stop();
    }

    public void PushSamples(float[] samples, int sampleOffset)
    {
        if (!isStreaming || ring == IntPtr.Zero) return;
        RarNative.rar_ring_push(ring, samples, samples.Length, sampleOffset);
    }

    void OnAudioFilterRead(float[] data, int channels)
    {
        if (!isStreaming || ring == IntPtr.Zero) return;
        RarNative.rar_ring_drain(ring, data, data.Length, channels);
    }

    void OnDestroy()
    {
        StopStreaming();
        if (ring != IntPtr.Zero) { RarNative.rar_ring_destroy(ring); ring = IntPtr.Zero; }
    }
}

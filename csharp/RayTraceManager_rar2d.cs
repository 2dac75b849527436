// RayTraceManager_rar2d.cs (install as Assets/Script/RayTraceManager.cs) -- the live orchestrator component with its GPU work re-pointed at librar2d.
//
// Same serialized fields and Unity messages as the reference component (so existing scenes keep their
// inspector values), different body: the ComputeShader/ComputeBuffer/AsyncGPUReadback plumbing is
// replaced by calls into the C-ABI (RarNative).  What changes in behaviour is listed in DESIGN.md:
// Trace and ProcessHits are one native call, so accumFrames advances in RunSimulation; IR buffers are
// native slots 0/1; there is no debug texture.
//
// NOT COMPILED IN THIS REPOSITORY'S CI (no dotnet/mono/Unity in the build image); the Python mirror
// realisticaudioraytracing2d_b200/host/ray_trace_manager.py runs the same sequence under test.
using System;
using System.Collections;
using System.Collections.Generic;
using UnityEngine;
using Helpers;
using Rar2D;

public class RayTraceManager : MonoBehaviour
{
    [Header("Native")] public int cudaDevice = 0;
    public int gridThreshold = 64;   // wall count from which traces use the uniform grid (identical results)

    [Header("Simulation")]
    [Range(10, 100000)] public int rayCount = 1000;
    [Range(1, 10)] public int maxBounces = 5;
    public float speedOfSound = 343f;
    public bool dynamicObstacles = false;

    [Header("Audio")]
    public AudioClip inputClip;
    public AudioManager audioManager;
    public int sampleRate = 48000;
    [Range(0.1f, 10f)] public float inputGain = 1.0f;
    [Range(0.1f, 5.0f)] public float reverbDuration = 2f;
    public bool loop = true;

    [Header("Scene")]
    public Transform source, listener;
    [Range(0.1f, 5f)] public float listenerRadius = 0.5f;
    public List<GameObject> obstacleObjects;

    [Header("Debug")]
    [Range(5, 100)] public int debugRayCount = 100;

    IntPtr native = IntPtr.Zero;
    Segment[] walls = new Segment[0];
    float[] clipSamples;
    int activeSlot, accumFrames, pendingSamples, chunkSamples, streamOffset;
    readonly int[] slotLength = { -1, -1 };

    int IrLength => (int)(sampleRate * reverbDuration);

    void Awake()
    {
        if (!RarNative.Ok(IntPtr.Zero, RarNative.rar_create(cudaDevice, out native), "rar_create")) native = IntPtr.Zero;
    }

    void Start() => UpdateGeometry();

    void Update()
    {
        if (native == IntPtr.Zero || !source || !listener) return;
        RunSimulation();
        if (Input.GetKeyDown(KeyCode.Space) && audioManager)
        {
            if (audioManager.IsStreaming) audioManager.StopStreaming(); else StartStreaming();
        }
        if (Input.GetKeyDown(KeyCode.R)) { ResetIR(); audioManager?.StopStreaming(); }
    }

    void FixedUpdate()
    {
        if (native == IntPtr.Zero || !audioManager || !audioManager.IsStreaming) return;
        if (dynamicObstacles) UpdateGeometry();
        pendingSamples += Mathf.RoundToInt(Time.fixedDeltaTime * sampleRate);
        if (pendingSamples < chunkSamples) return;
        if (streamOffset >= clipSamples.Length)
        {
            if (loop) streamOffset = 0; else audioManager.StopStreaming();
        }
        if (!audioManager.IsStreaming) return;
        StartCoroutine(ProcessChunk(streamOffset, chunkSamples, Mathf.Max(1, accumFrames), ActiveSlot()));
        activeSlot ^= 1;                 // ping-pong: later traces go to the other slot
        streamOffset += chunkSamples;
        ResetIR();
        pendingSamples -= chunkSamples;
    }

    IEnumerator ProcessChunk(int sampleOffset, int chunkLen, int accumCount, int slot)
    {
        int inputLen = Mathf.Min(chunkLen, clipSamples.Length - sampleOffset);
        if (inputLen <= 0) yield break;
        var chunk = new float[inputLen];
        Array.Copy(clipSamples, sampleOffset, chunk, 0, inputLen);
        if (!RarNative.Ok(native, RarNative.rar_convolve_begin(native, slot, chunk, inputLen, accumCount, out int ticket), "convolve_begin"))
            yield break;
        int state;
        while ((state = RarNative.rar_poll(native, ticket)) == 0) yield return null;
        if (state < 0) yield break;
        var result = new float[inputLen + slotLength[slot]];
        if (RarNative.Ok(native, RarNative.rar_convolve_end(native, ticket, result, result.Length), "convolve_end"))
            audioManager.PushSamples(result, sampleOffset);
    }

    void StartStreaming()
    {
        streamOffset = 0;
        pendingSamples = 0;
        chunkSamples = Mathf.RoundToInt(sampleRate * audioManager.chunkDuration);
        clipSamples = LoadSample(inputClip);
        ResetIR();
        audioManager.StartStreaming(reverbDuration);
    }

    float[] LoadSample(AudioClip clip)
    {
        int n = clip.samples, ch = clip.channels;
        var interleaved = new float[n * ch];
        clip.GetData(interleaved, 0);
        var mono = new float[n];
        for (int i = 0; i < n; i++)
        {
            float acc = 0;
            for (int c = 0; c < ch; c++) acc += interleaved[i * ch + c];
            mono[i] = acc / ch;
        }
        if (clip.frequency == sampleRate) return mono;
        float ratio = (float)clip.frequency / sampleRate;
        var resampled = new float[Mathf.RoundToInt(n / ratio)];
        for (int i = 0; i < resampled.Length; i++)
        {
            float pos = i * ratio;
            int i0 = Mathf.FloorToInt(pos), i1 = Mathf.Min(i0 + 1, n - 1);
            resampled[i] = Mathf.Lerp(mono[i0], mono[i1], pos - i0);
        }
        return resampled;
    }

    void ResetIR()
    {
        accumFrames = 0;
        int slot = ActiveSlot();
        if (RarNative.Ok(native, RarNative.rar_ir_clear(native, slot, IrLength, 1), "ir_clear")) slotLength[slot] = IrLength;
    }

    void RunSimulation()
    {
        if (walls.Length == 0 && obstacleObjects != null && obstacleObjects.Count > 0) UpdateGeometry();
        var p = new RarTraceParams
        {
            sourceX = source.position.x, sourceY = source.position.y,
            listenerX = listener.position.x, listenerY = listener.position.y,
            listenerRadius = listenerRadius, speedOfSound = speedOfSound, inputGain = inputGain,
            maxBounceCount = maxBounces, rngStateOffset = (uint)Time.frameCount, rayCount = rayCount,
            debugRayCount = debugRayCount, sampleRate = sampleRate, impulseLength = IrLength,
            bands = 1, timeDivisor = 1f, flags = walls.Length >= gridThreshold ? RarNative.FlagUseGrid : 0u, rayBegin = 0, rayEnd = 0,
        };
        if (RarNative.Ok(native, RarNative.rar_trace(native, ref p, ActiveSlot()), "trace")) accumFrames++;
    }

    int ActiveSlot()
    {
        for (int s = 0; s < 2; s++)
            if (slotLength[s] != IrLength && RarNative.Ok(native, RarNative.rar_ir_clear(native, s, IrLength, 1), "ir_clear"))
                slotLength[s] = IrLength;
        return activeSlot;
    }

    void UpdateGeometry()
    {
        walls = SceneToData2D.GetSegmentsFromColliders(obstacleObjects).ToArray();
        if (native != IntPtr.Zero) RarNative.Ok(native, RarNative.rar_set_walls(native, walls, walls.Length), "set_walls");
    }

    void OnDestroy()
    {
        RarNative.rar_destroy(native);   // NULL-tolerant, like buffer?.Release()
        native = IntPtr.Zero;
    }
}

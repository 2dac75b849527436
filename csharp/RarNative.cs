// RarNative.cs -- P/Invoke binding of librar2d (include/rar2d.h) for the Unity host.
//
// Drop this file and csharp/RayTraceManager_rar2d.cs into Assets/Script/ (keeping the existing .meta of
// RayTraceManager.cs so the scenes' script guid 2913e124... still resolves), and put librar2d.so in
// Assets/Plugins/x86_64/.  Struct layouts are the reference's own: `Segment` is the 40-byte
// LayoutKind.Sequential struct of Helpers/SceneHelper.cs:15-22 and is passed to rar_set_walls as is.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no dotnet/mono/Unity.  The same entry points
// are exercised through ctypes (realisticaudioraytracing2d_b200/_capi.py) by the test-suite, and
// tests/test_capi_symbols.py checks that every symbol named here is exported.
using System;
using System.Runtime.InteropServices;

namespace Rar2D
{
    [StructLayout(LayoutKind.Sequential)]
    public struct RarTraceParams
    {
        public float sourceX, sourceY;
        public float listenerX, listenerY;
        public float listenerRadius, speedOfSound, inputGain;
        public int maxBounceCount;
        public uint rngStateOffset;
        public int rayCount;
        public int debugRayCount;
        public int sampleRate;
        public int impulseLength;
        public int bands;
        public float timeDivisor;
        public uint flags;
        public long rayBegin, rayEnd;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct RarCounters
    {
        public ulong rayBounces, nearestTests, shadowTests, directHits, neeHits;
    }

    public static class RarNative
    {
        const string Lib = "rar2d";

        public const uint FlagExactRayCount = 1u;
        public const uint FlagCountTests = 2u;
        public const uint FlagCountExecuted = 4u;
        public const uint FlagUseGrid = 8u;

        [DllImport(Lib)] public static extern int rar_version();
        [DllImport(Lib)] public static extern int rar_create(int device, out IntPtr ctx);
        [DllImport(Lib)] public static extern int rar_destroy(IntPtr ctx);
        [DllImport(Lib)] public static extern IntPtr rar_last_error(IntPtr ctx);
        [DllImport(Lib)] public static extern int rar_set_stream(IntPtr ctx, IntPtr cudaStream);
        [DllImport(Lib)] public static extern int rar_sync(IntPtr ctx);

        // Segment[] is blittable (40 bytes, sequential): pinned and passed without marshalling.
        [DllImport(Lib)] public static extern int rar_set_walls(IntPtr ctx, [In] Segment[] segments, int n);
        [DllImport(Lib)] public static extern int rar_set_wall_band_absorption(IntPtr ctx, [In] float[] absorption, int n, int bands);

        [DllImport(Lib)] public static extern int rar_set_air_absorption(IntPtr ctx, [In] float[] alphaPerMetre, int bands);
        [DllImport(Lib)] public static extern int rar_ir_clear(IntPtr ctx, int slot, int impulseLength, int bands);
        [DllImport(Lib)] public static extern int rar_ir_read(IntPtr ctx, int slot, [Out] float[] dst, long n);
        [DllImport(Lib)] public static extern int rar_ir_read_fixed(IntPtr ctx, int slot, [Out] long[] dst, long n);
        [DllImport(Lib)] public static extern int rar_ir_write(IntPtr ctx, int slot, [In] float[] ir, int impulseLength, int bands);
        [DllImport(Lib)] public static extern int rar_ir_device_ptr(IntPtr ctx, int slot, out IntPtr devicePtr, out long nWords);

        [DllImport(Lib)] public static extern int rar_ir_read_begin(IntPtr ctx, int slot, long n, out int ticket);
        [DllImport(Lib)] public static extern int rar_ir_read_end(IntPtr ctx, int ticket, [Out] float[] output, long n);
        [DllImport(Lib)] public static extern int rar_allreduce_slots([In] IntPtr[] contexts, int n, int slot);
        // LoadSample (RayTraceManager.cs:135-167) for a batch of clips on the GPU
        [DllImport(Lib)] public static extern long rar_prepared_length(long samples, int clipFrequency, int sampleRate);
        [DllImport(Lib)] public static extern int rar_prepare_clips(IntPtr ctx, [In] float[] raw, long samples, int channels, int clipFrequency, int sampleRate, int nClips, [Out] float[] output, long outStride);
        [DllImport(Lib)] public static extern int rar_prepare_clips_device(IntPtr ctx, IntPtr dRaw, long samples, int channels, int clipFrequency, int sampleRate, int nClips, IntPtr dOut, long outStride);
        // one process per GPU: all-reduce kernel over CUDA-IPC peer memory (handles are 80 bytes each)
        [DllImport(Lib)] public static extern int rar_exchange_create(IntPtr ctx, long capacityWords, [Out] byte[] handle);
        [DllImport(Lib)] public static extern int rar_exchange_connect(IntPtr ctx, int rank, int world, [In] byte[] handles);
        [DllImport(Lib)] public static extern int rar_exchange_allreduce(IntPtr ctx, int slot, int mode);
        [DllImport(Lib)] public static extern int rar_exchange_status(IntPtr ctx);
        [DllImport(Lib)] public static extern int rar_exchange_destroy(IntPtr ctx);

        [DllImport(Lib)] public static extern int rar_trace(IntPtr ctx, ref RarTraceParams p, int slot);
        [DllImport(Lib)] public static extern int rar_trace_interleaved(IntPtr ctx, ref RarTraceParams p, int slot, int rank, int world, int chunkLog2);
        [DllImport(Lib)] public static extern int rar_trace_frames(IntPtr ctx, ref RarTraceParams p, int slot, int nFrames);
        [DllImport(Lib)] public static extern int rar_trace_listeners(IntPtr ctx, ref RarTraceParams p, [In] float[] listenersXY, int nListeners, int firstSlot);
        [DllImport(Lib)] public static extern int rar_trace_hits(IntPtr ctx, ref RarTraceParams p, IntPtr hits, IntPtr keys, long capacity, out long count);
        [DllImport(Lib)] public static extern int rar_get_counters(IntPtr ctx, out RarCounters c, int reset);
        [DllImport(Lib)] public static extern int rar_get_debug_rays(IntPtr ctx, [Out] UnityEngine.Vector4[] dst, long nFloat4);

        // banded model: filter-bank synthesis of a banded slot (RayTraceManagerComplex's WindowSize layout)
        [DllImport(Lib)] public static extern int rar_set_band_edges(IntPtr ctx, [In] float[] edgesHz, int bands, int sampleRate);
        [DllImport(Lib)] public static extern int rar_synthesize_ir(IntPtr ctx, int slot, [Out] float[] dst, long n);

        [DllImport(Lib)] public static extern int rar_convolve(IntPtr ctx, int slot, [In] float[] input, int inLen, int accumCount, [Out] float[] output, int outLen);
        [DllImport(Lib)] public static extern int rar_convolve_begin(IntPtr ctx, int slot, [In] float[] input, int inLen, int accumCount, out int ticket);
        [DllImport(Lib)] public static extern int rar_poll(IntPtr ctx, int ticket);
        [DllImport(Lib)] public static extern int rar_convolve_end(IntPtr ctx, int ticket, [Out] float[] output, int outLen);

        [DllImport(Lib)] public static extern int rar_conv_create(IntPtr ctx, int nStreams, int block, int maxIrLen, out IntPtr conv);
        [DllImport(Lib)] public static extern int rar_conv_destroy(IntPtr conv);
        [DllImport(Lib)] public static extern int rar_conv_set_ir(IntPtr conv, int stream, [In] float[] ir, int irLen, float scale);
        [DllImport(Lib)] public static extern int rar_conv_set_ir_from_slot(IntPtr conv, int stream, int slot, int accumCount);
        [DllImport(Lib)] public static extern int rar_conv_set_irs(IntPtr conv, int firstStream, int n, [In] float[] irs, int irLen, long irStride, float scale);
        [DllImport(Lib)] public static extern int rar_conv_set_irs_from_slots(IntPtr conv, int firstStream, int n, [In] int[] slots, [In] int[] accumCounts);
        [DllImport(Lib)] public static extern int rar_conv_update_ir(IntPtr conv, int stream, [In] float[] ir, int irLen, float scale);
        [DllImport(Lib)] public static extern int rar_conv_update_ir_from_slot(IntPtr conv, int stream, int slot, int accumCount);
        [DllImport(Lib)] public static extern int rar_conv_reset(IntPtr conv);
        [DllImport(Lib)] public static extern int rar_conv_process(IntPtr conv, [In] float[] input, [Out] float[] output);
        [DllImport(Lib)] public static extern int rar_conv_process_device(IntPtr conv, IntPtr dIn, IntPtr dOut);
        [DllImport(Lib)] public static extern long rar_conv_bytes_per_block(IntPtr conv);

        // AudioManager's ring as a native lock-free SPSC structure (PushSamples on the main thread, OnAudioFilterRead on the audio thread)
        [DllImport(Lib)] public static extern int rar_ring_create(int outputSampleRate, float reverbDuration, out IntPtr ring);
        [DllImport(Lib)] public static extern int rar_ring_destroy(IntPtr ring);
        [DllImport(Lib)] public static extern int rar_ring_reset(IntPtr ring);
        [DllImport(Lib)] public static extern int rar_ring_stop(IntPtr ring);
        [DllImport(Lib)] public static extern int rar_ring_size(IntPtr ring);
        [DllImport(Lib)] public static extern int rar_ring_is_pinned(IntPtr ring);
        [DllImport(Lib)] public static extern long rar_ring_frames_drained(IntPtr ring);
        [DllImport(Lib)] public static extern int rar_ring_push(IntPtr ring, [In] float[] samples, int n, long sampleOffset);
        [DllImport(Lib)] public static extern int rar_ring_drain(IntPtr ring, [In, Out] float[] data, int dataLength, int channels);
        [DllImport(Lib)] public static extern int rar_conv_process_to_ring(IntPtr conv, [In] float[] input, [In] IntPtr[] rings, long sampleOffset);

        [DllImport(Lib)] public static extern int rar_device_info(IntPtr ctx, out int smCount, out int smClockKhz, out int smemOptinBytes);
        [DllImport(Lib)] public static extern int rar_measure_fp32_peak(IntPtr ctx, out double laneOpsPerSecond);
        [DllImport(Lib)] public static extern int rar_selftest_arithmetic(IntPtr ctx, long nSamples, uint seed, [Out] ulong[] mismatches5);
        [DllImport(Lib)] public static extern int rar_debug_grid(IntPtr ctx, out int nx, out int ny, out long nItems, out ulong digest);
        [DllImport(Lib)] public static extern long rar_launch_count(IntPtr ctx);

        public static string LastError(IntPtr ctx) => Marshal.PtrToStringAnsi(rar_last_error(ctx));

        // The reference's convention is "null-guard and silently skip" (RayTraceManager.cs:52,119,222);
        // a failed native call is logged and skipped the same way.
        public static bool Ok(IntPtr ctx, int status, string what)
        {
            if (status >= 0) return true;
            UnityEngine.Debug.LogWarning($"rar2d: {what} failed ({status}): {LastError(ctx)}");
            return false;
        }
    }
}

"""Host-side mirrors of the reference's C# components that need no GPU: AudioManager ring buffer
(AudioManager.cs:45-69), ComputeHelper thread arithmetic (Helpers/ComputeHelper.cs:25-32), SceneToData2D
(Helpers/SceneHelper.cs:29-110), LoadSample (RayTraceManager.cs:135-167), sharding helpers."""
import numpy as np
import pytest

from realisticaudioraytracing2d_b200.host.audio_manager import AudioManager
from realisticaudioraytracing2d_b200.host.compute_helper import ComputeHelper
from realisticaudioraytracing2d_b200.host.scene_helper import (AcousticSurface, AudioMaterial, BoxCollider2D,
                                                               CircleCollider2D, GameObject, PolygonCollider2D,
                                                               SceneToData2D, Transform)
from realisticaudioraytracing2d_b200.host.sharding import dispatched_threads, shard_range


def test_audio_manager_ring_buffer():
    am = AudioManager(outputSampleRate=1000)
    am.PushSamples(np.ones(10, np.float32), 0)                      # not streaming: ignored (:47)
    am.StartStreaming(0.5)
    assert am.IsStreaming and am.bufferSize == 1500                 # ceil(sampleRate * (reverb + 1)) (:30)
    am.PushSamples(np.arange(1, 11, dtype=np.float32), 5)
    am.PushSamples(np.full(10, 100, np.float32), 10)                # overlap-add (:52)
    out = np.zeros(40, np.float32)
    am.OnAudioFilterRead(out, 2)                                    # 20 frames, duplicated over 2 channels (:61-67)
    mono = out[::2]
    assert np.array_equal(out[::2], out[1::2])
    expect = np.zeros(20, np.float32)
    expect[5:15] += np.arange(1, 11)
    expect[10:20] += 100
    assert np.array_equal(mono, expect)
    assert am.readHead == 20 and not am.ringBuffer[:20].any()       # drained and zeroed
    am.PushSamples(np.ones(20, np.float32), 1490)                   # wraps around the ring
    assert am.ringBuffer[1490:].sum() == 10 and am.ringBuffer[:10].sum() == 10
    am.StopStreaming()
    assert not am.IsStreaming


def test_compute_helper_thread_groups():
    assert ComputeHelper.GetThreadGroupCount(15000) == 235 and ComputeHelper.DispatchedThreads(15000) == 15040
    assert ComputeHelper.DispatchedThreads(1000) == 1024 and ComputeHelper.DispatchedThreads(64) == 64
    assert dispatched_threads(15000) == 15040 and dispatched_threads(15000, exact=True) == 15000
    ComputeHelper.Release(None, None)                               # null-tolerant (:223-226)


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 15040, 1 << 20):
        for world in (1, 2, 3, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 3, 2)


def test_scene_to_data_box_circle_polygon():
    mat = AudioMaterial(0.2, 0.3, 0.4, 1.5)
    box = GameObject(Transform((1.0, 2.0), (0.0, 1.0), (4.0, 2.0)), BoxCollider2D(size=(1.0, 1.0), offset=(0.0, 0.0)), AcousticSurface(mat))
    segs = SceneToData2D.GetSegmentsFromColliders([box])
    assert len(segs) == 4
    assert np.allclose(segs["start"], [(-1, 1), (3, 1), (3, 3), (-1, 3)]) and np.allclose(segs["end"][3], (-1, 1))
    assert np.allclose(segs["normal"], [(0, -1), (1, 0), (0, 1), (-1, 0)])   # outward for a CCW loop (:92-93)
    assert np.all(segs["absorption"] == np.float32(0.2)) and np.all(segs["ior"] == np.float32(1.5))
    # negative scale flips the winding sign so the normals still point outward (:81)
    flipped = GameObject(Transform((0.0, 0.0), (0.0, 1.0), (-2.0, 2.0)), BoxCollider2D(), AcousticSurface(mat))
    f = SceneToData2D.GetSegmentsFromColliders([flipped])
    mid = (f["start"] + f["end"]) / 2
    assert np.all(np.einsum("ij,ij->i", mid, f["normal"]) > 0)
    circ = GameObject(Transform((0.0, 0.0), (0.0, 1.0), (1.0, 1.0)), CircleCollider2D(radius=2.0), AcousticSurface(mat))
    c = SceneToData2D.GetSegmentsFromColliders([circ])
    assert len(c) == 32 and np.allclose(np.hypot(*c["start"].T), 2.0, atol=1e-6)       # CIRCLE_RESOLUTION (:26)
    poly = GameObject(Transform(), PolygonCollider2D(paths=[[(0, 0), (1, 0), (0, 1)], [(2, 2), (3, 2), (3, 3), (2, 3)]]), AcousticSurface(mat))
    assert len(SceneToData2D.GetSegmentsFromColliders([poly])) == 7
    off = GameObject(Transform(), BoxCollider2D(enabled=False), AcousticSurface(mat))
    assert len(SceneToData2D.GetSegmentsFromColliders([off, GameObject()])) == 0        # disabled / no collider (:34)
    with pytest.raises(AttributeError):
        SceneToData2D.GetSegmentsFromColliders([GameObject(Transform(), BoxCollider2D(), None)])  # null surface (:102-103)


def test_load_sample_mono_mix_and_resample():
    from realisticaudioraytracing2d_b200.host.ray_trace_manager import AudioClip, RayTraceManager
    m = RayTraceManager.__new__(RayTraceManager)                    # LoadSample touches no GPU state
    m.sampleRate = 48000
    stereo = np.array([0.25, 0.75, -1.0, 1.0, 0.5, 0.5], np.float32)
    assert np.array_equal(m.LoadSample(AudioClip(stereo, 2, 48000)), np.array([0.5, 0.0, 0.5], np.float32))
    ramp = np.arange(441, dtype=np.float32)
    out = m.LoadSample(AudioClip(ramp, 1, 44100))
    assert len(out) == 480                                          # RoundToInt(samples / ratio) (:153)
    ratio = np.float32(44100) / np.float32(48000)
    assert np.allclose(out, np.minimum(np.arange(480, dtype=np.float32) * ratio, 440), atol=1e-3)


def test_interleaved_chunks_partition_the_dispatch():
    from realisticaudioraytracing2d_b200.host.sharding import interleaved_chunks
    for total, world, k in [(15040, 8, 12), (15040, 3, 5), (1 << 20, 8, 14), (100, 4, 5), (0, 2, 6), (4849664, 8, 14)]:
        seen = np.zeros(total, np.int32)
        sizes = []
        for r in range(world):
            ranges = interleaved_chunks(total, r, world, k)
            sizes.append(sum(e - b for b, e in ranges))
            for b, e in ranges:
                assert b % (1 << k) == 0 and 0 < e - b <= (1 << k)
                seen[b:e] += 1
        assert np.all(seen == 1)                                   # every thread id exactly once
        assert max(sizes) - min(sizes) <= (1 << k)                 # shares differ by at most one chunk

"""The native playback ring (csrc/ring.cu) against the Python mirror of Assets/Script/AudioManager.cs.

The ring needs no GPU, so these run in the CPU suite.  The stress test drives it from two real threads (ctypes calls
release the GIL): a producer doing the reference's chunk cadence -- overlapping pushes of chunk + IR samples every
`chunk` samples (RayTraceManager.cs:82,121-122) -- and a consumer draining audio-callback-sized blocks.  The consumer
only takes frames that are final (no later push reaches them), and the producer never laps the read head, so the result
is independent of the interleaving and must equal the mirror's sequential run sample for sample."""
import threading
import time

import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import _capi
from realisticaudioraytracing2d_b200.host.audio_manager import AudioManager, NativeAudioManager


def test_ring_matches_the_mirror_call_by_call():
    rng = np.random.default_rng(0)
    ring, am = _capi.Ring(48000, 0.05), AudioManager(48000)
    am.StartStreaming(0.05)
    assert ring.size == am.bufferSize == 50400
    off = 0
    for k in range(60):
        n = int(rng.integers(0, 30000))
        x = rng.standard_normal(n).astype(np.float32)
        ring.push(x, off)
        am.PushSamples(x, off)
        off += int(rng.integers(0, 9000))
        ch = int(rng.integers(1, 3))
        a = np.full(int(rng.integers(0, 5000)), 7.0, np.float32)
        b = a.copy()
        ring.drain(a, ch)
        am.OnAudioFilterRead(b, ch)
        assert np.array_equal(a, b), k                      # incl. the untouched tail when len % channels != 0
    assert ring.frames_drained > 0


def test_a_chunk_longer_than_the_ring_laps_and_accumulates():
    ring, am = _capi.Ring(1000, 0.0), AudioManager(1000)
    am.StartStreaming(0.0)
    x = np.arange(2500, dtype=np.float32)
    ring.push(x, 999_999_999_999)                           # 64-bit offsets; the mirror takes Python ints
    am.PushSamples(x, 999_999_999_999)
    a, b = np.zeros(1000, np.float32), np.zeros(1000, np.float32)
    ring.drain(a, 1)
    am.OnAudioFilterRead(b, 1)
    assert np.array_equal(a, b) and a.any()


def test_stop_and_reset():
    ring = _capi.Ring(1000, 1.0)
    ring.push(np.ones(10, np.float32), 0)
    ring.stop()                                             # StopStreaming: pushes and drains are no-ops
    ring.push(np.ones(10, np.float32), 0)
    d = np.full(4, 9.0, np.float32)
    ring.drain(d, 1)
    assert np.all(d == 9.0)
    ring.reset()                                            # StartStreaming again: silent, read head 0
    d = np.full(20, 9.0, np.float32)
    ring.drain(d, 1)
    assert not d.any() and ring.frames_drained == 20
    with pytest.raises(_capi.RarError):
        ring.push(np.ones(4, np.float32), -1)


def test_two_thread_stress_equals_the_sequential_mirror():
    sr, reverb, chunk, tail, n_chunks, block = 48000, 1.5, 4800, 72000, 150, 1024
    rng = np.random.default_rng(5)
    chunks = [rng.standard_normal(chunk + tail).astype(np.float32) for _ in range(8)]
    ring = _capi.Ring(sr, reverb)
    size = ring.size
    assert size == 120000 and chunk + tail < size
    final = [0]                                             # frames below this index receive no further push
    total = n_chunks * chunk
    got = np.zeros(total, np.float32)
    errors = []

    def producer():
        try:
            for k in range(n_chunks):
                off = k * chunk
                while off + chunk + tail - ring.frames_drained > size:      # never lap the read head
                    time.sleep(0)
                ring.push(chunks[k % 8], off)               # overlaps the 15 pushes before it
                final[0] = (k + 1) * chunk
        except Exception as e:                              # noqa: BLE001
            errors.append(e)

    def consumer():
        try:
            pos = 0
            buf = np.zeros(2 * block, np.float32)
            while pos < total:
                n = min(block, final[0] - pos)
                if n <= 0:
                    time.sleep(0)
                    continue
                view = buf[: 2 * n]
                ring.drain(view, 2)                         # stereo output device: the sample goes to both channels
                assert np.array_equal(view[0::2], view[1::2])
                got[pos:pos + n] = view[0::2]
                pos += n
        except Exception as e:                              # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=producer), threading.Thread(target=consumer)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors and not any(t.is_alive() for t in threads)

    am = AudioManager(sr)                                   # the mirror, sequentially: push k, then drain what became final
    am.StartStreaming(reverb)
    want = np.zeros(total, np.float32)
    for k in range(n_chunks):
        am.PushSamples(chunks[k % 8], k * chunk)
        d = np.zeros(chunk, np.float32)
        am.OnAudioFilterRead(d, 1)
        want[k * chunk:(k + 1) * chunk] = d
    assert np.array_equal(got, want) and np.count_nonzero(got) > total // 2


def test_native_audio_manager_has_the_mirror_interface():
    rng = np.random.default_rng(2)
    a, b = AudioManager(44100), NativeAudioManager(44100)
    for m in (a, b):
        m.PushSamples(np.ones(8, np.float32), 0)            # not streaming yet: ignored (:47)
        m.StartStreaming(0.25)
    assert a.bufferSize == b.bufferSize == 55125 and b.IsStreaming
    for k in range(20):
        x = rng.standard_normal(3000).astype(np.float32)
        for m in (a, b):
            m.PushSamples(x, k * 1111)
        da, db = np.zeros(2048, np.float32), np.zeros(2048, np.float32)
        a.OnAudioFilterRead(da, 2)
        b.OnAudioFilterRead(db, 2)
        assert np.array_equal(da, db)
    for m in (a, b):
        m.OnDestroy()
        assert not m.IsStreaming

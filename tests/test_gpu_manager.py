"""The host mirror of RayTraceManager / RayTraceManagerComplex driving the CUDA path the way Unity drives
the reference: Start, Update per frame, FixedUpdate per 0.02 s, 0.1 s chunks convolved with ping-pong IRs
and overlap-added into the AudioManager ring (RayTraceManager.cs:45-133)."""
import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import _capi, scenes
from realisticaudioraytracing2d_b200.host.audio_manager import AudioManager
from realisticaudioraytracing2d_b200.host.ray_trace_manager import AudioClip, RayTraceManager, RayTraceManagerComplex
from realisticaudioraytracing2d_b200.host.scene_helper import AcousticSurface, BoxCollider2D, GameObject, Transform
from tests.common import oracle_params, rel_l2

pytestmark = pytest.mark.gpu


def _smoll_objects():
    R90, R57 = (0.7071068, 0.7071068), (0.47792548, 0.8784004)
    spec = [((0.0, 10.0), (0.0, 1.0), (100.0, 1.0), scenes.BORDER), ((0.01, -5.0), (0.0, 1.0), (100.0, 1.0), scenes.BORDER),
            ((-20.0, 0.0), R90, (20.0, 1.0), scenes.BORDER), ((20.0, 0.0), R90, (20.0, 1.0), scenes.BORDER),
            ((-11.8, 7.18), R57, (100.0, 1.0), scenes.MATERIAL)]
    return [GameObject(Transform(p, q, s), BoxCollider2D(), AcousticSurface(m)) for p, q, s, m in spec]


def _manager(ctx, cls=RayTraceManager):
    m = cls(context=ctx)
    m.rayCount, m.maxBounces, m.reverbDuration = 15000, 5, 1.5       # SmollRoom.unity:155-164
    m.source, m.listener = Transform((-18.0, 9.0)), Transform((0.0, -3.68))
    m.obstacleObjects = _smoll_objects()
    return m


def test_update_accumulates_frames_like_the_oracle(ctx, oracle):
    m = _manager(ctx)
    m.Start()
    m.ResetIR()
    for _ in range(3):
        m.Update()
    assert m.accumFrames == 3 and m.frameCount == 3
    sc = scenes.smoll_room()
    assert m.activeSegments.tobytes() == sc.walls.tobytes()
    hist = np.zeros(72000, np.int64)
    for f in (1, 2, 3):
        P = oracle.make_params(source_x=-18.0, source_y=9.0, listener_x=0.0, listener_y=-3.68, ray_count=15000,
                               max_bounce_count=5, rng_state_offset=f, impulse_length=72000, debug_ray_count=100)
        oracle.trace(sc.walls.view(oracle.SEGMENT_DTYPE), P, hist=hist)
    assert np.array_equal(m.ReadActiveIR().view(np.uint32), oracle.ir_to_float(hist).view(np.uint32))
    paths = m.GetDebugRayPaths().reshape(100, 6, 4)
    assert np.allclose(paths[:, 0, :2], (-18.0, 9.0))


def test_streaming_chunks_match_per_chunk_reference_convolution(ctx, oracle):
    """Each 0.1 s chunk is convolved with the IR accumulated since the previous chunk, divided by the number of
    frames, and overlap-added at its sample offset: replay the same schedule with the oracle."""
    m = _manager(ctx)
    m.rayCount = 2000
    m.audioManager = AudioManager(outputSampleRate=48000)
    clip = scenes.synthetic_clip(12000)
    m.inputClip = AudioClip(clip, 1, 48000)
    m.loop = False
    m.Start()
    m.StartStreaming()
    assert m.chunkSamples == 4800
    sc = scenes.smoll_room()
    expect = np.zeros(48000 * 3, np.float32)
    hist = np.zeros(72000, np.int64)
    frames_in_ir = 0
    for step in range(18):                                           # 18 fixed steps = 3 chunks + drain
        m.Update()                                                   # one rendered frame per fixed step
        P = oracle.make_params(source_x=-18.0, source_y=9.0, listener_x=0.0, listener_y=-3.68, ray_count=2000,
                               max_bounce_count=5, rng_state_offset=m.frameCount, impulse_length=72000, debug_ray_count=100)
        oracle.trace(sc.walls.view(oracle.SEGMENT_DTYPE), P, hist=hist)
        frames_in_ir += 1
        before = m.nextStreamingOffset
        m.FixedUpdate()
        if m.nextStreamingOffset != before and before < len(clip):   # a chunk was launched with the IR so far
            y = oracle.convolve(clip[before: before + 4800], oracle.ir_to_float(hist), max(1, frames_in_ir))
            expect[before: before + len(y)] += y
            hist[:] = 0
            frames_in_ir = 0
    for _ in range(50):                                              # let the coroutines finish
        m.Update()
        if not m._coroutines:
            break
    assert not m._coroutines
    got = m.audioManager.ringBuffer.copy()
    assert np.abs(expect).max() > 0
    assert rel_l2(got[: len(expect)][: len(got)], expect[: len(got)]) <= 1e-4


def test_bake_audio_whole_clip(ctx, oracle):
    m = _manager(ctx, RayTraceManagerComplex)
    m.Start()
    m.ResetIR()
    m.Update()
    m.Update()
    clip = scenes.synthetic_clip()
    stereo = np.repeat(clip, 2)
    m.inputClip = AudioClip(stereo, 2, 48000)
    out = m.BakeAudio()
    ir = m.ReadActiveIR()
    want = oracle.convolve(clip, ir, 2)
    want = want / np.abs(want).max()
    assert out.shape == (len(clip) + 72000,) and abs(np.abs(out).max() - 1.0) < 1e-6     # PlayResult peak-normalises
    assert rel_l2(out, want) <= 1e-4


def test_load_samples_batched_equals_load_sample(ctx):
    """RayTraceManager.LoadSamples (GPU, rar_prepare_clips) == LoadSample (the C#'s CPU arithmetic), clip by clip."""
    m = RayTraceManager(context=ctx)
    m.sampleRate = 48000
    rng = np.random.default_rng(5)
    clips = [AudioClip(rng.uniform(-1, 1, 4410 * 2).astype(np.float32), 2, 44100),
             AudioClip(rng.uniform(-1, 1, 3000).astype(np.float32), 1, 48000),
             AudioClip(rng.uniform(-1, 1, 4410 * 2).astype(np.float32), 2, 44100),
             AudioClip(rng.uniform(-1, 1, 2205 * 3).astype(np.float32), 3, 22050)]
    got = m.LoadSamples(clips)
    for c, g in zip(clips, got):
        assert np.array_equal(g, m.LoadSample(c))


def test_streaming_manager_with_a_running_convolver(ctx, oracle):
    """RayTraceManagerStreaming: the chunk cadence of the reference over one running partitioned convolver whose
    response is switched, with a one-block cross-fade, to the IR accumulated since the previous chunk.  Replay the
    schedule with the oracle: block g is the direct-form convolution of the whole input so far with the response
    of the moment (blend of old and new in the first block after a switch)."""
    from realisticaudioraytracing2d_b200.host.ray_trace_manager import RayTraceManagerStreaming
    m = _manager(ctx, RayTraceManagerStreaming)
    m.rayCount, m.reverbDuration = 2000, 0.25
    n_ir = 12000
    m.audioManager = AudioManager(outputSampleRate=48000)
    clip = scenes.synthetic_clip(12000)
    m.inputClip = AudioClip(clip, 1, 48000)
    m.loop = False
    m.Start()
    m.StartStreaming()
    sc = scenes.smoll_room()
    hist = np.zeros(n_ir, np.int64)
    frames_in_ir = 0
    irs, fed = [], []                                               # response and input length after each chunk
    for step in range(18):
        m.Update()
        P = oracle.make_params(source_x=-18.0, source_y=9.0, listener_x=0.0, listener_y=-3.68, ray_count=2000,
                               max_bounce_count=5, rng_state_offset=m.frameCount, impulse_length=n_ir, debug_ray_count=100)
        oracle.trace(sc.walls.view(oracle.SEGMENT_DTYPE), P, hist=hist)
        frames_in_ir += 1
        before = m.nextStreamingOffset
        m.FixedUpdate()
        if m.nextStreamingOffset != before and before < len(clip):
            irs.append(oracle.ir_to_float(hist) / np.float32(max(1, frames_in_ir)))
            fed.append(min(before + 4800, len(clip)))
            hist[:] = 0
            frames_in_ir = 0
    assert fed == [4800, 9600, 12000] and m._out_pos == (12000 // 256) * 256 and len(m._carry) == 12000 % 256
    m.DrainTail()
    x = np.concatenate([clip, np.zeros(m._out_pos - len(clip), np.float32)])
    w = (np.arange(256, dtype=np.float32) + 1) / 256
    full = [oracle.convolve(x, h, 1)[: m._out_pos] for h in irs]
    expect = np.zeros(m._out_pos, np.float32)
    first_block = [0] + [f // 256 for f in fed[:-1]]                 # first block processed under response c
    for g in range(m._out_pos // 256):
        c = max(k for k, fb in enumerate(first_block) if fb <= g)
        blk = slice(g * 256, (g + 1) * 256)
        if c > 0 and g == first_block[c]:
            expect[blk] = full[c - 1][blk] + w * (full[c][blk] - full[c - 1][blk])
        else:
            expect[blk] = full[c][blk]
    got = m.audioManager.ringBuffer[: m._out_pos]
    assert np.abs(expect).max() > 0
    assert rel_l2(got, expect) <= 1e-4
    m.OnDestroy()

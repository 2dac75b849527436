"""LoadSample (RayTraceManager.cs:135-167), SURVEY 8f-3.  CPU part: the C oracle against the independently written
numpy mirror of the same C# text and against hand-computed values; rar_prepared_length (host-only entry point of the
C-ABI) against both.  GPU part: rar_prepare_clips against the oracle, bit for bit."""
import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import _capi
from realisticaudioraytracing2d_b200.host.ray_trace_manager import AudioClip, RayTraceManager

SHAPES = [(1, 1, 44100, 48000), (2, 1, 44100, 48000), (441, 1, 44100, 48000), (1000, 2, 48000, 48000),
          (1000, 2, 22050, 48000), (999, 3, 96000, 48000), (4410, 6, 44100, 48000), (5000, 2, 48000, 44100),
          (12345, 2, 32000, 48000), (7, 1, 8000, 48000), (48000, 2, 47999, 48000),
          (100000, 1, 144000000, 48000), (30000, 2, 192000, 44100), (9000, 1, 4000, 48000)]


def _mirror(raw, samples, channels, freq, rate):
    m = RayTraceManager.__new__(RayTraceManager)                    # LoadSample touches no GPU state
    m.sampleRate = rate
    return m.LoadSample(AudioClip(raw, channels, freq))


def _clip(samples, channels, seed=0):
    return np.random.default_rng(seed).uniform(-1, 1, samples * channels).astype(np.float32)


@pytest.mark.parametrize("samples,channels,freq,rate", SHAPES)
def test_oracle_equals_the_numpy_mirror(oracle, samples, channels, freq, rate):
    raw = _clip(samples, channels, samples)
    want = _mirror(raw, samples, channels, freq, rate)
    got = oracle.load_sample(raw, samples, channels, freq, rate)
    assert got.dtype == np.float32 and np.array_equal(got, want)
    assert _capi.prepared_length(samples, freq, rate) == len(want)


def test_known_answers(oracle):
    stereo = np.array([0.25, 0.75, -1.0, 1.0, 0.5, 0.5], np.float32)
    assert np.array_equal(oracle.load_sample(stereo, 3, 2, 48000, 48000), np.array([0.5, 0.0, 0.5], np.float32))
    # 2:1 decimation picks every other sample exactly (t == 0); 1:2 interpolation alternates samples and midpoints
    x = np.array([1, 2, 4, 8, 16, 32], np.float32)
    assert np.array_equal(oracle.load_sample(x, 6, 1, 96000, 48000), np.array([1, 4, 16], np.float32))
    up = oracle.load_sample(x, 6, 1, 24000, 48000)
    assert np.array_equal(up, np.array([1, 1.5, 2, 3, 4, 6, 8, 12, 16, 24, 32, 32], np.float32))   # idx1 clamps at the end
    # RoundToInt rounds half to even: 5 samples at ratio 2 -> 2.5 -> 2, 7 -> 3.5 -> 4
    assert _capi.prepared_length(5, 96000, 48000) == 2 and _capi.prepared_length(7, 96000, 48000) == 4
    assert _capi.prepared_length(0, 44100, 48000) == 0 and _capi.prepared_length(10, 48000, 48000) == 10
    assert len(oracle.load_sample(np.zeros(7, np.float32), 7, 1, 96000, 48000)) == 4


@pytest.mark.gpu
@pytest.mark.parametrize("samples,channels,freq,rate", SHAPES)
def test_gpu_prepare_clips_bit_exact(ctx, oracle, samples, channels, freq, rate):
    n_clips = 3
    raw = _clip(samples * n_clips, channels, 7 + samples)
    got = ctx.prepare_clips(raw, samples, channels, freq, rate, n_clips)
    per = samples * channels
    for k in range(n_clips):
        want = oracle.load_sample(raw[k * per:(k + 1) * per], samples, channels, freq, rate)
        assert got.shape == (n_clips, len(want)) and np.array_equal(got[k], want)


@pytest.mark.gpu
def test_gpu_prepare_clips_edges_and_errors(ctx):
    assert ctx.prepare_clips(np.zeros(0, np.float32), 0, 2, 44100, 48000, 4).shape == (4, 0)
    assert ctx.prepare_clips(np.zeros(0, np.float32), 10, 2, 44100, 48000, 0).shape == (0, 11)
    lib, h = ctx._lib, ctx._h
    buf = np.zeros(64, np.float32)
    assert lib.rar_prepare_clips(h, buf.ctypes.data, 8, 0, 44100, 48000, 1, buf.ctypes.data, 64) == -1     # channels
    assert lib.rar_prepare_clips(h, buf.ctypes.data, 8, 1, 0, 48000, 1, buf.ctypes.data, 64) == -1         # frequency
    assert lib.rar_prepare_clips(h, buf.ctypes.data, 8, 1, 24000, 48000, 1, buf.ctypes.data, 15) == -1     # stride < 16
    assert lib.rar_prepare_clips(h, None, 8, 1, 24000, 48000, 1, buf.ctypes.data, 16) == -1
    # a stride larger than the prepared length leaves the padding untouched
    out = np.full((2, 20), 9.0, np.float32)
    raw = np.arange(16, dtype=np.float32)
    assert lib.rar_prepare_clips(h, raw.ctypes.data, 8, 1, 24000, 48000, 2, out.ctypes.data, 20) == 0
    assert (out[:, 16:] == 9.0).all() and out[1, 0] == 8.0 and out[0, 1] == 0.5


@pytest.mark.gpu
def test_gpu_prepare_clips_device_arrays(ctx, oracle):
    """rar_prepare_clips_device on arrays already resident in HBM; a misaligned stereo array is refused."""
    import torch
    samples, ch, freq, rate, n_clips = 5000, 2, 44100, 48000, 4
    raw = _clip(samples * n_clips, ch, 3)
    n = _capi.prepared_length(samples, freq, rate)
    d_raw = torch.from_numpy(np.concatenate([raw, np.zeros(2, np.float32)])).cuda()
    d_out = torch.zeros((n_clips, n + 7), dtype=torch.float32, device="cuda")
    ctx.prepare_clips_device(d_raw.data_ptr(), samples, ch, freq, rate, n_clips, d_out.data_ptr(), n + 7)
    ctx.sync()
    got = d_out.cpu().numpy()
    for k in range(n_clips):
        want = oracle.load_sample(raw[k * samples * ch:(k + 1) * samples * ch], samples, ch, freq, rate)
        assert np.array_equal(got[k, :n], want) and not got[k, n:].any()
    with pytest.raises(_capi.RarError) as e:
        ctx.prepare_clips_device(d_raw.data_ptr() + 4, samples, ch, freq, rate, n_clips, d_out.data_ptr(), n + 7)
    assert e.value.code == -1


@pytest.mark.gpu
def test_gpu_prepare_clips_full_size_batch(ctx, oracle):
    """64 stereo clips of 10 s at 44.1 kHz -> 48 kHz (the clip shape of the bench leg), every clip against the oracle."""
    samples, n_clips = 441000, 64
    raw = _clip(samples * n_clips, 2, 99)
    got = ctx.prepare_clips(raw, samples, 2, 44100, 48000, n_clips)
    assert got.shape == (n_clips, 480000)
    for k in range(n_clips):
        assert np.array_equal(got[k], oracle.load_sample(raw[k * samples * 2:(k + 1) * samples * 2], samples, 2, 44100, 48000))

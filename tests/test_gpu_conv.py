"""GPU parity of the partitioned overlap-save FFT convolution against the oracle's direct form
(AudioConvolve.compute:13-31), through the C-ABI.  Bar: output within 1e-4 relative L2."""
import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import _capi, scenes
from tests.common import capi_params, oracle_params, oracle_walls, rel_l2, trace_kwargs

pytestmark = pytest.mark.gpu

TOL = 1e-4  # north star: convolved audio <= 1e-4 relative L2


def _ir(n, seed=3):
    return scenes.decaying_noise_ir(n, seed, decay_s=0.4)


@pytest.mark.parametrize("n_in,n_ir", [(4800, 72000), (42624, 72000), (1000, 300), (256, 256), (257, 255), (5000, 1), (1, 1000), (3, 7)])
def test_convolve_matches_direct_form(ctx, oracle, n_in, n_ir):
    rng = np.random.default_rng(n_in * 31 + n_ir)
    x = rng.uniform(-0.5, 0.5, n_in).astype(np.float32)
    h = _ir(n_ir)
    ctx.ir_write(0, h)
    hq = ctx.ir_read(0, n_ir)                      # the IR as the slot holds it (Q23.40 quantised)
    assert np.abs(hq - h).max() <= 2.0 ** -40
    got = ctx.convolve(0, x, accum_count=3, ir_len=n_ir)
    want = oracle.convolve(x, hq, 3)
    assert got.shape == want.shape == (n_in + n_ir,)
    assert rel_l2(got, want) <= TOL
    assert got[-1] == 0.0                          # output has N+M samples, the last is always 0


def test_epsilon_gate_and_accum_semantics(ctx, oracle):
    rng = np.random.default_rng(5)
    h = _ir(2000)
    ctx.ir_write(0, h)
    hq = ctx.ir_read(0, 2000)
    x = rng.uniform(-0.5, 0.5, 3000).astype(np.float32)
    x[::3] = rng.uniform(-1e-4, 1e-4, len(x[::3])).astype(np.float32)   # |x| <= eps contributes nothing (:25)
    got = ctx.convolve(0, x, 2, 2000)
    assert rel_l2(got, oracle.convolve(x, hq, 2)) <= TOL
    xz = x.copy()
    xz[np.abs(xz) <= np.float32(1e-4)] = 0
    assert np.array_equal(got, ctx.convolve(0, xz, 2, 2000))
    # accumCount <= 0 writes zeros (:30)
    assert not ctx.convolve(0, x, 0, 2000).any()
    # an impulse reproduces ir/accumCount, a delayed impulse shifts it
    imp = np.zeros(700, np.float32)
    imp[0] = 1.0
    y = ctx.convolve(0, imp, 4, 2000)
    assert rel_l2(y[:2000], hq / 4) <= TOL
    imp = np.zeros(700, np.float32)
    imp[333] = 1.0
    y = ctx.convolve(0, imp, 1, 2000)
    assert rel_l2(y[333:2333], hq) <= TOL and np.abs(y[:333]).max() <= 1e-6 * np.abs(hq).max()


def test_async_tickets(ctx, oracle):
    rng = np.random.default_rng(9)
    h = _ir(5000)
    ctx.ir_write(1, h)
    hq = ctx.ir_read(1, 5000)
    xs = [rng.uniform(-0.5, 0.5, 4800).astype(np.float32) for _ in range(3)]
    tickets = [ctx.convolve_begin(1, x, 2) for x in xs]
    assert len(set(tickets)) == 3
    import time
    t0 = time.time()
    while not all(ctx.poll(t) for t in tickets):
        assert time.time() - t0 < 30
    for t, x in zip(tickets, xs):
        assert rel_l2(ctx.convolve_end(t, 9800), oracle.convolve(x, hq, 2)) <= TOL
    with pytest.raises(_capi.RarError):
        ctx.poll(tickets[0])                       # retired


def test_trace_then_convolve_config1(ctx, oracle):
    """BASELINE config 1 end to end: bundled room, 3 accumulated frames, IR build + convolution of one clip."""
    sc = scenes.smoll_room()
    kw = trace_kwargs(sc)
    n = kw["impulse_length"]
    ctx.set_walls(sc.walls)
    ctx.ir_clear(0, n, 1)
    hist = np.zeros(n, np.int64)
    for f in (11, 12, 13):
        ctx.trace(capi_params(_capi, dict(kw, rng_state_offset=f)), 0)
        oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, dict(kw, rng_state_offset=f)), hist=hist)
    ir = ctx.ir_read(0, n)
    assert np.array_equal(ir.view(np.uint32), oracle.ir_to_float(hist).view(np.uint32))
    clip = scenes.synthetic_clip()
    got = ctx.convolve(0, clip, 3, n)
    want = oracle.convolve(clip, ir, 3)
    assert rel_l2(got, want) <= TOL


def test_streaming_convolver_matches_direct_form(ctx, oracle):
    """Config 5 at reduced size: 6 streams with different IR lengths, 40 blocks of 256."""
    rng = np.random.default_rng(21)
    S, B, n_blocks, max_ir = 6, 256, 40, 5000
    cv = _capi.Convolver(ctx, S, B, max_ir)
    lens = [5000, 4097, 256, 1, 300, 2048]
    irs = [_ir(L, seed=s) for s, L in enumerate(lens)]
    for s, h in enumerate(irs):
        cv.set_ir(s, h, scale=0.5)
    x = rng.uniform(-1, 1, (S, n_blocks * B)).astype(np.float32)
    x[:, ::11] *= 1e-5
    y = np.concatenate([cv.process(x[:, k * B:(k + 1) * B]) for k in range(n_blocks)], axis=1)
    for s in range(S):
        want = oracle.convolve(x[s], irs[s] * np.float32(0.5), 1)[: n_blocks * B]
        assert rel_l2(y[s], want) <= TOL, s
    # reset restarts the streams
    cv.reset()
    y2 = cv.process(x[:, :B])
    assert np.array_equal(y2, y[:, :B])
    cv.destroy()


def test_streaming_convolver_from_traced_slot(ctx, oracle):
    sc = scenes.smoll_room()
    kw = trace_kwargs(sc, impulse_length=24000)
    ctx.set_walls(sc.walls)
    ctx.ir_clear(0, 24000, 1)
    ctx.trace(capi_params(_capi, kw), 0)
    ir = ctx.ir_read(0, 24000)
    cv = _capi.Convolver(ctx, 2, 256, 24000)
    cv.set_ir_from_slot(0, 0, 1)
    cv.set_ir_from_slot(1, 0, 2)
    rng = np.random.default_rng(2)
    x = rng.uniform(-1, 1, (2, 100 * 256)).astype(np.float32)
    y = np.concatenate([cv.process(x[:, k * 256:(k + 1) * 256]) for k in range(100)], axis=1)
    assert rel_l2(y[0], oracle.convolve(x[0], ir, 1)[: 100 * 256]) <= TOL
    assert rel_l2(y[1], oracle.convolve(x[1], ir, 2)[: 100 * 256]) <= TOL
    cv.destroy()


def test_streaming_full_size_properties(ctx):
    """Config 5 shape at full partition count on a few streams (10 s IR = 1875 partitions):
    impulse in -> IR out (block by block), and linearity."""
    S, B, n_ir = 4, 256, 480000
    cv = _capi.Convolver(ctx, S, B, n_ir)
    irs = [_ir(n_ir, seed=40 + s) for s in range(S)]
    for s in range(S):
        cv.set_ir(s, irs[s])
    x0 = np.zeros((S, B), np.float32)
    x0[:, 5] = 1.0
    outs = [cv.process(x0)]
    z = np.zeros((S, B), np.float32)
    for _ in range(30):
        outs.append(cv.process(z))
    y = np.concatenate(outs, axis=1)
    for s in range(S):
        assert rel_l2(y[s, 5:], irs[s][: y.shape[1] - 5]) <= TOL
    cv.destroy()


def test_streaming_config5_full_size_determinism_and_linearity(ctx):
    """BASELINE config 5 at full size (256 streams x 480 000-tap IRs, 1875 partitions): the reduction order is
    fixed, so two runs are bit-identical; and the convolver is linear in its input."""
    S, B, n_ir = 256, 256, 480000
    cv = _capi.Convolver(ctx, S, B, n_ir)
    base = _ir(n_ir, seed=77)
    for s in range(S):
        cv.set_ir(s, np.roll(base, 37 * s))
    rng = np.random.default_rng(8)
    x1 = rng.uniform(-1, 1, (3, S, B)).astype(np.float32)
    x2 = rng.uniform(-1, 1, (3, S, B)).astype(np.float32)

    def run(x):
        cv.reset()
        return np.stack([cv.process(x[k]) for k in range(3)])

    y1 = run(x1)
    assert np.array_equal(run(x1), y1)                       # deterministic
    y2, y12 = run(x2), run(x1 + x2)
    assert rel_l2(y12, y1 + y2) <= 1e-5                      # linear
    # first block of stream 0: plain convolution with the head of its IR
    want = np.convolve(x1[0, 0].astype(np.float64), base[:B].astype(np.float64))[:B]
    assert rel_l2(y1[0, 0], want) <= TOL
    assert cv.bytes_per_block() > 1.9e9
    cv.destroy()


def test_streaming_ir_update_with_crossfade(ctx, oracle):
    """rar_conv_update_ir: the block after an update is y_old + w (y_new - y_old), w[n] = (n+1)/256,
    with both responses applied to the same input history; later blocks use the new response; streams that were not
    updated are untouched.  Expected values come from the oracle's direct-form convolution of the whole history."""
    rng = np.random.default_rng(33)
    S, B, n_blocks, max_ir = 6, 256, 24, 3000
    cv = _capi.Convolver(ctx, S, B, max_ir)
    old = [_ir(L, seed=10 + s) for s, L in enumerate([3000, 1500, 700, 256, 2049, 900])]
    new = [_ir(L, seed=50 + s) for s, L in enumerate([2000, 3000, 1, 300, 2049, 900])]
    for s in range(S):
        cv.set_ir(s, old[s])
    x = rng.uniform(-1, 1, (S, n_blocks * B)).astype(np.float32)
    w = (np.arange(B, dtype=np.float32) + 1) / B
    current = list(old)
    fades = {5: [1, 3], 6: [1], 12: [0, 1, 2, 3, 4], 13: [5]}       # block -> streams updated just before it
    out = []
    for k in range(n_blocks):
        fading = fades.get(k, [])
        for s in fading:
            if k == 6:
                cv.update_ir(s, old[0])                             # replaced before the block is processed ...
            cv.update_ir(s, new[s] if current[s] is old[s] else old[s])
        if k == 13:
            cv.update_ir(4, old[0])
            cv.set_ir(4, current[4])                                # ... and a hard set cancels a pending update
        y = cv.process(x[:, k * B:(k + 1) * B])
        for s in range(S):
            hist = x[s, : (k + 1) * B]
            y_old = oracle.convolve(hist, current[s], 1)[k * B:(k + 1) * B]
            if s in fading:
                nxt = new[s] if current[s] is old[s] else old[s]
                y_new = oracle.convolve(hist, nxt, 1)[k * B:(k + 1) * B]
                want = y_old + w * (y_new - y_old)
                current[s] = nxt
            else:
                want = y_old
            assert rel_l2(y[s], want) <= TOL, (k, s)
        out.append(y)
    cv.destroy()


def test_streaming_ir_update_from_traced_slot(ctx, oracle):
    """The ping/pong cadence of the reference's streaming path inside the convolver: the IR traced into the other
    slot replaces the current one with a one-block cross-fade."""
    sc = scenes.smoll_room()
    n = 6000
    ctx.set_walls(sc.walls)
    irs = []
    for slot, frame in ((0, 1), (1, 2)):
        ctx.ir_clear(slot, n, 1)
        ctx.trace(capi_params(_capi, trace_kwargs(sc, impulse_length=n, rng_state_offset=frame)), slot)
        irs.append(ctx.ir_read(slot, n))
    cv = _capi.Convolver(ctx, 1, 256, n)
    cv.set_ir_from_slot(0, 0, 1)
    x = np.random.default_rng(4).uniform(-1, 1, (1, 40 * 256)).astype(np.float32)
    w = (np.arange(256, dtype=np.float32) + 1) / 256
    for k in range(40):
        if k == 20:
            cv.update_ir_from_slot(0, 1, 1)
        y = cv.process(x[:, k * 256:(k + 1) * 256])[0]
        hist = x[0, : (k + 1) * 256]
        a = oracle.convolve(hist, irs[0], 1)[k * 256:(k + 1) * 256]
        b = oracle.convolve(hist, irs[1], 1)[k * 256:(k + 1) * 256]
        want = a if k < 20 else (a + w * (b - a) if k == 20 else b)
        assert rel_l2(y, want) <= TOL, k
    cv.destroy()


def test_batched_response_load_equals_per_stream_loads(ctx, oracle):
    """rar_conv_set_irs (one call, alternating pinned staging halves, no stream synchronisation) against rar_conv_set_ir
    per stream; the responses are long enough for several staging groups."""
    rng = np.random.default_rng(12)
    S, n = 24, 400_000                                   # 24 x 1.6 MB: three 16 MB staging groups
    irs = (rng.standard_normal((S, n)) * np.exp(-np.arange(n) / 50_000.0)).astype(np.float32) * np.float32(0.02)
    x = rng.uniform(-1, 1, (S, 256 * 3)).astype(np.float32)
    out = []
    for batched in (True, False):
        cv = _capi.Convolver(ctx, S, 256, n)
        try:
            if batched:
                cv.set_irs(0, irs, 0.5)
                irs_after = irs.copy()                    # the call has copied: the caller's array is free again
            else:
                for s in range(S):
                    cv.set_ir(s, irs[s], 0.5)
            out.append(np.concatenate([cv.process(x[:, k * 256:(k + 1) * 256]) for k in range(3)], axis=1))
        finally:
            cv.destroy()
    assert np.array_equal(out[0], out[1]) and out[0].any()
    assert np.array_equal(irs, irs_after)
    want = oracle.convolve(x[5], irs[5] * np.float32(0.5), 1)[: 256 * 3]
    assert rel_l2(out[0][5], want) <= 1e-4


def test_streaming_convolver_writes_straight_into_rings(ctx, oracle):
    """rar_conv_process_to_ring: every block of stream s lands in rings[s] at its sample offset; draining the ring
    gives the same samples rar_conv_process returns."""
    rng = np.random.default_rng(21)
    S, n, blocks = 3, 2000, 12
    irs = (rng.standard_normal((S, n)) * 0.05).astype(np.float32)
    x = rng.uniform(-1, 1, (S, 256 * blocks)).astype(np.float32)
    outs = []
    for to_ring in (False, True):
        cv = _capi.Convolver(ctx, S, 256, n)
        rings = [_capi.Ring(48000, 0.5), None, _capi.Ring(48000, 0.5)]
        try:
            cv.set_irs(0, irs)
            if to_ring:
                for k in range(blocks):
                    cv.process_to_rings(x[:, k * 256:(k + 1) * 256], rings, k * 256)
                y = np.zeros((S, 256 * blocks), np.float32)
                for s in (0, 2):
                    rings[s].drain(y[s], 1)
                assert rings[0].pinned
            else:
                y = np.concatenate([cv.process(x[:, k * 256:(k + 1) * 256]) for k in range(blocks)], axis=1)
            outs.append(y)
        finally:
            cv.destroy()
            for r in rings:
                if r is not None:
                    r.destroy()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][2], outs[1][2]) and not outs[1][1].any()
    assert rel_l2(outs[1][2], oracle.convolve(x[2], irs[2], 1)[: 256 * blocks]) <= 1e-4

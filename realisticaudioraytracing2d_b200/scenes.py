"""Scene fixtures and synthetic workloads for the hot path.

`smoll_room` / `big_room` rebuild the reference's two live scenes from the transforms serialized in
Assets/Scenes/SmollRoom.unity and "Assets/Scenes/Big Room.unity" through the host mirror of
SceneToData2D (so the binary32 segment values are generated, not typed in).  `shoebox` and `maze`
are the synthetic BASELINE.json configs 2 and 3.  Everything here is deterministic.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .host.scene_helper import (SEGMENT_DTYPE, AcousticSurface, AudioMaterial, BoxCollider2D, GameObject,
                                SceneToData2D, Transform)

# Assets/Script/Border.asset:15-18 and Assets/Script/Material.asset:15-18
BORDER = AudioMaterial(absorption=0.507, scattering=0.5, transmission=0.271, ior=0.01)
MATERIAL = AudioMaterial(absorption=0.148, scattering=1.0, transmission=1.0, ior=0.6)


@dataclass
class Scene:
    name: str
    walls: np.ndarray                      # SEGMENT_DTYPE[n]
    source: tuple
    listener: tuple
    listener_radius: float = 0.5
    speed_of_sound: float = 343.0
    input_gain: float = 1.0
    ray_count: int = 1000
    max_bounces: int = 5
    sample_rate: int = 48000
    reverb_duration: float = 2.0
    band_absorption: Optional[np.ndarray] = None   # float32 [n, bands] (build extension, config 3)
    extra: dict = field(default_factory=dict)

    @property
    def impulse_length(self) -> int:
        # RayTraceManager.cs:172,181,214: (int)(sampleRate * reverbDuration) in binary32
        return int(np.float32(self.sample_rate) * np.float32(self.reverb_duration))


def _box(name, pos, zw, scale, mat) -> GameObject:
    return GameObject(Transform(pos, zw, scale), BoxCollider2D(size=(1.0, 1.0), offset=(0.0, 0.0)),
                      AcousticSurface(mat), name)


_ROT90 = (0.7071068, 0.7071068)      # SmollRoom.unity:554,1236
_ROT57 = (0.47792548, 0.8784004)     # SmollRoom.unity:803


def smoll_room() -> Scene:
    """Assets/Scenes/SmollRoom.unity: obstacleObjects order :169-174, transforms :902-904, :410-412,
    :1236-1238, :554-556, :803-805; manager parameters :155-178; source :1334, listener :673."""
    objs = [
        _box("Wall", (0.0, 10.0), (0.0, 1.0), (100.0, 1.0), BORDER),
        _box("Wall (1)", (0.01, -5.0), (0.0, 1.0), (100.0, 1.0), BORDER),
        _box("Wall (2)", (-20.0, 0.0), _ROT90, (20.0, 1.0), BORDER),
        _box("Wall (3)", (20.0, 0.0), _ROT90, (20.0, 1.0), BORDER),
        _box("Wall (4)", (-11.8, 7.18), _ROT57, (100.0, 1.0), MATERIAL),
    ]
    return Scene("SmollRoom", SceneToData2D.GetSegmentsFromColliders(objs), source=(-18.0, 9.0),
                 listener=(0.0, -3.68), listener_radius=0.5, speed_of_sound=343.0, input_gain=1.0,
                 ray_count=15000, max_bounces=5, sample_rate=48000, reverb_duration=1.5)


def big_room() -> Scene:
    """"Assets/Scenes/Big Room.unity": transforms :902-904, :410-412, :1139-1141, :554-556, :803-805;
    inputGain 100 (:161); source :1237, listener :673."""
    objs = [
        _box("Wall", (0.0, 100.0), (0.0, 1.0), (1000.0, 1.0), BORDER),
        _box("Wall (1)", (0.01, -50.0), (0.0, 1.0), (1000.0, 1.0), BORDER),
        _box("Wall (2)", (-200.0, 0.0), _ROT90, (200.0, 1.0), BORDER),
        _box("Wall (3)", (200.0, 0.0), _ROT90, (200.0, 1.0), BORDER),
        _box("Wall (4)", (-118.8, 71.8), _ROT57, (1000.0, 10.0), MATERIAL),
    ]
    return Scene("Big Room", SceneToData2D.GetSegmentsFromColliders(objs), source=(-183.8, 87.1),
                 listener=(0.0, -3.68), listener_radius=0.5, speed_of_sound=343.0, input_gain=100.0,
                 ray_count=15000, max_bounces=5, sample_rate=48000, reverb_duration=1.5)


def _segment(a, b, normal, mat: AudioMaterial):
    return (np.asarray(a, np.float32), np.asarray(b, np.float32), np.asarray(normal, np.float32),
            np.float32(mat.absorption), np.float32(mat.scattering), np.float32(mat.transmission), np.float32(mat.ior))


def shoebox(width: float = 10.0, height: float = 6.0, absorption: float = 0.1, scattering: float = 0.0,
            transmission: float = 0.0, ior: float = 1.0, ray_count: int = 1 << 20, max_bounces: int = 32,
            reverb_duration: float = 1.0) -> Scene:
    """BASELINE.json config 2: an axis-aligned rectangular room of 4 segments.  Normals point INTO the
    room, as the room-facing faces of the reference's box walls do (their outward normal is the room's
    inward one); with normals pointing out of the room the shadow ray of Raytrace2D.compute:105 would
    start behind the wall it just hit and every next-event estimate would be blocked."""
    mat = AudioMaterial(absorption, scattering, transmission, ior)
    w, h = width, height
    segs = [
        _segment((0, 0), (w, 0), (0, 1), mat),
        _segment((w, 0), (w, h), (-1, 0), mat),
        _segment((w, h), (0, h), (0, -1), mat),
        _segment((0, h), (0, 0), (1, 0), mat),
    ]
    walls = np.zeros(4, dtype=SEGMENT_DTYPE)
    for i, s in enumerate(segs):
        walls[i] = s
    return Scene(f"shoebox {w:g}x{h:g}", walls, source=(2.5, 1.7), listener=(7.3, 4.1), listener_radius=0.5,
                 ray_count=ray_count, max_bounces=max_bounces, reverb_duration=reverb_duration)


def maze(n_segments: int = 10000, size: float = 100.0, bands: int = 8, seed: int = 1234,
         ray_count: int = 1 << 26, max_bounces: int = 64, reverb_duration: float = 1.0,
         scattering: float = 0.1) -> Scene:
    """BASELINE.json config 3: `n_segments` axis-aligned thin walls on a square grid inside a closed
    bounding box (4 of the segments), so no ray escapes.  Interior cell edges are chosen without
    replacement by a seeded generator at roughly 40 % density; per-wall, per-band absorption is drawn
    uniform(0.02, 0.10) and the broadband absorption is the band mean (<= 0.10 keeps 64 bounces above
    the 1e-3 cut-off of Raytrace2D.compute:122)."""
    rng = np.random.default_rng(seed)
    n_inner = n_segments - 4
    g = 2
    while 2 * g * (g - 1) * 0.4 < n_inner:
        g += 1
    cell = np.float32(size / g)
    n_edges = 2 * g * (g - 1)
    chosen = np.sort(rng.choice(n_edges, size=n_inner, replace=False))
    band_abs = rng.uniform(0.02, 0.10, size=(n_segments, max(bands, 1))).astype(np.float32)
    flip = rng.integers(0, 2, size=n_segments)
    walls = np.zeros(n_segments, dtype=SEGMENT_DTYPE)
    s = np.float32(size)
    outer = [((0, 0), (s, 0), (0, 1)), ((s, 0), (s, s), (-1, 0)), ((s, s), (0, s), (0, -1)), ((0, s), (0, 0), (1, 0))]
    for i, (a, b, n) in enumerate(outer):
        m = AudioMaterial(float(band_abs[i].mean(dtype=np.float32)), scattering, 0.0, 1.0)
        walls[i] = _segment(a, b, n, m)
    half = g * (g - 1)
    for k, e in enumerate(chosen):
        i = 4 + k
        if e < half:   # vertical edge between cell (cx, cy) and (cx+1, cy)
            cx, cy = e % (g - 1), e // (g - 1)
            x = np.float32(cx + 1) * cell
            a, b = (x, np.float32(cy) * cell), (x, np.float32(cy + 1) * cell)
            n = (1.0, 0.0) if flip[i] else (-1.0, 0.0)
        else:          # horizontal edge between cell (cx, cy) and (cx, cy+1)
            e2 = e - half
            cx, cy = e2 % g, e2 // g
            y = np.float32(cy + 1) * cell
            a, b = (np.float32(cx) * cell, y), (np.float32(cx + 1) * cell, y)
            n = (0.0, 1.0) if flip[i] else (0.0, -1.0)
        m = AudioMaterial(float(band_abs[i].mean(dtype=np.float32)), scattering, 0.0, 1.0)
        walls[i] = _segment(a, b, n, m)
    c = float(cell)
    src = (c * (g // 3 + 0.5), c * (g // 3 + 0.47))
    lis = (c * (g // 3 + 2.5), c * (g // 3 + 1.53))
    return Scene(f"maze {n_segments} segs seed {seed}", walls, source=src, listener=lis, listener_radius=0.5,
                 ray_count=ray_count, max_bounces=max_bounces, reverb_duration=reverb_duration,
                 band_absorption=band_abs if bands > 1 else None, extra={"grid": g, "cell": c, "bands": bands})


def synthetic_clip(n_samples: int = 42624, seed: int = 7, amplitude: float = 0.5) -> np.ndarray:
    """Stand-in for Assets/Script/bruh.mp3 (48 kHz, ~0.888 s): no MP3 decoder exists in the image, so
    config 1 uses uniform(-amplitude, amplitude) noise of the same length."""
    rng = np.random.default_rng(seed)
    return rng.uniform(-amplitude, amplitude, size=n_samples).astype(np.float32)


def decaying_noise_ir(n_taps: int, seed: int, decay_s: float = 1.5, sample_rate: int = 48000) -> np.ndarray:
    """Config 5 impulse responses: exponentially decaying seeded noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_taps, dtype=np.float32) / np.float32(sample_rate)
    env = np.exp(-t * np.float32(6.9 / decay_s)).astype(np.float32)
    return (rng.standard_normal(n_taps).astype(np.float32) * env * np.float32(0.05)).astype(np.float32)

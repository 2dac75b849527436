"""Host mirror of the reference's scene-ingestion types (Assets/Script/Helpers/SceneHelper.cs,
AudioMaterial.cs, AudioSurface.cs).

The reference walks Unity `Collider2D` components; Unity is not available here, so the small
classes below stand in for `Transform`, `BoxCollider2D`, `CircleCollider2D`, `PolygonCollider2D`,
`GameObject`, `AcousticSurface` and `AudioMaterial` with the fields the walk actually reads.
`SceneToData2D.GetSegmentsFromColliders` keeps the reference's name, argument meaning, segment
order and binary32 arithmetic (Helpers/SceneHelper.cs:29-110).  This is host-side, O(segments)
work: it is not accelerated, it only produces the 40-byte `Segment` records the CUDA path uploads.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

f32 = np.float32

# Helpers/SceneHelper.cs:15-22 `Segment` (LayoutKind.Sequential) with the nested `AudioMat` (:8-14):
# start, end, normal (Vector2 each) then absorption, scattering, transmission, ior.  40 bytes.
SEGMENT_DTYPE = np.dtype(
    [("start", "<f4", (2,)), ("end", "<f4", (2,)), ("normal", "<f4", (2,)),
     ("absorption", "<f4"), ("scattering", "<f4"), ("transmission", "<f4"), ("ior", "<f4")]
)
assert SEGMENT_DTYPE.itemsize == 40

CIRCLE_RESOLUTION = 32  # Helpers/SceneHelper.cs:26


@dataclass
class AudioMaterial:
    """AudioMaterial.cs:6-20 (ScriptableObject): defaults and meaning of the four coefficients."""
    absorption: float = 0.1
    scattering: float = 0.5
    transmission: float = 0.0
    ior: float = 1.0


@dataclass
class AcousticSurface:
    """AudioSurface.cs:3-6."""
    material: AudioMaterial = field(default_factory=AudioMaterial)


@dataclass
class Transform:
    """The subset of UnityEngine.Transform the walk uses: world position, z-rotation, lossy scale."""
    position: Sequence[float] = (0.0, 0.0)
    rotation_zw: Sequence[float] = (0.0, 1.0)  # quaternion (0, 0, z, w)
    lossyScale: Sequence[float] = (1.0, 1.0)

    @staticmethod
    def from_degrees(position, degrees, scale) -> "Transform":
        h = math.radians(degrees) * 0.5
        return Transform(position, (math.sin(h), math.cos(h)), scale)

    def TransformPoint(self, p) -> np.ndarray:
        """world = position + Rz(q) * (scale (.) p), evaluated in binary32 without contraction."""
        qz, qw = f32(self.rotation_zw[0]), f32(self.rotation_zw[1])
        two = f32(2.0)
        r00 = f32(1.0) - two * (qz * qz)
        r01 = -(two * (qz * qw))
        r10 = two * (qz * qw)
        r11 = r00
        lx = f32(p[0]) * f32(self.lossyScale[0])
        ly = f32(p[1]) * f32(self.lossyScale[1])
        x = (r00 * lx + r01 * ly) + f32(self.position[0])
        y = (r10 * lx + r11 * ly) + f32(self.position[1])
        return np.array([x, y], dtype=f32)


@dataclass
class Collider2D:
    enabled: bool = True


@dataclass
class BoxCollider2D(Collider2D):
    size: Sequence[float] = (1.0, 1.0)
    offset: Sequence[float] = (0.0, 0.0)


@dataclass
class CircleCollider2D(Collider2D):
    radius: float = 0.5
    offset: Sequence[float] = (0.0, 0.0)


@dataclass
class PolygonCollider2D(Collider2D):
    paths: List[Sequence[Sequence[float]]] = field(default_factory=list)

    @property
    def pathCount(self) -> int:
        return len(self.paths)

    def GetPath(self, i: int):
        return self.paths[i]


@dataclass
class GameObject:
    transform: Transform = field(default_factory=Transform)
    collider: Optional[Collider2D] = None
    surface: Optional[AcousticSurface] = None
    name: str = ""


def _normalized(v: np.ndarray) -> np.ndarray:
    # UnityEngine.Vector2.normalized: v / magnitude when magnitude > 1e-5, else zero.
    mag = f32(np.sqrt(v[0] * v[0] + v[1] * v[1]))
    if mag > f32(1e-5):
        return np.array([v[0] / mag, v[1] / mag], dtype=f32)
    return np.zeros(2, dtype=f32)


class SceneToData2D:
    """Helpers/SceneHelper.cs:24-110."""

    @staticmethod
    def GetSegmentsFromColliders(objects: Sequence[GameObject]) -> np.ndarray:
        segs: list = []
        for obj in objects:
            col = obj.collider
            if col is None or not col.enabled:  # :34
                continue
            mat = SceneToData2D.ResolveMaterial(obj)  # :37
            if isinstance(col, PolygonCollider2D):  # :39-46
                for i in range(col.pathCount):
                    SceneToData2D.AddLoopToSegments(obj.transform, col.GetPath(i), segs, mat)
            elif isinstance(col, BoxCollider2D):  # :47-56
                hx, hy = f32(col.size[0]) * f32(0.5), f32(col.size[1]) * f32(0.5)
                ox, oy = f32(col.offset[0]), f32(col.offset[1])
                pts = [(ox - hx, oy - hy), (ox + hx, oy - hy), (ox + hx, oy + hy), (ox - hx, oy + hy)]
                SceneToData2D.AddLoopToSegments(obj.transform, pts, segs, mat)
            elif isinstance(col, CircleCollider2D):  # :57-67
                pts = []
                for i in range(CIRCLE_RESOLUTION):
                    angle = f32(f32(i) / f32(CIRCLE_RESOLUTION)) * f32(math.pi) * f32(2.0)
                    c, s = f32(math.cos(float(angle))), f32(math.sin(float(angle)))  # Mathf.Cos = (float)Math.Cos
                    pts.append((f32(col.offset[0]) + c * f32(col.radius), f32(col.offset[1]) + s * f32(col.radius)))
                SceneToData2D.AddLoopToSegments(obj.transform, pts, segs, mat)
            else:
                # :68-71 logs "COLLIDER NOT SUPPORTED YET" and carries on
                continue
        out = np.zeros(len(segs), dtype=SEGMENT_DTYPE)
        for i, s in enumerate(segs):
            out[i] = s
        return out

    @staticmethod
    def AddLoopToSegments(trans: Transform, localPoints, outSegments: list, material: AudioMaterial) -> None:
        sx, sy = f32(trans.lossyScale[0]), f32(trans.lossyScale[1])
        winding = f32(1.0) if sx * sy >= 0 else f32(-1.0)  # Mathf.Sign (:81)
        n = len(localPoints)
        for i in range(n):
            p1, p2 = localPoints[i], localPoints[(i + 1) % n]
            start, end = trans.TransformPoint(p1), trans.TransformPoint(p2)  # :89-90
            d = _normalized(end - start)  # :92
            normal = np.array([d[1] * winding, -d[0] * winding], dtype=f32)  # :93
            outSegments.append((start, end, normal, f32(material.absorption), f32(material.scattering),
                                f32(material.transmission), f32(material.ior)))

    @staticmethod
    def ResolveMaterial(obj: GameObject) -> AudioMaterial:
        # :99-110 dereferences surface.material unconditionally (a missing AcousticSurface is a
        # NullReferenceException in the reference); mirror that as an error, not a silent default.
        if obj.surface is None or obj.surface.material is None:
            raise AttributeError(f"GameObject '{obj.name}' has no AcousticSurface.material")
        return obj.surface.material

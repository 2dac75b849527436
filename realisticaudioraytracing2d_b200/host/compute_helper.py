"""Host mirror of the parts of Helpers/ComputeHelper.cs the managers use, re-pointed at the C-ABI.

The reference's helper wraps Unity's ComputeBuffer/ComputeShader API: `Dispatch` (ceil-div group count,
Helpers/ComputeHelper.cs:25-32), `CreateStructuredBuffer` overloads (:83-153), `CreateAppendBuffer`
(:61-81), `Release` (:209-247), `ReadbackData` (:534-539).  With the fused CUDA path there is no hit
append buffer and no per-kernel dispatch, so what remains is: wall upload, IR-slot management, thread
count arithmetic and release.  Names are kept so the manager code reads like the reference's.
"""
from __future__ import annotations

import math

import numpy as np

from .. import _capi


class ComputeHelper:
    @staticmethod
    def GetThreadGroupCount(numIterationsX: int, groupSize: int = 64) -> int:
        """Mathf.CeilToInt(numIterationsX / (float)groupSize) (Helpers/ComputeHelper.cs:27-31).  The
        float division is reproduced because it is what decides how many threads Trace really runs."""
        return int(math.ceil(float(np.float32(numIterationsX) / np.float32(groupSize))))

    @staticmethod
    def DispatchedThreads(numIterationsX: int, groupSize: int = 64) -> int:
        return ComputeHelper.GetThreadGroupCount(numIterationsX, groupSize) * groupSize

    @staticmethod
    def CreateStructuredBuffer(ctx: "_capi.Context", segments: np.ndarray) -> int:
        """CreateStructuredBuffer(ref wallBuffer, activeSegments) (:114-125): (re)allocate + SetData."""
        ctx.set_walls(segments)
        return len(segments)

    @staticmethod
    def CreateIRBuffer(ctx: "_capi.Context", slot: int, count: int, bands: int = 1) -> None:
        """CreateStructuredBuffer<float>(ref irBuffer, len) (:83-95) + ClearImpulse: slots are always
        zero-initialised here (the reference leaves new buffers undefined until the first ResetIR)."""
        ctx.ir_clear(slot, count, bands)

    @staticmethod
    def ReadbackData(ctx: "_capi.Context", slot: int, count: int) -> np.ndarray:
        """ReadbackData<float>(buffer) (:534-539)."""
        return ctx.ir_read(slot, count)

    @staticmethod
    def Release(*objects) -> None:
        """Release(params ComputeBuffer[]) (:209-226): null-tolerant."""
        for o in objects:
            if o is not None and hasattr(o, "destroy"):
                o.destroy()

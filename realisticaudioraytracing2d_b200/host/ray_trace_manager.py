"""Host mirror of Assets/Script/RayTraceManager.cs (the live orchestrator) with its GPU calls
re-pointed at the C-ABI (include/rar2d.h) instead of Unity compute shaders.

Public fields and method names are the reference's (RayTraceManager.cs:8-34 and :45-281); Unity's
frame loop is replaced by the caller invoking Start()/Update()/FixedUpdate().  Differences that follow
from the fused CUDA path, all documented in DESIGN.md:
  * Trace + ProcessHits are one call, so `accumFrames` advances in RunSimulation, not 1-3 frames later
    in an AsyncGPUReadback callback (:209, :233);
  * IR slots hold 64-bit fixed point; `GetActiveIRBuffer` returns a slot index;
  * DrawIR / OnGUI / OnDrawGizmos (:235-243, :252-279) are editor visualisation and are not mirrored
    (debug ray paths are still available through GetDebugRayPaths).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .. import _capi
from .audio_manager import AudioManager
from .compute_helper import ComputeHelper
from .scene_helper import GameObject, SceneToData2D, Transform


@dataclass
class AudioClip:
    """The subset of UnityEngine.AudioClip that LoadSample reads (:137-150)."""
    data: np.ndarray            # interleaved float32, samples * channels
    channels: int = 1
    frequency: int = 48000

    @property
    def samples(self) -> int:
        return len(self.data) // self.channels

    def GetData(self) -> np.ndarray:
        return np.asarray(self.data, dtype=np.float32)


def _round_to_int(x) -> int:
    """Mathf.RoundToInt: round half to even."""
    return int(np.rint(np.float32(x)))


class RayTraceManager:
    def __init__(self, device: int = 0, context: Optional[_capi.Context] = None):
        # [Header("Simulation")] :12-16
        self.rayCount = 1000
        self.maxBounces = 5
        self.speedOfSound = 343.0
        self.dynamicObstacles = False
        # [Header("Audio")] :18-24
        self.inputClip: Optional[AudioClip] = None
        self.audioManager: Optional[AudioManager] = None
        self.sampleRate = 48000
        self.inputGain = 1.0
        self.reverbDuration = 2.0
        self.loop = True
        # [Header("Scene")] :26-29
        self.source: Optional[Transform] = None
        self.listener: Optional[Transform] = None
        self.listenerRadius = 0.5
        self.obstacleObjects: List[GameObject] = []
        # [Header("Debug")] :31-34
        self.debugRayCount = 100
        # not in the reference: wall count from which traces use RAR_FLAG_USE_GRID (results are identical)
        self.gridThreshold = 64
        # private state :36-41
        self.activeSegments = None
        self.fullInputSamples = None
        self.activeIRIndex = 0
        self.accumFrames = 0
        self.samplesSinceLastChunk = 0
        self.chunkSamples = 0
        self.nextStreamingOffset = 0
        # stand-ins for UnityEngine.Time
        self.frameCount = 0
        self.fixedDeltaTime = 0.02          # ProjectSettings/TimeManager.asset:6
        self._owns_ctx = context is None
        self._ctx = context if context is not None else _capi.Context(device)
        self._coroutines = []
        self._slot_len = [-1, -1]

    # ---- Unity messages -------------------------------------------------------------------------
    def Start(self) -> None:                                             # :45-48
        self.UpdateGeometry()

    def Update(self) -> None:                                            # :50-62 (keyboard handling omitted)
        self.frameCount += 1
        if self.source is None or self.listener is None:
            return
        self.RunSimulation()
        # Unity resumes `yield return null` coroutines once per frame after Update
        self._coroutines = [c for c in self._coroutines if self._advance(c)]

    @staticmethod
    def _advance(co) -> bool:
        try:
            next(co)
            return True
        except StopIteration:
            return False

    def FixedUpdate(self) -> None:                                       # :64-89
        am = self.audioManager
        if am is None or not am.IsStreaming:
            return
        if self.dynamicObstacles:
            self.UpdateGeometry()
        samplesThisFrame = _round_to_int(np.float32(self.fixedDeltaTime) * np.float32(self.sampleRate))
        self.samplesSinceLastChunk += samplesThisFrame
        if self.samplesSinceLastChunk >= self.chunkSamples:
            if self.nextStreamingOffset >= len(self.fullInputSamples):
                if self.loop:
                    self.nextStreamingOffset = 0
                else:
                    am.StopStreaming()
            if am.IsStreaming:
                co = self.ProcessChunk(self.nextStreamingOffset, self.chunkSamples, max(1, self.accumFrames),
                                       self.GetActiveIRBuffer())
                if self._advance(co):
                    self._coroutines.append(co)
                self.activeIRIndex = 1 - self.activeIRIndex
                self.nextStreamingOffset += self.chunkSamples
                self.ResetIR()
                self.samplesSinceLastChunk -= self.chunkSamples

    def ProcessChunk(self, sampleOffset: int, chunkLen: int, accumCount: int, ir: int):   # :91-123
        inputLen = min(chunkLen, len(self.fullInputSamples) - sampleOffset)
        if inputLen <= 0:
            return
        irLen = self._slot_len[ir]
        outputLen = inputLen + irLen
        chunk = self.fullInputSamples[sampleOffset: sampleOffset + inputLen]
        ticket = self._ctx.convolve_begin(ir, chunk, accumCount)         # SetData + Dispatch + AsyncGPUReadback.Request
        try:
            while not self._ctx.poll(ticket):                            # while (!req.done) yield return null
                yield None
        except _capi.RarError:                                           # if (req.hasError) yield break
            return
        result = self._ctx.convolve_end(ticket, outputLen)               # req.GetData<float>().CopyTo(result)
        self.audioManager.PushSamples(result, sampleOffset)

    def StartStreaming(self) -> None:                                    # :125-133
        self.nextStreamingOffset = 0
        self.samplesSinceLastChunk = 0
        self.chunkSamples = _round_to_int(np.float32(self.sampleRate) * np.float32(self.audioManager.chunkDuration))
        self.fullInputSamples = self.LoadSample(self.inputClip)
        self.ResetIR()
        self.audioManager.StartStreaming(self.reverbDuration)

    def LoadSample(self, clip: AudioClip) -> np.ndarray:                 # :135-167
        raw = clip.GetData()
        ch = clip.channels
        mono = np.zeros(clip.samples, dtype=np.float32)
        for c in range(ch):                                              # sequential channel sum (:144-146)
            mono = mono + raw[c::ch][: clip.samples]
        mono = (mono / np.float32(ch)).astype(np.float32)
        if clip.frequency == self.sampleRate:
            return mono
        ratio = np.float32(clip.frequency) / np.float32(self.sampleRate)
        newLength = _round_to_int(np.float32(clip.samples) / ratio)
        i = np.arange(newLength, dtype=np.float32)
        srcIdx = (i * ratio).astype(np.float32)
        idx0 = np.floor(srcIdx).astype(np.int64)
        idx1 = np.minimum(idx0 + 1, len(mono) - 1)
        t = np.clip(srcIdx - idx0.astype(np.float32), 0, 1).astype(np.float32)
        a, b = mono[idx0], mono[idx1]
        return (a + (b - a) * t).astype(np.float32)                      # Mathf.Lerp

    def LoadSamples(self, clips) -> list:
        """LoadSample for many sources at once on the GPU (rar_prepare_clips): clips of the same shape
        (samples, channels, frequency) are prepared by one launch per shape group; results are bit-identical
        to LoadSample and come back in the order given."""
        out = [None] * len(clips)
        groups = {}
        for i, c in enumerate(clips):
            groups.setdefault((c.samples, c.channels, c.frequency), []).append(i)
        for (samples, ch, freq), idx in groups.items():
            raw = np.concatenate([np.asarray(clips[i].GetData(), dtype=np.float32)[: samples * ch] for i in idx])
            res = self._ctx.prepare_clips(raw, samples, ch, freq, self.sampleRate, len(idx))
            for k, i in enumerate(idx):
                out[i] = res[k]
        return out

    def _ir_length(self) -> int:
        return int(np.float32(self.sampleRate) * np.float32(self.reverbDuration))   # (int)(sampleRate * reverbDuration)

    def ResetIR(self) -> None:                                           # :169-177
        self.accumFrames = 0
        length = self._ir_length()
        slot = self.GetActiveIRBuffer()
        ComputeHelper.CreateIRBuffer(self._ctx, slot, length)            # ClearImpulse
        self._slot_len[slot] = length

    def RunSimulation(self) -> None:                                     # :179-210 fused with :220-233
        irLength = self._ir_length()
        if self.activeSegments is None:
            self.UpdateGeometry()
        slot = self.GetActiveIRBuffer()
        p = _capi.make_trace_params(
            source=self.source.position, listener=self.listener.position, listener_radius=self.listenerRadius,
            speed_of_sound=self.speedOfSound, input_gain=self.inputGain, max_bounce_count=self.maxBounces,
            rng_state_offset=self.frameCount, ray_count=self.rayCount, debug_ray_count=self.debugRayCount,
            sample_rate=self.sampleRate, impulse_length=irLength,
            # large scenes: look walls up through the uniform grid (identical results)
            flags=_capi.RAR_FLAG_USE_GRID if len(self.activeSegments) >= self.gridThreshold else 0)
        self._ctx.trace(p, slot)                                         # Trace + ProcessHits
        self.accumFrames += 1                                            # OnSimulationFinished :233

    def GetActiveIRBuffer(self) -> int:                                  # :212-218
        length = self._ir_length()
        for s in (0, 1):
            if self._slot_len[s] != length:                              # CreateStructuredBuffer<float>(ref ..., len)
                ComputeHelper.CreateIRBuffer(self._ctx, s, length)
                self._slot_len[s] = length
        return 0 if self.activeIRIndex == 0 else 1

    def UpdateGeometry(self) -> None:                                    # :246-250
        self.activeSegments = SceneToData2D.GetSegmentsFromColliders(self.obstacleObjects)
        ComputeHelper.CreateStructuredBuffer(self._ctx, self.activeSegments)

    def GetDebugRayPaths(self) -> np.ndarray:                            # debugRayPaths (:39, :207)
        return self._ctx.get_debug_rays(max(100, self.debugRayCount) * (self.maxBounces + 1))

    def ReadActiveIR(self) -> np.ndarray:
        """The active float IR (un-normalised sum over accumFrames), for inspection and tests."""
        return self._ctx.ir_read(self.GetActiveIRBuffer(), self._ir_length())

    def OnDestroy(self) -> None:                                         # :281
        if self._owns_ctx:
            ComputeHelper.Release(self._ctx)
        self._ctx = None


class RayTraceManagerStreaming(RayTraceManager):
    """The same component over a convolver that keeps running (not in the reference).

    The reference convolves every chunk from scratch with the IR of that moment and overlap-adds the
    `inputLen + irLen` results (ProcessChunk, :91-123).  Here one partitioned streaming convolver
    (rar_conv_*) carries the input history: when a chunk is due its IR slot becomes the convolver's
    response -- with a one-block cross-fade from the previous one (rar_conv_update_ir_from_slot) -- and
    the chunk's samples go through in blocks of 256; samples that do not fill a block wait for the next
    chunk.  Output blocks are pushed to the AudioManager at consecutive offsets.
    """

    BLOCK = 256

    def StartStreaming(self) -> None:
        super().StartStreaming()
        old = getattr(self, "_conv", None)
        if old is not None:
            old.destroy()
        self._conv = _capi.Convolver(self._ctx, 1, self.BLOCK, max(1, self._ir_length()))
        self._carry = np.zeros(0, dtype=np.float32)
        self._out_pos = 0
        self._have_ir = False

    def ProcessChunk(self, sampleOffset: int, chunkLen: int, accumCount: int, ir: int):
        inputLen = min(chunkLen, len(self.fullInputSamples) - sampleOffset)
        if inputLen <= 0:
            return
        if self._have_ir:
            self._conv.update_ir_from_slot(0, ir, accumCount)
        else:
            self._conv.set_ir_from_slot(0, ir, accumCount)
            self._have_ir = True
        data = np.concatenate([self._carry, self.fullInputSamples[sampleOffset: sampleOffset + inputLen]])
        n_blocks = len(data) // self.BLOCK
        for b in range(n_blocks):
            y = self._conv.process(data[b * self.BLOCK:(b + 1) * self.BLOCK][None, :])[0]
            self.audioManager.PushSamples(y, self._out_pos)
            self._out_pos += self.BLOCK
        self._carry = data[n_blocks * self.BLOCK:]
        if False:  # a coroutine in the base class; nothing to wait for here
            yield None

    def DrainTail(self) -> None:
        """Pushes the remaining input and the reverberation tail (zero input for one IR length) through."""
        n = len(self._carry) + self._ir_length()
        data = np.concatenate([self._carry, np.zeros(n - len(self._carry) + (-n) % self.BLOCK, dtype=np.float32)])
        for b in range(len(data) // self.BLOCK):
            y = self._conv.process(data[b * self.BLOCK:(b + 1) * self.BLOCK][None, :])[0]
            self.audioManager.PushSamples(y, self._out_pos)
            self._out_pos += self.BLOCK
        self._carry = np.zeros(0, dtype=np.float32)

    def OnDestroy(self) -> None:
        conv = getattr(self, "_conv", None)
        if conv is not None:
            conv.destroy()
            self._conv = None
        super().OnDestroy()


class RayTraceManagerComplex(RayTraceManager):
    """The offline use-case of Assets/Script/RayTraceManagerComplex.cs: BakeAudio (:170-227) convolves the
    whole clip with the single accumulated IR in one call and PlayResult (:228-245) peak-normalises it.
    (The experimental banded deposit of that script is covered by the banded trace of the C-ABI, not here.)"""

    def BakeAudio(self) -> Optional[np.ndarray]:
        if self.inputClip is None:
            return None                                                  # Debug.LogError("Assign an Input Clip!")
        clip = self.inputClip
        raw = clip.GetData()
        mono = np.zeros(clip.samples, dtype=np.float32)
        for c in range(clip.channels):
            mono = mono + raw[c::clip.channels][: clip.samples]
        mono = (mono / np.float32(clip.channels)).astype(np.float32)
        slot = self.GetActiveIRBuffer()
        irLen = self._ir_length()
        result = self._ctx.convolve(slot, mono, max(1, self.accumFrames), irLen)   # synchronous GetData (:209)
        return self.PlayResult(result)

    @staticmethod
    def PlayResult(data: np.ndarray) -> np.ndarray:
        maxVol = float(np.max(np.abs(data))) if len(data) else 0.0
        if maxVol > 0.0001:
            data = (data * np.float32(1.0 / maxVol)).astype(np.float32)
        return data

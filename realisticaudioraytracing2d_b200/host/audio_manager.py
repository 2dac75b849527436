"""Host mirror of Assets/Script/AudioManager.cs: the playback sink of the streaming path.

A lock-protected float ring buffer: `PushSamples` overlap-adds convolved chunks at an absolute sample
offset (:45-54), `OnAudioFilterRead` (the audio thread in Unity) drains and zeroes it (:56-69).  It keeps
the reference's API so that the manager code and the tests read like the reference.

`AudioManager` is the literal mirror (a Python lock, numpy arrays).  `NativeAudioManager` has the same
interface over the library's lock-free single-producer / single-consumer ring (rar_ring_*, csrc/ring.cu):
what the C# AudioManager binds so that the audio callback never waits for the main thread's push.
"""
from __future__ import annotations

import math
import threading

import numpy as np


class AudioManager:
    def __init__(self, outputSampleRate: int = 48000, chunkDuration: float = 0.1):
        self.chunkDuration = chunkDuration          # [Range(0.05, 1.0)] (:5)
        self.sampleRate = outputSampleRate          # AudioSettings.outputSampleRate (:16)
        self.ringBuffer = None
        self.readHead = 0
        self.bufferSize = 0
        self.bufferLock = threading.Lock()
        self.isStreaming = False

    @property
    def IsStreaming(self) -> bool:                  # :12
        return self.isStreaming

    def StartStreaming(self, reverbDuration: float) -> None:   # :26-36
        if self.isStreaming:
            self.StopStreaming()
        self.bufferSize = int(math.ceil(float(np.float32(self.sampleRate) * (np.float32(reverbDuration) + np.float32(1.0)))))
        self.ringBuffer = np.zeros(self.bufferSize, dtype=np.float32)
        self.readHead = 0
        self.isStreaming = True

    def StopStreaming(self) -> None:                # :38-43
        if not self.isStreaming:
            return
        self.isStreaming = False

    def PushSamples(self, samples: np.ndarray, sampleOffset: int) -> None:   # :45-54
        if not self.isStreaming or self.ringBuffer is None:
            return
        samples = np.asarray(samples, dtype=np.float32)
        with self.bufferLock:
            writePos = sampleOffset % self.bufferSize
            idx = (writePos + np.arange(len(samples))) % self.bufferSize
            np.add.at(self.ringBuffer, idx, samples)   # += with wrap-around (a chunk may lap the ring)

    def OnAudioFilterRead(self, data: np.ndarray, channels: int) -> None:    # :56-69
        if not self.isStreaming or self.ringBuffer is None:
            return
        n = len(data) // channels
        with self.bufferLock:
            idx = (self.readHead + np.arange(n)) % self.bufferSize
            s = self.ringBuffer[idx].copy()
            self.ringBuffer[idx] = 0
            self.readHead = int((self.readHead + n) % self.bufferSize)
        data[: n * channels] = np.repeat(s, channels)

    def OnDestroy(self) -> None:                    # :71
        self.StopStreaming()


class NativeAudioManager(AudioManager):
    """AudioManager over the native lock-free ring (include/rar2d.h rar_ring_*)."""

    def __init__(self, outputSampleRate: int = 48000, chunkDuration: float = 0.1):
        super().__init__(outputSampleRate, chunkDuration)
        self._ring = None

    def StartStreaming(self, reverbDuration: float) -> None:
        from .. import _capi
        if self.isStreaming:
            self.StopStreaming()
        if self._ring is not None:
            self._ring.destroy()
        self._ring = _capi.Ring(self.sampleRate, reverbDuration)
        self.bufferSize = self._ring.size
        self.isStreaming = True

    def StopStreaming(self) -> None:
        if not self.isStreaming:
            return
        self.isStreaming = False
        if self._ring is not None:
            self._ring.stop()

    def PushSamples(self, samples: np.ndarray, sampleOffset: int) -> None:
        if not self.isStreaming or self._ring is None:
            return
        self._ring.push(samples, sampleOffset)

    def OnAudioFilterRead(self, data: np.ndarray, channels: int) -> None:
        if not self.isStreaming or self._ring is None:
            return
        self._ring.drain(data, channels)

    def OnDestroy(self) -> None:
        self.StopStreaming()
        if self._ring is not None:
            self._ring.destroy()
            self._ring = None

"""Recorded audio as a source waveform (reference core/waveforms.py:33-231).

Mirror of the reference's ``AudioFileWaveform`` for boxes where the reference package is absent: a WAV file (other formats
through ``soundfile`` when it is installed) becomes float32 samples in [-1, 1], one channel or the mean of all, trimmed to
``start_time`` / ``duration``; ``waveform(t, dt)`` -- the interface of ``GaussianPulse.waveform`` -- looks the times up in
the recording resampled to 1/dt (FFT resampling, cached per rate), zero past the end unless ``loop`` is set.  The solver
evaluates it once per chunk of steps into the device's waveform table like any other source.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path

import numpy as np


def _to_unit_float(data: np.ndarray) -> np.ndarray:
    """Integer PCM to [-1, 1) float32 (8-bit WAV is unsigned), floats unchanged."""
    if data.dtype == np.int16:
        return data.astype(np.float32) / 32768.0
    if data.dtype == np.int32:
        return data.astype(np.float32) / 2147483648.0
    if data.dtype == np.uint8:
        return (data.astype(np.float32) - 128) / 128.0
    return data.astype(np.float32)


@dataclass
class AudioFileWaveform:
    filepath: str | Path
    amplitude: float = 1.0
    channel: int | str = "mix"            # index of a channel, or "mix" = mean of all
    start_time: float = 0.0
    duration: float | None = None
    loop: bool = False
    _samples: np.ndarray | None = field(default=None, init=False, repr=False)
    _native_sr: int | None = field(default=None, init=False, repr=False)
    _resampled: dict = field(default_factory=dict, init=False, repr=False)

    def __post_init__(self) -> None:
        self._load_audio()

    def _load_audio(self) -> None:
        path = Path(self.filepath)
        if not path.exists():
            raise FileNotFoundError(f"Audio file not found: {path}")
        if path.suffix.lower() == ".wav":
            from scipy.io import wavfile
            rate, data = wavfile.read(path)
        else:
            try:
                import soundfile
            except ImportError as e:
                raise ImportError(f"soundfile package required for {path.suffix} files. Install with: pip install soundfile") from e
            data, rate = soundfile.read(path)
        self._native_sr = rate
        data = _to_unit_float(data)
        if data.ndim > 1:
            if self.channel == "mix":
                data = np.mean(data, axis=1)
            elif self.channel >= data.shape[1]:
                raise ValueError(f"Channel {self.channel} requested but file only has {data.shape[1]} channels")
            else:
                data = data[:, self.channel]
        first = int(self.start_time * rate)
        if first >= len(data):
            raise ValueError(f"start_time {self.start_time}s is beyond end of file ({len(data) / rate:.2f}s)")
        last = None if self.duration is None else first + int(self.duration * rate)
        self._samples = data[first:last].astype(np.float32)

    def _resample_for_dt(self, dt: float) -> np.ndarray:
        rate = int(round(1.0 / dt))
        if rate not in self._resampled:
            if rate == self._native_sr:
                self._resampled[rate] = self._samples
            else:
                from scipy import signal
                self._resampled[rate] = signal.resample(self._samples, int(len(self._samples) * rate / self._native_sr)).astype(np.float32)
        return self._resampled[rate]

    def waveform(self, t, dt: float):
        track = self._resample_for_dt(dt)
        at = (t / dt).astype(np.int64)
        if self.loop:
            return self.amplitude * track[at % len(track)]
        out = np.zeros_like(t, dtype=np.float32)
        ok = (at >= 0) & (at < len(track))
        out[ok] = track[at[ok]]
        return self.amplitude * out

    @property
    def duration_seconds(self) -> float:
        return 0.0 if self._samples is None or self._native_sr is None else len(self._samples) / self._native_sr

    @property
    def native_sample_rate(self) -> int:
        return self._native_sr or 0

    @property
    def num_samples(self) -> int:
        return 0 if self._samples is None else len(self._samples)

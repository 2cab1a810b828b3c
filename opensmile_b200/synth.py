"""Deterministic synthetic PCM: a voiced-like harmonic source plus noise.

x(t) = 0.1 * sum_{k=1..19} sin(k*phi(t))/k + 0.02*N(0,1),  f0(t) = 120 + 30 sin(2 pi 0.5 t),
clipped to [-1, 1], scaled by 32767 -> int16.  numpy `default_rng(seed)` drives the noise.
"""
import numpy as np


def voiced_pcm(n_samples, sample_rate=16000, seed=0, n_chan=1):
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    f0 = 120.0 + 30.0 * np.sin(2 * np.pi * 0.5 * t + 0.37 * seed)
    phi = 2 * np.pi * np.cumsum(f0) / sample_rate
    x = np.zeros(n_samples)
    for k in range(1, 20):
        x += np.sin(k * phi) / k
    x *= 0.1
    chans = []
    for _ in range(n_chan):
        y = x + 0.02 * rng.standard_normal(n_samples)
        chans.append(np.clip(y, -1.0, 1.0))
    y = np.stack(chans, axis=1).reshape(-1) if n_chan > 1 else chans[0]
    return np.round(y * 32767.0).astype(np.int16)


def mixed_pcm(n_samples, sample_rate=16000, seed=0):
    """voiced_pcm with stretches replaced by loud noise (unvoiced) and near silence: exercises the
    voiced / unvoiced decisions of the pitch chain (segments of 50..375 ms, kinds cycle voiced, noise,
    voiced, silence)."""
    rng = np.random.default_rng(seed)
    x = voiced_pcm(n_samples, sample_rate, seed=seed).astype(np.float64)
    seg = 0
    pos = 0
    while pos < n_samples:
        ln = int(rng.integers(sample_rate // 20, sample_rate * 3 // 8))
        kind = seg % 4
        m = min(ln, n_samples - pos)
        if kind == 1:
            x[pos:pos + m] = rng.normal(0, 600, size=m)
        if kind == 3:
            x[pos:pos + m] = rng.normal(0, 3, size=m)
        pos += ln
        seg += 1
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def stereo_mixed_pcm(n_samples, sample_rate=16000, seed=0):
    """interleaved stereo: left = mixed_pcm, right = the same signal 37 samples earlier, scaled by 0.8, plus
    its own noise (the mono mixdown therefore differs from either channel)"""
    a = mixed_pcm(n_samples + 64, sample_rate, seed=seed).astype(np.float64)
    rng = np.random.default_rng(seed + 100)
    left = a[64:]
    right = 0.8 * a[27:27 + n_samples] + rng.normal(0, 40, size=n_samples)
    x = np.stack([left, right], axis=1)
    return np.clip(np.round(x), -32768, 32767).astype(np.int16).reshape(-1)

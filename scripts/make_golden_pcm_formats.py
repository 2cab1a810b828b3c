"""Generate tests/golden/pcm_formats.npz with the UNMODIFIED reference (oracle/_ref/SMILExtract): for WAV files in every sample
format smilePcm_convertSamples / smilePcm_convertFloatSamples accept (8 / 24 / 32 bit integer, 24 valid bits in a 32-bit container,
32-bit float; mono and stereo) the reference's `wave` level (an extra cHtkSink on it: exact float32 samples after conversion and
mixdown) and its MFCC12_0_D_A rows.  The file's data chunk is stored as bytes.  Build container only."""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402
from opensmile_b200.synth import voiced_pcm  # noqa: E402

REF = "/root/reference/config"
N = 4000
# name -> (osm_b200_pcm_format, WAV format tag, bits per sample, bytes per sample, channels)
VARIANTS = {"s8_mono": (2, 1, 8, 1, 1), "s8_stereo": (2, 1, 8, 1, 2), "s24_mono": (3, 1, 24, 3, 1), "s24_stereo": (3, 1, 24, 3, 2),
            "s24in32_mono": (4, 1, 24, 4, 1), "s32_mono": (5, 1, 32, 4, 1), "s32_stereo": (5, 1, 32, 4, 2),
            "f32_mono": (1, 3, 32, 4, 1), "f32_stereo": (1, 3, 32, 4, 2)}


def encode(x, fmt):
    """x: float64 [frames, chan] in [-1, 1] -> bytes of interleaved little-endian samples"""
    if fmt == 2:
        return np.round(x * 127).astype(np.int8).tobytes()
    if fmt == 3:
        v = np.round(x * 8388607).astype(np.int32).reshape(-1)
        return np.stack([(v & 0xFF), (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=1).astype(np.uint8).tobytes()
    if fmt == 4:
        return np.round(x * 8388607).astype("<i4").tobytes()        # negative values keep their sign bits above bit 23
    if fmt == 5:
        return np.round(x * 2147483000).astype("<i4").tobytes()
    return x.astype("<f4").tobytes()


def write_wav_raw(path, data, tag, bits, bps, nchan, sr):
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, tag, nchan, sr, sr * bps * nchan, bps * nchan, bits))
        f.write(b"data" + struct.pack("<I", len(data)) + data)


def main():
    base = voiced_pcm(N, 16000, seed=21).astype(np.float64) / 32767.0
    other = voiced_pcm(N, 16000, seed=22).astype(np.float64) / 32767.0
    out = {}
    for name, (fmt, tag, bits, bps, nchan) in VARIANTS.items():
        x = base[:, None] if nchan == 1 else np.stack([base, 0.7 * other], axis=1)
        data = encode(x, fmt)
        with tempfile.TemporaryDirectory() as d:
            wav = os.path.join(d, "in.wav")
            write_wav_raw(wav, data, tag, bits, bps, nchan, 16000)
            conf = os.path.join(d, "w.conf")
            txt = open(os.path.join(REF, "mfcc", "MFCC12_0_D_A.conf")).read().replace("\\{../shared/", "\\{" + REF + "/shared/")
            txt += "\n[componentInstances:cComponentManager]\ninstance[dbgwave].type=cHtkSink\n[dbgwave:cHtkSink]\nreader.dmLevel=wave\nfilename=%s\nparmKind=9\n" % os.path.join(d, "wave.htk")
            open(conf, "w").write(txt)
            r = subprocess.run([refrun.SMILEXTRACT, "-C", conf, "-I", wav, "-O", os.path.join(d, "o.htk"), "-l", "1"], capture_output=True, text=True)
            if r.returncode:
                print(r.stderr[-3000:])
                sys.exit(1)
            w, _ = refrun.read_htk(os.path.join(d, "wave.htk"))
            m, _ = refrun.read_htk(os.path.join(d, "o.htk"))
        out["data_" + name] = np.frombuffer(data, np.uint8)
        out["wave_" + name] = w.reshape(-1).astype(np.float32)
        out["mfcc_" + name] = m.astype(np.float32)
        print(name, len(data), w.shape, m.shape, float(np.abs(w).max()))
    out["variants"] = np.array(list(VARIANTS))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pcm_formats.npz"), **out)


if __name__ == "__main__":
    main()

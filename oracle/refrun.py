"""Run the UNMODIFIED reference (oracle/_ref/SMILExtract, built by `make -C oracle ref`).

TEST INFRASTRUCTURE ONLY.  Used to pin the C restatement, to generate tests/golden/ and as the
CPU baseline of bench.py (`cpu_baseline.kind == "reference"`, `--impl reference`).
"""
import os
import struct
import subprocess
import tempfile
import wave

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
SMILEXTRACT = os.path.join(REF_DIR, "SMILExtract")
CONFIG_DIR = os.path.join(REF_DIR, "config")


def available():
    return os.access(SMILEXTRACT, os.X_OK) and os.path.isdir(CONFIG_DIR)


def write_wav(path, pcm, sample_rate, n_chan=1):
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(n_chan)
        w.setsampwidth(2)
        w.setframerate(int(sample_rate))
        w.writeframes(pcm.tobytes())


def read_wav(path):
    with wave.open(path, "rb") as w:
        assert w.getsampwidth() == 2
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
        return pcm, w.getframerate(), w.getnchannels()


def read_htk(path):
    """HTK parameter file (iocore/htkSink.cpp:53,90-106): 12-byte big-endian header + BE f32."""
    with open(path, "rb") as f:
        hdr = f.read(12)
        n, period, size, kind = struct.unpack(">iihh", hdr)
        data = np.frombuffer(f.read(), dtype=">f4").astype(np.float32)
    return data.reshape(n, size // 4), dict(n=n, period=period, size=size, kind=kind)


def run_config(conf_rel, wav_path, out_path, extra=()):
    cmd = [SMILEXTRACT, "-C", os.path.join(CONFIG_DIR, conf_rel), "-I", wav_path,
           "-O", out_path, "-l", "0", "-nologfile"] + list(extra)
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def extract(conf_rel, pcm, sample_rate, n_chan=1, conf_text=None):
    """pcm int16 -> [T, n_out] float32 through the reference binary (HTK sink)."""
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        wav = os.path.join(d, "in.wav")
        out = os.path.join(d, "out.htk")
        write_wav(wav, pcm, sample_rate, n_chan)
        if conf_text is not None:
            conf = os.path.join(CONFIG_DIR, conf_rel)
            with open(conf, "w") as f:
                f.write(conf_text)
        run_config(conf_rel, wav, out)
        data, hdr = read_htk(out)
    return data

"""The value formatting of the file sinks on the device (opensmile_b200/csrc/sinks.cu): cCsvSink / cHtkSink files
written from rows resident in HBM are byte-identical to the host writers (which are pinned byte for byte against the reference's
files on the CPU), on LLD-like rows and on adversarial values (integer valued, dyadic ties of the 7-digit rounding, denormals,
non-finite values that take the host path); the session's file extraction uses them."""
import os

import numpy as np
import pytest

from opensmile_b200.synth import mixed_pcm, voiced_pcm

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _rows(seed, n, k):
    rng = np.random.default_rng(seed)
    r = (rng.standard_normal((n, k)) * 10.0 ** rng.integers(-9, 6, size=(1, k))).astype(np.float32)
    r[:, 0] = np.round(r[:, 0])                                  # integer valued -> "%.0f"
    if k > 3:
        r[:, 1] = np.ldexp(rng.integers(1, 4096, n) * 2 + 1, rng.integers(-40, 10, n)).astype(np.float32)   # dyadic: exact ties
        r[::7, 2] = 0.0
        r[1::7, 2] = -0.0
        r[:, 3] = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)            # any bit pattern (nan / inf / denormal)
    if n > 6 and k > 4:
        r[5, 4] = 1e16
        r[6, 4] = 9.9999995e6
    return r


@pytest.mark.parametrize("n,k", [(1, 1), (257, 39), (1000, 130), (3, 200)])
def test_device_writers_equal_the_host_writers(tmp_path, n, k):
    import torch
    from opensmile_b200 import session as S
    rows = _rows(n * 131 + k, n, k)
    d = torch.from_numpy(rows).cuda()
    names = ["c%d" % i for i in range(k)]
    S.write_csv(tmp_path / "h.csv", rows, names, 0.01, instance_name="x", n_time_frames=max(n - 2, 0))
    S.write_csv_device(tmp_path / "d.csv", d, n, names, 0.01, instance_name="x", n_time_frames=max(n - 2, 0))
    assert (tmp_path / "h.csv").read_bytes() == (tmp_path / "d.csv").read_bytes()
    S.write_htk(tmp_path / "h.htk", rows, 0.01)
    S.write_htk_device(tmp_path / "d.htk", d, n, k, 0.01)
    assert (tmp_path / "h.htk").read_bytes() == (tmp_path / "d.htk").read_bytes()


def test_session_files_come_from_the_device_sinks(tmp_path):
    """extract_files (device sinks: HTK payload, CSV and ARFF value text) against write_files on the rows of extract_pcm (host
    formatting): identical files"""
    import wave
    from opensmile_b200.session import Session
    conf = os.path.join(HERE, "configs", "mfcc_e_d_a.conf")
    pcms = [mixed_pcm(24000, 16000, seed=3), voiced_pcm(16000, 16000, seed=9)]
    wavs = []
    for i, x in enumerate(pcms):
        w = tmp_path / ("in%d.wav" % i)
        with wave.open(str(w), "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(np.ascontiguousarray(x, dtype="<i2").tobytes())
        wavs.append(str(w))
    s = Session(conf, options={"instname": "utt7"}, device=0)
    frames = s.extract_files(wavs, [str(tmp_path / ("d%d.htk" % i)) for i in range(2)], [str(tmp_path / ("d%d.csv" % i)) for i in range(2)],
                             [str(tmp_path / ("d%d.arff" % i)) for i in range(2)])
    off = np.concatenate([[0], np.cumsum([len(x) for x in pcms])]).astype(np.int64)
    rows, fo = s.extract_pcm(np.concatenate(pcms), off, 16000.0, 1)
    assert list(frames) == list(np.diff(fo))
    s.write_files(rows, fo, 16000.0, 1, n_samples=[len(x) for x in pcms], htk_paths=[str(tmp_path / ("h%d.htk" % i)) for i in range(2)],
                  csv_paths=[str(tmp_path / ("h%d.csv" % i)) for i in range(2)], arff_paths=[str(tmp_path / ("h%d.arff" % i)) for i in range(2)])
    s.close()
    for i in range(2):
        assert (tmp_path / ("d%d.htk" % i)).read_bytes() == (tmp_path / ("h%d.htk" % i)).read_bytes()
        assert (tmp_path / ("d%d.csv" % i)).read_bytes() == (tmp_path / ("h%d.csv" % i)).read_bytes()
        assert (tmp_path / ("d%d.arff" % i)).read_bytes() == (tmp_path / ("h%d.arff" % i)).read_bytes()      # every value "%e" (cArffSink)

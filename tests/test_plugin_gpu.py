"""GPU box: the REFERENCE's SMILExtract (dynamic build of the unmodified sources) with the B200 plugin in ./plugins/ writes the
reference's own sink files from rows computed by the CUDA plan.  Wave source, data memory, sinks, tick loop and end-of-input
handling are the reference's; only the chain between the wave level and the lld level is replaced (plugin/lldBlockB200.cpp)."""
import os
import subprocess
import sys
import wave

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402  (HTK reader only)

PLUG = os.path.join(ROOT, "plugin")
SMILE = os.path.join(ROOT, "oracle", "_ref_dyn", "SMILExtract")
REFCONF = os.path.join(ROOT, "oracle", "_ref", "config")
GOLD = os.path.join(HERE, "golden")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.access(SMILE, os.X_OK) and os.path.exists(os.path.join(PLUG, "plugins", "libosm_b200_plugin.so"))),
                                 reason="dynamic reference build / plugin not built")]


def _wav(path, pcm, sr, nchan=1):
    with wave.open(path, "wb") as w:
        w.setnchannels(nchan); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes(np.ascontiguousarray(pcm, "<i2").tobytes())


def _run(args):
    r = subprocess.run([SMILE] + args, cwd=PLUG, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "(ERR)" not in r.stdout, r.stdout
    return r.stdout


def _percol(got, ref):
    return float((np.abs(got - ref) / (np.abs(ref).max(axis=0) + 1e-30)).max())


def test_reference_smilextract_with_plugin_writes_the_golden_htk(tmp_path):
    g = np.load(os.path.join(GOLD, "mfcc_example_44k1.npz"))           # example-audio/opensmile.wav, config[0]
    wav, htk, csv = str(tmp_path / "in.wav"), str(tmp_path / "o.htk"), str(tmp_path / "o.csv")
    _wav(wav, g["pcm"], int(g["sample_rate"]))
    _run(["-C", "config/MFCC12_0_D_A_b200.conf", "-graphconf", os.path.join(REFCONF, "mfcc", "MFCC12_0_D_A.conf"),
          "-sinkconf", os.path.join(REFCONF, "shared", "standard_data_output_lldonly.conf.inc"),
          "-I", wav, "-O", htk, "-csvoutput", csv, "-l", "1"])
    rows, hdr = refrun.read_htk(htk)
    assert rows.shape == (202, 39) and hdr["period"] == 100000 and hdr["size"] == 156 and hdr["kind"] == 9
    assert os.path.getsize(htk) == 31524                                # the reference's file size
    assert _percol(rows, g["lld"]) < 1e-5
    head = open(csv).readline().strip().split(";")
    assert head[:2] == ["name", "frameTime"] and head[2] == "pcm_fftMag_mfcc[0]" and head[-1] == "pcm_fftMag_mfcc_de_de[12]"
    body = np.loadtxt(csv, delimiter=";", skiprows=1, usecols=range(1, 41))
    assert np.allclose(body[:, 0], np.arange(202) * 0.01)
    assert np.abs(body[:, 1:] - g["lld"]).max() <= 1e-5 * np.abs(g["lld"]).max() + 1e-6 * np.abs(g["lld"]).max()


def test_plugin_stereo_mixdown_is_recovered_exactly(tmp_path):
    """a stereo file: the reference's wave source mixes down to mono floats; the plugin re-encodes them as a two-channel
    int16 carrier with the same sums, so the rows equal a direct plan run on the file's own samples bit for bit"""
    from opensmile_b200 import Plan, components_mfcc12_0_d_a
    from opensmile_b200.synth import voiced_pcm
    pcm = voiced_pcm(30000, 16000, seed=4, n_chan=2)
    wav, htk = str(tmp_path / "in.wav"), str(tmp_path / "o.htk")
    _wav(wav, pcm, 16000, 2)
    _run(["-C", "config/MFCC12_0_D_A_b200.conf", "-graphconf", os.path.join(REFCONF, "mfcc", "MFCC12_0_D_A.conf"),
          "-sinkconf", os.path.join(REFCONF, "shared", "standard_data_output_lldonly.conf.inc"), "-I", wav, "-O", htk, "-l", "1"])
    rows, _ = refrun.read_htk(htk)
    p = Plan(components_mfcc12_0_d_a(16000.0, n_channels=2), "lld", device=0)
    direct = p.run_host(pcm, np.array([0, 30000], np.int64))
    p.close()
    assert rows.shape == direct.shape and np.array_equal(rows, direct)

"""The reference-side binding, compiled for real: plugin/plugins/libosm_b200_plugin.so (cLldBlockB200, built against the
reference's own headers) loaded by the reference's own plugin loader in the dynamic build of the UNMODIFIED reference
(oracle/_ref_dyn/SMILExtract, `make -C oracle refdyn`; plugin loader: src/core/componentManager.cpp:212-425).

CPU box: the plugin loads, registers its component type, a configuration using it validates (unknown fields are rejected by
the reference's own config manager), and without a CUDA device the run fails loudly -- there is no CPU path.
The GPU half is tests/test_plugin_gpu.py."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PLUG = os.path.join(ROOT, "plugin")
SMILE = os.path.join(ROOT, "oracle", "_ref_dyn", "SMILExtract")
REFCONF = os.path.join(ROOT, "oracle", "_ref", "config")

pytestmark = pytest.mark.skipif(not (os.access(SMILE, os.X_OK) and os.path.exists(os.path.join(PLUG, "plugins", "libosm_b200_plugin.so"))),
                                reason="dynamic reference build / plugin not built (make -C oracle refdyn && make -C plugin; needs /root/reference)")


def run(args, **kw):
    return subprocess.run([SMILE] + args, cwd=PLUG, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120, **kw)


def test_plugin_exports_the_loader_entry_point():
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(PLUG, "plugins", "libosm_b200_plugin.so")],
                         stdout=subprocess.PIPE, text=True).stdout
    assert " T registerPluginComponent" in out


def test_reference_lists_the_plugin_component():
    r = run(["-L"])
    assert "found plugin : 'libosm_b200_plugin.so'" in r.stdout
    assert "+++ 'cLldBlockB200' +++" in r.stdout
    assert "cannot open plugin" not in r.stdout


def test_reference_config_manager_knows_the_fields():
    r = run(["-H", "cLldBlockB200"])
    for f in ("graphConf", "captureTo", "graphOption", "device", "reader", "writer"):
        assert f in r.stdout, r.stdout


def test_unknown_field_is_rejected_by_the_reference(tmp_path):
    conf = open(os.path.join(PLUG, "config", "MFCC12_0_D_A_b200.conf")).read().replace("captureTo = lld", "captureTo = lld\nbogusField = 3")
    p = tmp_path / "bad.conf"
    p.write_text(conf)
    r = run(["-C", str(p), "-I", "x.wav", "-sinkconf", os.path.join(REFCONF, "shared", "standard_data_output_lldonly.conf.inc")])
    assert "bogusField" in r.stdout and "ERR" in r.stdout


def test_fails_loudly_without_cuda(tmp_path):
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "opensmile_b200", "libosm_b200.so"))
    if lib.osm_b200_device_count() > 0:
        pytest.skip("a CUDA device is present: covered by tests/test_plugin_gpu.py")
    import numpy as np
    import wave
    wav = str(tmp_path / "in.wav")
    with wave.open(wav, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes((np.arange(8000) % 200).astype("<i2").tobytes())
    out = str(tmp_path / "o.htk")
    r = run(["-C", "config/MFCC12_0_D_A_b200.conf", "-graphconf", os.path.join(REFCONF, "mfcc", "MFCC12_0_D_A.conf"),
             "-sinkconf", os.path.join(REFCONF, "shared", "standard_data_output_lldonly.conf.inc"), "-I", wav, "-O", out])
    assert "no usable CUDA device" in r.stdout and "no CPU fallback" in r.stdout
    assert not os.path.exists(out) or os.path.getsize(out) <= 12        # nothing was computed

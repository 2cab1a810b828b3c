"""Generate tests/golden/ from the UNMODIFIED reference (oracle/_ref/SMILExtract).

Run in the build container (needs /root/reference for `make -C oracle ref`):
    python scripts/make_golden.py
Fixtures (all small):
  mfcc_example_44k1.npz   config[0]: config/mfcc/MFCC12_0_D_A.conf on example-audio/opensmile.wav
                          (pcm int16 44.1 kHz mono, 90112 samples; lld float32 [202, 39])
  mfcc_synth16k_s0.npz    MFCC12_0_D_A on the seeded synthetic 16 kHz signal (seed 0, 80000
                          samples -> [498, 39]); only the OUTPUT is stored, the input is
                          regenerated from opensmile_b200.synth.voiced_pcm(80000, 16000, seed=0)
  plp_goldens.npz         config/plp/PLP_0_D_A.conf on the example wav ([202, 18]) and on a seeded
                          44.1 kHz stereo signal (voiced_pcm(44100, 44100, seed=2, n_chan=2) -> [98, 18])
  mfcc_taps16k_s1.npz     intermediate levels (fftmag, melspec, ft0) of the first 20 frames for
                          seed 1, 16 kHz, dumped with extra cHtkSink instances
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402
from opensmile_b200.synth import voiced_pcm  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

TAP_CONF = r"""
\{mfcc/MFCC12_0_D_A.conf}
[componentInstances:cComponentManager]
instance[tapmag].type=cHtkSink
instance[tapmel].type=cHtkSink
instance[tapft0].type=cHtkSink
[tapmag:cHtkSink]
reader.dmLevel = fftmag
filename = \cm[tapmag{tapmag.htk}:fftmag tap]
parmKind = 9
[tapmel:cHtkSink]
reader.dmLevel = melspec
filename = \cm[tapmel{tapmel.htk}:melspec tap]
parmKind = 9
[tapft0:cHtkSink]
reader.dmLevel = ft0
filename = \cm[tapft0{tapft0.htk}:ft0 tap]
parmKind = 9
"""


def main():
    assert refrun.available(), "build the reference first: make -C oracle ref"
    os.makedirs(GOLD, exist_ok=True)
    pcm, sr, nch = refrun.read_wav("/root/reference/example-audio/opensmile.wav")
    lld = refrun.extract("mfcc/MFCC12_0_D_A.conf", pcm, sr, nch)
    np.savez_compressed(os.path.join(GOLD, "mfcc_example_44k1.npz"), pcm=pcm, sample_rate=sr, lld=lld)
    print("example", lld.shape)

    pcm = voiced_pcm(80000, 16000, seed=0)
    lld = refrun.extract("mfcc/MFCC12_0_D_A.conf", pcm, 16000)
    np.savez_compressed(os.path.join(GOLD, "mfcc_synth16k_s0.npz"), lld=lld, crc=np.int64(pcm.astype(np.int64).sum()))
    print("synth16k", lld.shape)

    # PLP_0_D_A (config/plp/PLP_0_D_A.conf): example wav + 44.1 kHz STEREO synthetic (BASELINE cfg 5 shape)
    pcm, sr, nch = refrun.read_wav("/root/reference/example-audio/opensmile.wav")
    lld = refrun.extract("plp/PLP_0_D_A.conf", pcm, sr, nch)
    pcm2 = voiced_pcm(44100, 44100, seed=2, n_chan=2)
    lld2 = refrun.extract("plp/PLP_0_D_A.conf", pcm2, 44100, 2)
    np.savez_compressed(os.path.join(GOLD, "plp_goldens.npz"), example_lld=lld, stereo44k1_lld=lld2,
                        stereo_crc=np.int64(pcm2.astype(np.int64).sum()))
    print("plp", lld.shape, lld2.shape)

    # intermediate taps through extra sinks (first 20 frames kept)
    import tempfile
    pcm = voiced_pcm(16000, 16000, seed=1)
    with tempfile.TemporaryDirectory() as d:
        conf = os.path.join(refrun.CONFIG_DIR, "_taps_mfcc.conf")
        with open(conf, "w") as f:
            f.write(TAP_CONF)
        wav = os.path.join(d, "in.wav")
        refrun.write_wav(wav, pcm, 16000)
        refrun.run_config("_taps_mfcc.conf", wav, os.path.join(d, "out.htk"),
                          ["-tapmag", os.path.join(d, "mag.htk"), "-tapmel", os.path.join(d, "mel.htk"),
                           "-tapft0", os.path.join(d, "ft0.htk")])
        mag = refrun.read_htk(os.path.join(d, "mag.htk"))[0]
        mel = refrun.read_htk(os.path.join(d, "mel.htk"))[0]
        ft0 = refrun.read_htk(os.path.join(d, "ft0.htk"))[0]
        os.remove(conf)
    np.savez_compressed(os.path.join(GOLD, "mfcc_taps16k_s1.npz"), fftmag=mag[:20], melspec=mel[:20], ft0=ft0[:20],
                        n_frames=np.int64(mag.shape[0]))
    print("taps", mag.shape, mel.shape, ft0.shape)


if __name__ == "__main__":
    main()

"""Generate tests/golden/conf_goldens.npz: outputs of the UNMODIFIED reference (oracle/_ref/SMILExtract)
for whole configuration files, used by the session (conf front end) tests.

    python scripts/make_golden_conf.py        # needs `make -C oracle ref` (build container only)

Contents (inputs are regenerated in the tests from opensmile_b200.synth.voiced_pcm with the seeds below):
  mix16k            tests/configs/lld_mix.conf, voiced_pcm(16000, 16000, seed=3)           [97, 52]
  mix32k_stereo     tests/configs/lld_mix.conf, voiced_pcm(16000, 32000, seed=4, n_chan=2)   (FFT 1024 + 2048)
  mfcc_e            config/mfcc/MFCC12_E_D_A.conf, voiced_pcm(12000, 16000, seed=5)
  mfcc_e_short_<n>  the same config on the first n samples of seed 5, n = 400, 560, 720, 880 (1..4 frames)
  plp_e             config/plp/PLP_E_D_A.conf, voiced_pcm(12000, 16000, seed=6)
  cmp_ns            tests/configs/compare_ns.conf (ComParE_2016's LLD-path columns, 59 + 59 deltas),
                    voiced_pcm(16000, 16000, seed=7); cmp_ns_short_<n>: its first n samples, n = 960, 1100, 1300, 2000;
                    cmp_ns_44k: voiced_pcm(30000, 44100, seed=8)
  gemaps_ns         tests/configs/gemaps_ns.conf (eGeMAPSv02's LLD-path columns), voiced_pcm(16000, 16000, seed=10)
  mfcc_z            tests/configs/mfcc_0_d_a_z.conf (cFullinputMean on the statics), voiced_pcm(12000, 16000, seed=11);
                    mfcc_z_plain: tests/configs/mfcc_0_d_a.conf on the same input (statics before the mean subtraction);
                    ref_mfcc_0_z / ref_mfcc_e_z / ref_plp_0_z / ref_plp_e_z: the reference's shipped *_Z configurations, same input
  ref_prosody_acf   config/prosody/prosodyAcf.conf (cPitchACF + cIntensity loudness, sma3) on seed 11
  ref_audspec / ref_audspec_compat / ref_spectrogram / ref_demo1_energy_csv: the reference's config/audspec/*.conf,
                    config/spectrum/spectrogram.conf (first 4000 samples) and config/demo/demo1_energy.conf (its CSV file) on seed 11
  cmp_taps          static levels audR (26) | audSum | audRSum of compare_ns.conf for the same input (oracle pin)
  rasta_plp         tests/configs/rasta_plp.conf (RASTA-PLP cepstra 0..8 + delta), voiced_pcm(16000, 16000, seed=9)
  names_<case>      the CSV header's element names
  htk_bytes / csv_bytes   the reference's HTK and CSV files for case mfcc_e (instance name 'utt7')
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402
from opensmile_b200.synth import voiced_pcm  # noqa: E402


def run(conf, pcm, sr, nch, keep_files=False, csv_out=True):
    with tempfile.TemporaryDirectory() as d:
        wav, htk, csv = (os.path.join(d, x) for x in ("in.wav", "out.htk", "out.csv"))
        refrun.write_wav(wav, pcm, sr, nch)
        extra = ["-csvoutput", csv, "-instname", "utt7"] if csv_out else []
        subprocess.run([refrun.SMILEXTRACT, "-C", conf, "-I", wav, "-O", htk, "-l", "0"] + extra,
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        rows, _ = refrun.read_htk(htk)
        hdr = open(csv).readline().strip().split(";") if csv_out else []
        names = [h for h in hdr if h not in ("name", "frameIndex", "frameTime")]
        files = (open(htk, "rb").read(), open(csv, "rb").read()) if keep_files else None
    return rows, np.array(names), files


def main():
    assert refrun.available(), "build the reference first: make -C oracle ref"
    out = {}
    mix = os.path.join(ROOT, "tests", "configs", "lld_mix.conf")
    out["mix16k"], out["names_mix"], _ = run(mix, voiced_pcm(16000, 16000, seed=3), 16000, 1)
    out["mix32k_stereo"], _, _ = run(mix, voiced_pcm(16000, 32000, seed=4, n_chan=2), 32000, 2)
    mfe = os.path.join(refrun.CONFIG_DIR, "mfcc", "MFCC12_E_D_A.conf")
    pcm = voiced_pcm(12000, 16000, seed=5)
    out["mfcc_e"], out["names_mfcc_e"], files = run(mfe, pcm, 16000, 1, keep_files=True)
    out["htk_bytes"] = np.frombuffer(files[0], dtype=np.uint8)
    out["csv_bytes"] = np.frombuffer(files[1], dtype=np.uint8)
    for n in (400, 560, 720, 880):
        out["mfcc_e_short_%d" % n], _, _ = run(mfe, pcm[:n], 16000, 1)
    out["plp_e"], out["names_plp_e"], _ = run(os.path.join(refrun.CONFIG_DIR, "plp", "PLP_E_D_A.conf"),
                                              voiced_pcm(12000, 16000, seed=6), 16000, 1)
    cns = os.path.join(ROOT, "tests", "configs", "compare_ns.conf")
    pcm = voiced_pcm(16000, 16000, seed=7)
    out["cmp_ns"], out["names_cmp_ns"], _ = run(cns, pcm, 16000, 1)
    for n in (960, 1100, 1300, 2000):
        out["cmp_ns_short_%d" % n], _, _ = run(cns, pcm[:n], 16000, 1)
    out["cmp_ns_44k"], _, _ = run(cns, voiced_pcm(30000, 44100, seed=8), 44100, 1)
    out["rasta_plp"], _, _ = run(os.path.join(ROOT, "tests", "configs", "rasta_plp.conf"), voiced_pcm(16000, 16000, seed=9), 16000, 1,
                                 csv_out=False)
    gns = os.path.join(ROOT, "tests", "configs", "gemaps_ns.conf")
    out["gemaps_ns"], out["names_gemaps_ns"], _ = run(gns, voiced_pcm(16000, 16000, seed=10), 16000, 1)
    # cepstral mean subtraction (cFullinputMean): own config + the reference's shipped *_Z configurations
    pz = voiced_pcm(12000, 16000, seed=11)
    out["mfcc_z"], out["names_mfcc_z"], _ = run(os.path.join(ROOT, "tests", "configs", "mfcc_0_d_a_z.conf"), pz, 16000, 1)
    out["mfcc_z_plain"], _, _ = run(os.path.join(ROOT, "tests", "configs", "mfcc_0_d_a.conf"), pz, 16000, 1, csv_out=False)
    for key, rel in (("ref_mfcc_0_z", "mfcc/MFCC12_0_D_A_Z.conf"), ("ref_mfcc_e_z", "mfcc/MFCC12_E_D_A_Z.conf"),
                     ("ref_plp_0_z", "plp/PLP_0_D_A_Z.conf"), ("ref_plp_e_z", "plp/PLP_E_D_A_Z.conf")):
        out[key], _, _ = run(os.path.join(refrun.CONFIG_DIR, rel), pz, 16000, 1, csv_out=False)
    # further shipped configurations made of LLD-path components only
    out["ref_audspec"], _, _ = run(os.path.join(refrun.CONFIG_DIR, "audspec", "audspec.conf"), pz, 16000, 1, csv_out=False)
    out["ref_audspec_compat"], _, _ = run(os.path.join(refrun.CONFIG_DIR, "audspec", "audspec_compat.conf"), pz, 16000, 1, csv_out=False)
    out["ref_spectrogram"], _, _ = run(os.path.join(refrun.CONFIG_DIR, "spectrum", "spectrogram.conf"), pz[:4000], 16000, 1, csv_out=False)
    out["ref_prosody_acf"], out["names_ref_prosody_acf"], _ = run(os.path.join(refrun.CONFIG_DIR, "prosody", "prosodyAcf.conf"), pz, 16000, 1)
    with tempfile.TemporaryDirectory() as d:           # demo1_energy.conf writes CSV only (-O names the csv file)
        wav, csv = os.path.join(d, "in.wav"), os.path.join(d, "out.csv")
        refrun.write_wav(wav, pz, 16000, 1)
        subprocess.run([refrun.SMILEXTRACT, "-C", os.path.join(refrun.CONFIG_DIR, "demo", "demo1_energy.conf"), "-I", wav, "-O", csv, "-l", "0"],
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        out["ref_demo1_energy_csv"] = np.frombuffer(open(csv, "rb").read(), dtype=np.uint8)
    # taps of compare_ns.conf (the same graph with the HTK sink moved): RASTA-filtered bands and the two sums
    tap = os.path.join(ROOT, "tests", "configs", "_cmp_taps.conf")
    with open(tap, "w") as f:
        f.write(open(cns).read().replace("reader.dmLevel = lld;lld_de\nfilename = \\cm[output(O)", "reader.dmLevel = audR;audSum;audRSum\nfilename = \\cm[output(O)"))
    out["cmp_taps"], _, _ = run(tap, pcm, 16000, 1, csv_out=False)
    os.remove(tap)
    for k, v in out.items():
        print(k, v.shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "conf_goldens.npz"), **out)


if __name__ == "__main__":
    main()

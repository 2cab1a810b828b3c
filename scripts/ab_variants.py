"""A/B timing of library variants (opensmile_b200/variants/lib_*.so, built with -DOSM_OPT_* switches):
runs the cfg-2 batch through each and prints the fused kernel's time."""
import glob, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for so in sorted(glob.glob(os.path.join(root, "opensmile_b200", "variants", "lib_*.so"))):
    env = dict(os.environ, OSM_B200_LIB=so)
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "prof_target.py"), "12"], env=env, capture_output=True, text=True)
    print(os.path.basename(so), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])

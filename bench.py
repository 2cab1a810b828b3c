#!/usr/bin/env python
"""bench.py -- LLD frames/s of the fused MFCC12_0_D_A path (BASELINE.json configs[1]).

A "step" is one pass of the hot path over one batch: 2000 synthetic 16 kHz mono utterances of
80 240 samples = exactly 500 frames each = 1 000 000 LLD frames per GPU (weak scaling: every
rank owns its own batch; the path has no data-path collective, NCCL only carries the timing /
counter reduction).

  value : frames/s with PCM already resident in HBM (osm_b200_plan_run_device), CUDA events,
          barrier + synchronize on both sides, max over ranks
  e2e   : the same metric through the C ABI's host entry point (osm_b200_plan_run_host) with
          pinned HOST buffers: H2D of the PCM + kernels + D2H of the LLD rows inside the
          timed region, every step
  roofline     : the fused per-frame kernel, algorithmic bytes (476 B/frame) over
                 its CUDA-event duration vs the measured HBM copy bandwidth
  cpu_baseline : the UNMODIFIED reference (oracle/_ref/SMILExtract, one process per host core)
                 on a bounded sample of the same workload

`--impl reference` times the reference's own CPU implementation (SMILExtract) instead.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_UTT = 2000
UTT_LEN = 400 + 160 * 499        # 80 240 samples -> exactly 500 frames
FRAMES_PER_UTT = 500
SAMPLE_RATE = 16000
BYTES_PER_FRAME = 160 * 2 + 39 * 4   # algorithmic: each PCM sample read once, each LLD value written once
WORKLOAD = "MFCC12_0_D_A, synthetic 16 kHz mono int16, 2000 utterances x 500 frames = 1M frames per GPU"
WORKLOAD_KEY = "mfcc12"
REF_CONF = "mfcc/MFCC12_0_D_A.conf"
REF_OUT_OPT = "-O"
N_COLS = 39


def select_workload(name):
    """--workload compare16: BASELINE configs[3], the shipped config/compare16/ComParE_2016.conf (full LLD set,
    65 + 65 columns incl. the SHS pitch chain) on 3.0 s utterances: 296 LLD rows each."""
    global N_UTT, UTT_LEN, FRAMES_PER_UTT, BYTES_PER_FRAME, WORKLOAD, WORKLOAD_KEY, REF_CONF, REF_OUT_OPT, N_COLS
    if name == "mfcc12":
        return
    assert name in ("compare16", "egemaps")
    WORKLOAD_KEY = name
    N_UTT = int(os.environ.get("OSM_BENCH_N_UTT", "10000"))
    UTT_LEN = 48000
    FRAMES_PER_UTT = 296                 # min(295 + 1, 299 + 1) rows of the lld level (SURVEY.md 8a')
    REF_OUT_OPT = "-lldhtkoutput"
    if name == "egemaps":
        # BASELINE configs[2]: the shipped config/egemaps/v02/eGeMAPSv02.conf, 25 LLD columns incl. the formant / harmonics
        # chain.  Its kernels have not run on a device yet: use this workload only to measure them.
        N_COLS = 25
        REF_CONF = "egemaps/v02/eGeMAPSv02.conf"
        WORKLOAD = ("eGeMAPSv02 LLD set (config/egemaps/v02/eGeMAPSv02.conf unchanged, 25 columns), synthetic 16 kHz mono "
                    "int16, %d utterances x 3.0 s = %d rows per GPU" % (N_UTT, N_UTT * FRAMES_PER_UTT))
    else:
        N_COLS = 130
        REF_CONF = "compare16/ComParE_2016.conf"
        WORKLOAD = ("ComParE_2016 full LLD set (config/compare16/ComParE_2016.conf unchanged, 130 columns), synthetic 16 kHz mono "
                    "int16, %d utterances x 3.0 s = %d rows per GPU" % (N_UTT, N_UTT * FRAMES_PER_UTT))
    BYTES_PER_FRAME = 160 * 2 + N_COLS * 4


# ------------------------------------------------------------------------------------------
def synth_batch_torch(n_utt, utt_len, device, seed):
    """Voiced-like harmonic source + noise (SURVEY.md 8d formula), generated on the device."""
    import torch
    g = torch.Generator(device=device).manual_seed(1234 + seed)
    out = torch.empty(n_utt * utt_len, dtype=torch.int16, device=device)
    chunk = 100                                  # utterances per chunk (bounds temporaries)
    t = torch.arange(utt_len, device=device, dtype=torch.float32) / SAMPLE_RATE
    for u0 in range(0, n_utt, chunk):
        n = min(chunk, n_utt - u0)
        ph0 = torch.rand(n, 1, device=device, generator=g) * 6.2831853
        f0 = 120.0 + 30.0 * torch.sin(6.2831853 * 0.5 * t[None, :] + ph0)
        phi = 6.2831853 * torch.cumsum(f0, dim=1) / SAMPLE_RATE
        x = torch.zeros(n, utt_len, device=device)
        for k in range(1, 20):
            x += torch.sin(k * phi) / k
        x = 0.1 * x + 0.02 * torch.randn(n, utt_len, device=device, generator=g)
        x = (x.clamp(-1, 1) * 32767.0).round().to(torch.int16)
        out[u0 * utt_len:(u0 + n) * utt_len] = x.reshape(-1)
    return out


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.lines = []
        self.idx = gpu_index

    def start(self):
        if shutil.which("nvidia-smi") is None:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.03)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profile_traffic():
    """dram bytes per launch of the fused kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("lld_kernel_dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------
# reference arm: the unmodified SMILExtract over WAV files, one process per host core
# ------------------------------------------------------------------------------------------
def _ref_worker(args):
    from oracle import refrun
    files, outdir, key = args
    select_workload(key)
    n = 0
    for wav in files:
        out = os.path.join(outdir, "%s.%d.htk" % (os.path.basename(wav), os.getpid()))
        subprocess.run([refrun.SMILEXTRACT, "-C", os.path.join(refrun.CONFIG_DIR, REF_CONF), "-I", wav, REF_OUT_OPT, out,
                        "-l", "0"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        n += refrun.read_htk(out)[1]["n"]
        os.remove(out)
    return n


def reference_sample(n_files, workers, tmpdir, seed=0):
    """Run the reference on n_files synthetic utterances with `workers` parallel processes.
    Returns (frames, seconds)."""
    from concurrent.futures import ProcessPoolExecutor
    from opensmile_b200.synth import voiced_pcm
    from oracle import refrun
    base = [voiced_pcm(UTT_LEN, SAMPLE_RATE, seed=seed + i) for i in range(8)]
    files = []
    for i in range(n_files):
        w = os.path.join(tmpdir, "u%05d.wav" % i)
        refrun.write_wav(w, base[i % len(base)], SAMPLE_RATE)
        files.append(w)
    shards = [files[i::workers] for i in range(workers)]
    shards = [s for s in shards if s]
    with ProcessPoolExecutor(max_workers=len(shards)) as ex:
        list(ex.map(_ref_worker, [([files[0]], tmpdir, WORKLOAD_KEY)] * len(shards)))       # warm the page cache / binaries
        t0 = time.perf_counter()
        frames = sum(ex.map(_ref_worker, [(s, tmpdir, WORKLOAD_KEY) for s in shards]))
        dt = time.perf_counter() - t0
    for w in files:
        os.remove(w)
    return frames, dt


def cpu_baseline(n_files=None):
    from oracle import refrun
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        if refrun.available():
            if n_files is None:
                n_files = max(64, min(2000, 48 * cores))   # ~13 ms of CPU work per file
                if WORKLOAD_KEY in ("compare16", "egemaps"):
                    n_files = max(16, 4 * cores)           # ~55 ms of CPU work per 3 s file
            frames, dt = reference_sample(n_files, cores, tmp)
            return {"value": frames / dt, "unit": "frames/s", "cores": cores, "kind": "reference",
                    "sample": "%d of the %d utterances (%d frames) through oracle/_ref/SMILExtract -C "
                              "%s, one process per core, WAV in /dev/shm -> HTK out, %.2f s"
                              % (n_files, N_UTT, frames, REF_CONF, dt)}
        # the reference binary did not travel: time the C restatement instead (single thread)
        from opensmile_b200.synth import voiced_pcm
        from oracle import oracle
        pcm = voiced_pcm(UTT_LEN, SAMPLE_RATE, seed=0)
        n = 40
        t0 = time.perf_counter()
        for _ in range(n):
            oracle.mfcc_d_a(pcm, float(SAMPLE_RATE))
        dt = time.perf_counter() - t0
        return {"value": n * FRAMES_PER_UTT / dt, "unit": "frames/s", "cores": 1, "kind": "port",
                "sample": "%d utterances through oracle/liboracle.so (double-precision FFT restatement), %.2f s" % (n, dt)}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import refrun
    cores = os.cpu_count() or 1
    n_files = max(64, min(2000, 48 * cores))
    if WORKLOAD_KEY in ("compare16", "egemaps"):
        n_files = max(16, 4 * cores)
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        if not refrun.available():
            cb = cpu_baseline()
            v = cb["value"]
            ms = FRAMES_PER_UTT * 1e3 / v
            steps_done = 0
        else:
            for _ in range(min(args.warmup, 1)):
                reference_sample(max(8, n_files // 8), cores, tmp)
            tot_f, tot_t = 0, 0.0
            for s in range(args.steps):
                fr, dt = reference_sample(n_files, cores, tmp, seed=s)
                tot_f += fr; tot_t += dt
            v = tot_f / tot_t
            ms = tot_t / args.steps * 1e3
            cb = {"value": v, "unit": "frames/s", "cores": cores, "kind": "reference",
                  "sample": "per step %d of the %d utterances (%d frames) through oracle/_ref/SMILExtract, "
                            "one process per host core" % (n_files, N_UTT, n_files * FRAMES_PER_UTT)}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps({
        "impl": "reference", "metric": "LLD frames/sec (16kHz, 25ms/10ms)", "value": v, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference is single-threaded per process; %d processes" % cores},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args, rank, world, local_rank):
    import torch
    from opensmile_b200 import Plan, components_mfcc12_0_d_a
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    if WORKLOAD_KEY in ("compare16", "egemaps"):
        from opensmile_b200 import Session
        conf = os.path.join(ROOT, "oracle", "_ref", "config", *REF_CONF.split("/"))
        sess = Session(conf, options={"lldhtkoutput": "x.htk"}, device=-1)     # conf front end only; the plan below computes
        comps, level = sess.components(float(SAMPLE_RATE), 1)
        plan = Plan(list(comps), level, device=local_rank)
    else:
        plan = Plan(components_mfcc12_0_d_a(float(SAMPLE_RATE)), "lld", device=local_rank)
    off = np.arange(N_UTT + 1, dtype=np.int64) * UTT_LEN
    fo = plan.frame_offsets(off)
    rows = int(fo[-1])
    assert rows == N_UTT * FRAMES_PER_UTT
    d_pcm = synth_batch_torch(N_UTT, UTT_LEN, dev, seed=rank)
    d_out = torch.empty((rows, plan.num_elements), dtype=torch.float32, device=dev)
    h_pcm = torch.empty(N_UTT * UTT_LEN, dtype=torch.int16).pin_memory()
    h_pcm.copy_(d_pcm)
    h_out = torch.empty((rows, plan.num_elements), dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        plan.run_device(d_pcm, off, d_out=d_out, frame_offsets=fo)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lld_ms, post_ms, launches = [], [], 0
    # per-step kernel times need a sync each step; take them in a separate pass so the timed loop
    # below stays free of host synchronisation
    ev0.record()
    for _ in range(args.steps):
        plan.run_device(d_pcm, off, d_out=d_out, frame_offsets=fo)
        launches += plan.last_launch_count()
    ev1.record()
    barrier()
    dt_ms = ev0.elapsed_time(ev1)
    for _ in range(min(args.steps, 10)):
        plan.run_device(d_pcm, off, d_out=d_out, frame_offsets=fo)
        a, b = plan.last_kernel_times()
        lld_ms.append(a); post_ms.append(b)

    # ---- end to end through the host entry point (pinned host buffers); the clock sampler keeps
    # running over this second timed region ----
    for _ in range(max(1, min(args.warmup, 3))):
        plan.run_host(h_pcm, off, out=h_out, frame_offsets=fo)
    barrier()
    e2e_steps = args.steps
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.run_host(h_pcm, off, out=h_out, frame_offsets=fo)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    checksum = float(h_out[::997].double().abs().sum())
    assert np.isfinite(checksum)

    # the only communication of the job: SUM of frame counters, MAX of times over ranks (NCCL)
    from opensmile_b200.dist import reduce_counters
    frames_all, dt_s = reduce_counters(rows * args.steps, dt_ms * 1e-3, dist, dev)
    frames_e2e, e2e_s = reduce_counters(rows * e2e_steps, e2e_s, dist, dev)
    dt_ms, e2e_ms = dt_s * 1e3, e2e_s * 1e3

    if rank == 0:
        value = frames_all / (dt_ms * 1e-3)
        e2e_value = frames_e2e / (e2e_ms * 1e-3)
        peak, peak_src = measured_peak_hbm()
        k_ms = statistics.mean(lld_ms)
        achieved = rows * BYTES_PER_FRAME / (k_ms * 1e-3) / 1e9
        cb = cpu_baseline() if world == 1 else None
        line = {
            "metric": "LLD frames/sec (16kHz, 25ms/10ms)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": rows,
                       "l2": "no flush needed: per step %d MB PCM in + %d MB rows out exceed the 126 MB L2"
                             % (N_UTT * UTT_LEN * 2 // 1000000, rows * plan.num_elements * 4 // 1000000),
                       "parallelism": "utterance shards, one rank per GPU, no data-path collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h_pcm.numel() * 2),
                    "d2h_bytes_per_step": int(h_out.numel() * 4), "steps": e2e_steps,
                    "api": "osm_b200_plan_run_host (pinned host buffers)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": profile_traffic() if WORKLOAD_KEY == "mfcc12" else None,
                         "kernel": "lld_kernel<256,32,256,2,VEC2,MFCC>" if WORKLOAD_KEY == "mfcc12" else "FFT front-end passes (lld_kernel) of the %d launches of a step" % (launches // max(args.steps, 1)),
                         "kernel_ms": k_ms, "post_kernel_ms": statistics.mean(post_ms),
                         "algorithmic_bytes_per_launch": rows * BYTES_PER_FRAME, "peak_source": peak_src},
        }
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    plan.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mfcc12", choices=["mfcc12", "compare16", "egemaps"],
                    help="mfcc12 = BASELINE configs[1] (default, the quoted metric); compare16 = configs[3], full ComParE_2016 LLD set; "
                         "egemaps = configs[2], eGeMAPSv02 LLD set (kernels pending their first device run)")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

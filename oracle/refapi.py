"""The UNMODIFIED reference through its own C API, in process (oracle/_ref/libSMILEapi.so, progsrc/include/smileapi/SMILEapi.h).

TEST INFRASTRUCTURE ONLY (bench.py's cpu_baseline / --impl reference legs).

One `smile_initialize` per worker process, then per utterance `smile_run` + `smile_reset` -- the start-up-free way to push
many files through the reference: component registration, config parsing and process creation (which dominate a
one-SMILExtract-per-5-second-file run) are paid once per core instead of once per file.  The shipped
configurations read their input through cWaveSource, whose file name is fixed at initialisation, so every utterance is
copied to the worker's fixed /dev/shm path before its run (a page-cache copy of <= 300 KB).  `smile_reset` re-creates the
sinks, which truncates the output file of the previous run: rows are therefore counted by the frame rule, and verified once
per worker at the end (run without reset, then read the file).
"""
import ctypes as C
import os
import shutil
import time

from . import refrun

LIB = os.path.join(refrun.REF_DIR, "libSMILEapi.so")


class _Opt(C.Structure):
    _fields_ = [("name", C.c_char_p), ("value", C.c_char_p)]


def available():
    return os.path.exists(LIB) and os.path.isdir(refrun.CONFIG_DIR)


class Extractor:
    def __init__(self, conf_rel, out_opt, workdir, tag):
        self.L = C.CDLL(LIB)
        L = self.L
        L.smile_new.restype = C.c_void_p
        L.smile_initialize.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(_Opt), C.c_int, C.c_int, C.c_int, C.c_char_p]
        for f in ("smile_run", "smile_reset", "smile_free"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.smile_error_msg.argtypes = [C.c_void_p]
        L.smile_error_msg.restype = C.c_char_p
        self.wav = os.path.join(workdir, "in_%s.wav" % tag)
        self.out = os.path.join(workdir, "out_%s.htk" % tag)
        self.obj = L.smile_new()
        self.opts = (_Opt * 2)(_Opt(b"I", self.wav.encode()), _Opt(out_opt.lstrip("-").encode(), self.out.encode()))
        self.conf = os.path.join(refrun.CONFIG_DIR, conf_rel).encode()
        self.ready = False

    def _init(self):
        if self.L.smile_initialize(self.obj, self.conf, 2, self.opts, 0, 0, 0, None) != 0:
            raise RuntimeError("smile_initialize: %s" % self.L.smile_error_msg(self.obj))
        self.ready = True

    def run_file(self, src_wav, reset=True):
        shutil.copyfile(src_wav, self.wav)
        if not self.ready:
            self._init()                       # needs the input file to exist
        if self.L.smile_run(self.obj) != 0:
            raise RuntimeError("smile_run: %s" % self.L.smile_error_msg(self.obj))
        if reset and self.L.smile_reset(self.obj) != 0:
            raise RuntimeError("smile_reset: %s" % self.L.smile_error_msg(self.obj))

    def close(self):
        self.L.smile_free(self.obj)
        for p in (self.wav, self.out):
            if os.path.exists(p):
                os.remove(p)


def worker(args):
    """(files, conf_rel, out_opt, workdir, rows_per_file, warm) -> (rows, seconds of the timed loop, rows of the last file)
    The last file is run without a reset so that its output can be read back: the count check of the leg."""
    files, conf_rel, out_opt, workdir, rows_per_file, warm = args
    ex = Extractor(conf_rel, out_opt, workdir, str(os.getpid()))
    try:
        for f in files[:warm]:
            ex.run_file(f)
        t0 = time.perf_counter()
        for f in files[:-1]:
            ex.run_file(f)
        ex.run_file(files[-1], reset=False)
        dt = time.perf_counter() - t0
        ex.L.smile_free(ex.obj)                # finalises the sink's header
        ex.obj = ex.L.smile_new()
        n_last = refrun.read_htk(ex.out)[1]["n"]
        return len(files) * rows_per_file, dt, n_last
    finally:
        ex.close()

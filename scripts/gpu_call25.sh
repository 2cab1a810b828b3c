#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/dbg_gemaps_func.py > gpurun_out/c25_gemaps.txt 2>&1
tail -100 gpurun_out/c25_gemaps.txt

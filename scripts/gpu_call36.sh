#!/bin/bash
# round 2, GPU call 36: the rows widened late (melspec level / scales, cFFTmagphase variants, delta variants) on the GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_melspec_scales.py tests/test_spectrogram_variants.py tests/test_delta_variants.py -m gpu -q 2>&1 | grep -E "passed|failed|Error|assert|FAILED" | cut -c1-500 | tail -20 | tee gpurun_out/c36_new_rows.txt

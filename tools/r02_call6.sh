#!/bin/bash
# Round 2, call 6 (1 GPU): bounded-CTA rest updates + stream priorities (real overlap of the panel chain with the update)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== 1. parity subset"
timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_tcgen05_insitu.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r02c6_tests.log
echo "== 2. C4h / C4 / C3-size fits vs bounded-CTA size"
for ch in 0 4 8 16 32; do echo "chunk $ch"; AGP_OZAKI_CHUNK=$ch timeout 200 python tools/fit_once.py 32768 64 4; done
for ch in 0 8 16; do echo "chunk $ch (C4)"; AGP_OZAKI_CHUNK=$ch timeout 300 python tools/fit_once.py 65536 64 3; done
for ch in 0 8; do echo "chunk $ch (fp32 16384)"; AGP_OZAKI_CHUNK=$ch timeout 200 python tools/fit_once.py 16384 32 4 f32; done
for ch in 0 8; do echo "chunk $ch (8192)"; AGP_OZAKI_CHUNK=$ch timeout 200 python tools/fit_once.py 8192 16 5; done

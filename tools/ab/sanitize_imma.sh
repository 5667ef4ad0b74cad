#!/bin/bash
# compute-sanitizer over the integer-tensor-pipe u8 kernels (w5i::p25_ddc5_imma_kernel, w50i::p25_ddc50_imma_kernel): the chunk-phase /
# minimal-chunk test (both kernels, both power instantiations, pieces with warm-up slices that start in front of the tail) and smoke().
mkdir -p gpurun_out
{
echo "compute-sanitizer on a B200 (gpurun): w5i / w50i kernels (TMA-staged slices read as mma.sync u8 x s8 fragments)"
echo "== memcheck: pytest tests/test_gpu_round2.py -k imma_kernels"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "imma_kernels" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|FAILED|Invalid|at 0x" | tail -8
echo "== racecheck: the same"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "imma_kernels" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|FAILED|hazard" | tail -8
echo "== synccheck: the same"
timeout 200 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "imma_kernels" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|FAILED" | tail -5
echo "== initcheck (global memory): the same"
timeout 200 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "imma_kernels" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|FAILED|Uninit" | tail -5
} > gpurun_out/r02_sanitizer_imma.txt 2>&1
cat gpurun_out/r02_sanitizer_imma.txt

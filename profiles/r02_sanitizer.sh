#!/bin/bash
# compute-sanitizer over the kernel tests of the round-2 kernels (memcheck on all of them; racecheck on the shared-memory
# heavy ones).  Sizes are the tests' own; -k keeps the long parametrisations out.
OUT=gpurun_out; mkdir -p $OUT
SEL="tests/test_gpu_dense.py tests/test_gpu_field_fused.py tests/test_gpu_gridencoder.py tests/test_gpu_field_mlp.py"
{
echo "== memcheck: $SEL (+ occupancy update, fused step with the loss scaler)"
timeout 1500 compute-sanitizer --tool memcheck --target-processes all python -m pytest $SEL tests/test_gpu_render.py tests/test_gpu_fused_step.py -m gpu -q -x \
   -k "not 100003 and not 70001 and not at_scale and not nccl and not reference_bitfield" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -8
echo "== racecheck: dense sampler, fused field kernel, occupancy"
timeout 1500 compute-sanitizer --tool racecheck --target-processes all python -m pytest tests/test_gpu_dense.py tests/test_gpu_field_fused.py -m gpu -q -x \
   -k "sampler_kernels or fused_inference_forward or honours_the_device" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -8
} | tee $OUT/r02_sanitizer.txt

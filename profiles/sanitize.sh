#!/bin/bash
# compute-sanitizer (memcheck / racecheck / synccheck) over the tcgen05 + TMA + mbarrier kernels on a small grid:
# the dim-32 model case (30x14x12, 3 levels: winz 64->64, win 128/256, winp 32->32, fold2, per-tap, wgrad, attention).
# Usage (on a GPU box): bash profiles/sanitize.sh   -> gpurun_out/sanitize_<tool>.log + gpurun_out/sanitize_summary.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='test_denoiser_bf16_within_tolerance and dim32'
SELB='test_denoiser_backward_matches_oracle and dim32'
: > gpurun_out/sanitize_summary.txt
for tool in memcheck synccheck racecheck; do
  for sel in "$SEL" "$SELB"; do
    tag=$(echo "$sel" | cut -d' ' -f1)
    log=gpurun_out/sanitize_${tool}_${tag}.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "$sel" > "$log" 2>&1
    echo "[$tool] $sel: exit $? | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' "$log" | tr '\n' ' ')" >> gpurun_out/sanitize_summary.txt
  done
  log=gpurun_out/sanitize_${tool}_attention.log
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "test_attention and bf16" > "$log" 2>&1
  echo "[$tool] test_attention bf16 (tcgen05): exit $? | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' "$log" | tr '\n' ' ')" >> gpurun_out/sanitize_summary.txt
done
cat gpurun_out/sanitize_summary.txt

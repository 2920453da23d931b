#!/bin/bash
# compute-sanitizer over the kernels added this round (memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards in the environment kernels)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
K="copy_contact or local_poses or touch_events or removed_actor or material_table_gpu_matches_oracle or test_gpu_matches_oracle or hull_contacts or config3_shape or env_path_matches_oracle_ragged or state_export_matches"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > $O/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/sanitize_memcheck.log | head -12
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "env_path_matches_oracle_ragged or local_poses_gpu or test_gpu_matches_oracle" > $O/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $O/sanitize_racecheck.log | head -12

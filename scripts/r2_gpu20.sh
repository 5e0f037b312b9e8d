#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize.py > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -c "Invalid\|out of bounds\|Misaligned" gpurun_out/r2_sanitize_memcheck.log; tail -4 gpurun_out/r2_sanitize_memcheck.log

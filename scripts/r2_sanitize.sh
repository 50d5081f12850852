#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/sanitize_small.py > gpurun_out/r02_sanitize_plain.log 2>&1; echo "plain exit $?"; tail -4 gpurun_out/r02_sanitize_plain.log
for tool in memcheck synccheck racecheck; do
timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r02_sanitize_$tool.log 2>&1; echo "$tool exit $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE TOUR|hazard" gpurun_out/r02_sanitize_$tool.log | sort | uniq -c | head -12
done

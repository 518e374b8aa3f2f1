#!/bin/bash
# compute-sanitizer memcheck + racecheck over profiles/sanitize_cases.py (on the GPU box, from the repo root):
#   bash profiles/sanitize.sh r02      -> gpurun_out/r02_{memcheck,racecheck}.log (+ one-line summaries on stdout)
tag=${1:-rXX}
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_cases.py > gpurun_out/${tag}_$tool.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_CASES_OK' gpurun_out/${tag}_$tool.log | tr '\n' ' ')"
done

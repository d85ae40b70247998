#!/bin/bash
# compute-sanitizer over the small-shape pass (synccheck is mandatory for hand-rolled mbarrier / TMEM code)
mkdir -p gpurun_out
for tool in ${TOOLS:-memcheck synccheck racecheck}; do
  for part in ${PARTS:-cells text search}; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py $part > gpurun_out/sanitize_${tool}_${part}.log 2>&1
    echo "$tool $part rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_${part}.log | tail -n 1)"
  done
done

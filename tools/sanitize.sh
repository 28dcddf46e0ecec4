#!/bin/bash
# compute-sanitizer over the small end-to-end smoke (all three phases + GCO) and one test per
# phase-B path; run under gpurun.  Output: gpurun_out/sanitizer_<tool>.txt
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.txt 2>&1
  tail -4 gpurun_out/sanitizer_$tool.txt
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 \
  python -m pytest tests/test_gpu_paths.py -x -q -k "far_from or degree or strong or more_than" > gpurun_out/sanitizer_memcheck_paths.txt 2>&1
tail -5 gpurun_out/sanitizer_memcheck_paths.txt

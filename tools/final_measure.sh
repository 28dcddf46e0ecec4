#!/bin/bash
# usage (under gpurun): tools/final_measure.sh   -- the lines committed under profiles/r2_bench_*.json
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_cfg5band.json 2> gpurun_out/r2_bench_n1_cfg5band.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_n1_reference_arm.json 2>/dev/null
for wl in cfg1_chr21_example cfg2_chr21_chr22 cfg3_chr1_50kb cfg4_genome_50kb; do
  python bench.py --steps 10 --warmup 3 --workload $wl > gpurun_out/r2_bench_n1_$wl.json 2> gpurun_out/r2_bench_n1_$wl.err
done
python tools/pcie_probe.py > gpurun_out/r2_pcie.json
for f in gpurun_out/r2_bench_n1_*.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().splitlines()[-1])
if d.get("impl") == "reference":
    print(sys.argv[1], "reference arm", d["value"])
else:
    print(sys.argv[1].split("n1_")[1], "ms/step %.3f value %.3e frac %.3f B-frac %.3f e2e %.1f ms %.3e cpu %s vec %s" % (
        d["ms_per_step"], d["value"], d["roofline"]["step"]["frac_of_slower_roof"], d["roofline"]["frac"],
        d["e2e"]["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"] and "%.2e" % d["cpu_baseline"]["value"],
        d["cpu_vectorised"] and "%.2e" % d["cpu_vectorised"]["value"]))
PY
done
bash tools/sanitize.sh 2>&1 | tail -12

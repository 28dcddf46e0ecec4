nproc; free -g | head -2
( time python bench.py --steps 10 --warmup 3 --e2e-steps 3 > gpurun_out/${TAG}_cfg5.json 2> gpurun_out/${TAG}_cfg5.err ) 2>&1 | grep real
tail -3 gpurun_out/${TAG}_cfg5.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_cfg5.json')); print('cfg5', d['roofline']['phase_ms'], d['ms_per_step'], 'e2e', d['e2e'], d['cpu_baseline'], d['cpu_vectorised'])"
for wl in cfg1_chr21_example cfg2_chr21_chr22 cfg3_chr1_50kb cfg4_genome_50kb; do
( time python bench.py --steps 10 --warmup 3 --no-cpu --workload $wl --e2e-steps 3 > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err ) 2>&1 | grep real
tail -3 gpurun_out/${TAG}_$wl.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_$wl.json')); print('$wl', d['roofline']['phase_ms'], d['ms_per_step'], d['roofline']['step']['frac_of_slower_roof'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'value', d['value'])"
done

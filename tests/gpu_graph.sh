for w in c2 c1; do
timeout 200 python bench.py --workload $w --steps 400 --warmup 10 > gpurun_out/p_$w.json 2>gpurun_out/p_$w.err; python -c "
import json; d=json.load(open('gpurun_out/p_$w.json')); print('$w', '%.4g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['e2e']['d2h_bytes_per_step'], 'frac %.3f'%d['roofline']['frac'], 'cpu %.3g'%d['cpu_baseline']['value'])" || tail -5 gpurun_out/p_$w.err
done

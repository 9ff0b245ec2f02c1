timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --workload c4 --steps 400 --warmup 10 --no-cpu-baseline > gpurun_out/p_c4.json 2>gpurun_out/p_c4.err; python -c "
import json; d=json.load(open('gpurun_out/p_c4.json')); print('c4 400', '%.4g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['e2e']['d2h_bytes_per_step'], 'frac %.3f'%d['roofline']['frac'])" || tail -5 gpurun_out/p_c4.err

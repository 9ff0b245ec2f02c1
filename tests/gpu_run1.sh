set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python oracle/record_c3_polys.py 2>&1 | tail -3
timeout 600 python bench.py --steps 300 --warmup 10 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err; tail -3 gpurun_out/r2a_bench_c3.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2a_bench_c3.json'))
    print('c3 value %.4g ms/step %.4f e2e %.4g'%(d['value'],d['ms_per_step'],d['e2e']['value']))
    print([(k['name'],round(k['ms'],4),k.get('items_per_launch')) for k in d['roofline']['per_kernel']])
    print('ring_fraction',d['ring_fraction'],'success',d['success_rate'])
except Exception as ex: print('bench failed',ex)
PY
timeout 300 python bench.py --workload c2 --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
timeout 300 python bench.py --workload c1 --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/r2a_bench_c1.json 2> gpurun_out/r2a_bench_c1.err
python - <<'PY'
import json
for w in ('c2','c1'):
    try:
        d=json.load(open(f'gpurun_out/r2a_bench_{w}.json'))
        print(w,'value %.4g ms/step %.4f e2e %.4g'%(d['value'],d['ms_per_step'],d['e2e']['value']), [ (k['name'],round(k['ms'],4)) for k in d['roofline']['per_kernel']])
    except Exception as ex: print(w,'failed',ex)
PY

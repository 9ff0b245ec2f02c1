timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_parity.py -x -q -k "host" 2>&1 | tail -15
for opt in "" "host_graph=0"; do
MESHGEN_OPTIONS=$opt timeout 600 python bench.py --steps 400 --warmup 10 --no-cpu-baseline > gpurun_out/h_c3_$opt.json 2>gpurun_out/h_c3_$opt.err; python -c "
import json; d=json.load(open('gpurun_out/h_c3_$opt.json')); print('c3 $opt', '%.4g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['e2e']['d2h_bytes_per_step'], 'frac %.3f'%d['roofline']['frac'], d['gpu_launches'])"
done

timeout 600 python -m pytest tests/test_gpu_parity_r2.py -x -q -k "device_side_step_index or graph_captured" 2>&1 | tail -3
for mode in "" "--no-graph"; do
timeout 600 python bench.py --steps 400 --warmup 10 --no-cpu-baseline $mode > gpurun_out/g_c3$mode.json 2>gpurun_out/g_c3$mode.err; python -c "
import json; d=json.load(open('gpurun_out/g_c3$mode.json')); print('c3 $mode', '%.4g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], d['gpu_launches'], d['launch_mode'].get('kernels_per_step'), d['clocks'])"
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/g_c3_20.json 2>gpurun_out/g_c3_20.err; python -c "
import json; d=json.load(open('gpurun_out/g_c3_20.json')); print('c3 20 steps', '%.4g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], d['gpu_launches'])"

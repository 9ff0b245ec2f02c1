for w in c2 c1; do
for mode in "" "--no-graph"; do
timeout 600 python bench.py --workload $w --steps 400 --warmup 10 --no-cpu-baseline $mode > gpurun_out/g_$w$mode.json 2>gpurun_out/g_$w$mode.err; python -c "
import json; d=json.load(open('gpurun_out/g_$w$mode.json')); print('$w $mode', '%.4g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], d['gpu_launches'], d['launch_mode'], d['clocks'])" || tail -5 gpurun_out/g_$w$mode.err
done
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/g_c3_20.json 2>gpurun_out/g_c3_20.err; python -c "
import json; d=json.load(open('gpurun_out/g_c3_20.json')); print('c3 20 steps', '%.4g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], d['gpu_launches'])"

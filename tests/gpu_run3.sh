set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/driver_like.json 2> gpurun_out/driver_like.err; python -c "
import json; d=json.load(open('gpurun_out/driver_like.json')); print('driver-like', '%.4g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])"
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 | cut -c1-300
for s in 401 402 403; do timeout 900 python tests/soak.py --seed $s 2>&1 | tail -1; done

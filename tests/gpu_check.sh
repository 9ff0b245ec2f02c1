timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 400 --warmup 10 --no-cpu-baseline > gpurun_out/final_c3.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/final_c3.json')); print('c3', '%.4g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], d['clocks'])"

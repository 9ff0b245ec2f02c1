# final 1-GPU measurements of a round: every workload with the CPU arm, the reference arm, config 5, a soak seed
# (SKIP_TESTS=1: without the GPU test suite)
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in c3 c4 c2 c1; do
  timeout 600 python bench.py --workload $w --steps 400 --warmup 10 > gpurun_out/r2_bench_${w}_n1.json 2> gpurun_out/r2_bench_${w}_n1.err
done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c3_n1_driverlike.json 2> gpurun_out/r2_bench_c3_n1_driverlike.err
python - <<'PY'
import json
for w in ('c3','c4','c2','c1','c3_n1_driverlike'):
    try:
        d=json.load(open(f'gpurun_out/r2_bench_{w}_n1.json' if len(w) == 2 else f'gpurun_out/r2_bench_{w}.json'))
        print(w,'value %.4g ms/step %.4f e2e %.4g frac %.3f memo_frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['memo_frac']), [ round(k['ms']*1e3,1) for k in d['roofline']['per_kernel']], 'cpu', d['cpu_baseline'] and '%.3g'%d['cpu_baseline']['value'], d['clocks'])
    except Exception as ex: print(w,'failed',ex)
PY
timeout 600 python examples/sac_rollout.py --envs 65536 --steps 300 --graph > gpurun_out/r2_c5_n1_graph.json 2> gpurun_out/r2_c5_n1.err; cut -c1-300 gpurun_out/r2_c5_n1_graph.json
timeout 600 python examples/sac_rollout.py --envs 65536 --steps 300 --graph --tf32 > gpurun_out/r2_c5_n1_graph_tf32.json 2>> gpurun_out/r2_c5_n1.err; cut -c1-300 gpurun_out/r2_c5_n1_graph_tf32.json
timeout 600 python examples/sac_rollout.py --envs 65536 --steps 300 > gpurun_out/r2_c5_n1_eager.json 2>> gpurun_out/r2_c5_n1.err; cut -c1-300 gpurun_out/r2_c5_n1_eager.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; cut -c1-400 gpurun_out/r2_bench_ref.json
for s in ${SOAK_SEEDS:-201 202}; do timeout 900 python tests/soak.py --seed $s 2>&1 | tail -3; done

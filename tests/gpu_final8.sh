# final multi-GPU measurements of a round (one node): scaling of the default workload, config 4 and config 5
N=${1:-8}
set -x
P=29500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N)) bench.py --gpus $N --steps 400 --warmup 10 > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err
if [ "$N" = 8 ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --workload c4 --steps 400 --warmup 10 > gpurun_out/r2_bench_c4_n8.json 2> gpurun_out/r2_bench_c4_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 examples/sac_rollout.py --envs 65536 --steps 300 --graph > gpurun_out/r2_c5_n8_graph.json 2> gpurun_out/r2_c5_n8_graph.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 examples/sac_rollout.py --envs 65536 --steps 300 --graph --tf32 > gpurun_out/r2_c5_n8_graph_tf32.json 2> gpurun_out/r2_c5_n8_graph_tf32.err
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_c*_n[248].json') + glob.glob('gpurun_out/r2_c5_n8*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,'value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), 'e2e %.4g'%d['e2e']['value'] if 'e2e' in d else '', d.get('collective',{}).get('inside_timed_loop'), d['clocks'])
    except Exception as ex: print(f,'failed',ex)
PY

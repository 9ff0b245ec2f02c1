# 2-GPU check of the bench (driver-style launch) and of the config-5 rollout
set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 400 --warmup 10 > gpurun_out/r2_bench_c3_n2.json 2> gpurun_out/r2_bench_c3_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_c3_n2_driverlike.json 2> gpurun_out/r2_bench_c3_n2_driverlike.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29504 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err
python - <<'PY'
import json
for f in ('r2_bench_c3_n2', 'r2_bench_c3_n2_driverlike', 'r2_bench_ref_n2'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), 'e2e %.4g'%d['e2e']['value'] if 'e2e' in d else '', d.get('collective',{}).get('inside_timed_loop'), d.get('gpu_launches'), d.get('clocks'))
    except Exception as ex: print(f,'failed',ex); print(open(f'gpurun_out/{f}.err').read()[-1500:])
PY

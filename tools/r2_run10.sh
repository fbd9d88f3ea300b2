cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"; tail -2 gpurun_out/bench_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config 5 --steps 3 --warmup 3 > gpurun_out/bench_8gpu_config5.json 2> gpurun_out/bench_8gpu_config5.err; echo "bench8 c5 rc=$?"; tail -2 gpurun_out/bench_8gpu_config5.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
python - <<'P'
import json
for f in ('gpurun_out/bench_8gpu.json','gpurun_out/bench_8gpu_config5.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('gather'), d['worlds'])
    except Exception as e: print(f,'ERR',e)
P

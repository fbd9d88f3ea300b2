# the bench under torchrun on 8 GPUs (config 3 and config 5) + the 2-rank NCCL parity test
cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_final_bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --config 5 --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_final_bench_8gpu_config5.json 2> gpurun_out/bench_8gpu_config5.err; echo "bench8 c5 rc=$?"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -1
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_final_bench_8gpu*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], '%.4g'%d['value'], '%.1f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d.get('gather',{}))
    except Exception as e: print(f,'ERR',e)
P

# usage: bash tools/r2_scale.sh N  -- the bench under torchrun on N GPUs (config 3; config 5 too when N = 8)
cd $GRAFT_REPO_ROOT
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench$N rc=$?"
if [ "$N" = "8" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --config 5 --steps 3 --warmup 3 > gpurun_out/r02_bench_8gpu_config5.json 2> gpurun_out/bench_8gpu_config5.err; echo "bench8 c5 rc=$?"
fi
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -1
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_${N}gpu*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], '%.4g'%d['value'], '%.1f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d.get('gather',{}).get('ms'))
    except Exception as e: print(f,'ERR',e)
P

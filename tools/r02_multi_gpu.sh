# usage: bash tools/r02_multi_gpu.sh N
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
$TR bench.py --gpus $N --steps 4 --warmup 2 --no-saturated --no-cpu-baseline --seeds-per-batch 0 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo bench_rc=$?
$TR bench.py --gpus $N --sweep64 --warmup 1 > gpurun_out/r02_sweep64_n$N.json 2> gpurun_out/r02_sweep64_n$N.err; echo sweep_rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1]); print('bench', d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['value_leg_equals_e2e_leg'])
d=json.loads(open('gpurun_out/r02_sweep64_n$N.json').read().strip().splitlines()[-1]); print('sweep64', d['n_gpus'], d['value'], d['ms_per_step'], d['checksum_of_checksums'], d['failures'])
"

# one GPU call: the whole -m gpu suite, the microbench sweeps, the variant crossover sweep and the bench line
T=${1:-r02b}
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${T}_pytest_gpu.log; tail -4 gpurun_out/${T}_pytest_gpu.log
python -m guided_attention_b200.microbench > gpurun_out/${T}_microbench_k1k2_tail.jsonl 2>&1
python -m guided_attention_b200.microbench --self > gpurun_out/${T}_microbench_self.jsonl 2>&1
python tools/crossover_sweep.py > gpurun_out/${T}_crossover_sweep.jsonl 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo bench_rc=$?

python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu_4.log; tail -4 gpurun_out/r02_pytest_gpu_4.log
python -m guided_attention_b200.microbench > gpurun_out/r02_microbench_k1k2_tail.jsonl 2>&1
python -m guided_attention_b200.microbench --self > gpurun_out/r02_microbench_self.jsonl 2>&1
python tools/crossover_sweep.py > gpurun_out/r02_crossover_sweep.jsonl 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo bench_rc=$?

python bench.py --steps 6 --warmup 3 --host-control --no-saturated --no-cpu-baseline --seeds-per-batch 0 > gpurun_out/r02_bench_host_control.json 2>/dev/null; echo host_rc=$?
python bench.py --steps 6 --warmup 3 --no-saturated --no-cpu-baseline --seeds-per-batch 0 > gpurun_out/r02_bench_device_control.json 2>/dev/null; echo dev_rc=$?
python bench.py --sweep64 --warmup 2 > gpurun_out/r02_sweep64_n1.json 2>/dev/null; echo sweep_rc=$?
python -c "
import json
for f in ('host_control','device_control'):
    d=json.loads(open('gpurun_out/r02_bench_%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d['ms_per_step'])
d=json.loads(open('gpurun_out/r02_sweep64_n1.json').read().strip().splitlines()[-1]); print('sweep64', d['value'], d['ms_per_step'], d['checksum_of_checksums'], d['failures'])
"

# ncu launch list of the bench command (host-driven graph replay: ncu does not follow device-side conditional nodes)
CMD="python bench.py --steps 1 --warmup 0 --denoise-steps 3 --no-cpu-baseline --no-saturated --seeds-per-batch 0 --host-control"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 30000 --launch-count 8000 --csv \
  --log-file gpurun_out/r02b_ncu_launches.csv $CMD > gpurun_out/r02b_ncu_launches.log 2>&1; echo ncu_rc=$?
python tools/ncu_launch_summary.py gpurun_out/r02b_ncu_launches.csv "$CMD (skip 30000, count 8000)" > gpurun_out/r02b_ncu_launches_summary.txt; head -20 gpurun_out/r02b_ncu_launches_summary.txt
gzip -f gpurun_out/r02b_ncu_launches.csv

set -x
N="ncu --set full --clock-control none --import-source on -f"
M="python -m guided_attention_b200.microbench"
timeout 200 $N -k regex:cross_attn_fwd_tc_pipe --launch-skip 3 --launch-count 1 -o gpurun_out/final_k1_pipe_flat_d40_b128 $M --single fwd 128 4096 40 > /dev/null 2>&1; echo rc=$?
timeout 200 $N -k regex:cross_attn_fwd_tc_pipe --launch-skip 3 --launch-count 1 -o gpurun_out/final_k1_pipe_maps_d80_b128 $M --single fwd 128 1024 80 maps > /dev/null 2>&1; echo rc=$?
timeout 200 $N -k regex:cross_attn_bwd_tc_pipe --launch-skip 3 --launch-count 1 -o gpurun_out/final_k2_pipe_d40_b128 $M --single bwd 128 4096 40 > /dev/null 2>&1; echo rc=$?
timeout 200 $N -k regex:cross_attn_fwd_tc_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/final_k1_single_d160_b1 $M --single fwd 1 256 160 maps > /dev/null 2>&1; echo rc=$?
timeout 200 $N -k regex:cross_attn_bwd_tc_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/final_k2_single_d160_b1 $M --single bwd 1 256 160 maps > /dev/null 2>&1; echo rc=$?
timeout 200 $N -k regex:self_attn_fwd --launch-skip 3 --launch-count 1 -o gpurun_out/final_sa_fwd_b1 $M --single-self fwd 1 4096 40 > /dev/null 2>&1; echo rc=$?
timeout 200 $N -k regex:self_attn_bwd --launch-skip 4 --launch-count 2 -o gpurun_out/final_sa_bwd_b1 $M --single-self bwd 1 4096 40 > /dev/null 2>&1; echo rc=$?
timeout 200 $N -k regex:tail_fwd --launch-skip 3 --launch-count 1 -o gpurun_out/final_tail_fwd_s2048 $M --single-tail fwd 16 2048 > /dev/null 2>&1; echo rc=$?
timeout 200 $N -k regex:tail_bwd --launch-skip 3 --launch-count 1 -o gpurun_out/final_tail_bwd_s2048 $M --single-tail bwd 16 2048 > /dev/null 2>&1; echo rc=$?
timeout 200 $N -k regex:tail_fwd --launch-skip 3 --launch-count 1 -o gpurun_out/final_tail_fwd_s1 $M --single-tail fwd 16 1 > /dev/null 2>&1; echo rc=$?

# ncu --set full captures of the UNet-side fused kernels (third round-2 pass), one launch each; summarised by
# tools/ncu_summary.py into profiles/r02c_ncu_unet_ops_summary.json
N="ncu --set full --clock-control none --import-source on -f"; M="python -m guided_attention_b200.microbench"; O=gpurun_out/ncu; mkdir -p $O
timeout 300 $N -k "regex:gn_(stats|apply)_kernel" --launch-skip 6 --launch-count 2 -o $O/r02c__group_norm_fwd__n1_c320_hw4096 $M --single-group-norm fwd 1 320 64 > /dev/null 2>&1; echo rc=$?
timeout 300 $N -k "regex:gn_(stats|apply)_kernel" --launch-skip 6 --launch-count 2 -o $O/r02c__group_norm_fwd__n8_c640_hw4096 $M --single-group-norm fwd 8 640 64 > /dev/null 2>&1; echo rc=$?
timeout 300 $N -k "regex:gn_bwd_(sums|apply)_kernel" --launch-skip 6 --launch-count 2 -o $O/r02c__group_norm_bwd__n8_c640_hw4096 $M --single-group-norm fwdbwd 8 640 64 > /dev/null 2>&1; echo rc=$?
timeout 300 $N -k "regex:geglu_fwd_kernel" --launch-skip 3 --launch-count 1 -o $O/r02c__geglu_fwd__rows8192_inner1280 $M --single-geglu fwd 8192 1280 > /dev/null 2>&1; echo rc=$?
timeout 300 $N -k "regex:geglu_bwd_kernel" --launch-skip 3 --launch-count 1 -o $O/r02c__geglu_bwd__rows8192_inner1280 $M --single-geglu fwdbwd 8192 1280 > /dev/null 2>&1; echo rc=$?
ls -la $O | tail -6

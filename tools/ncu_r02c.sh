# round-2 third pass: source-level captures of the three lowest HBM fractions (K1 maps d=80, K1 / K2 d=40 no maps)
N="ncu --set full --clock-control none --import-source on -f"; M="python -m guided_attention_b200.microbench"; O=gpurun_out/ncu; mkdir -p $O
timeout 300 $N -k "regex:cross_attn_fwd_tc_(pipe|stream)" --launch-skip 3 --launch-count 1 -o $O/r02c__cross_attn_fwd__N1024_d80_mapsTrue $M --single fwd 256 1024 80 maps > /dev/null 2>&1; echo rc=$?
timeout 300 $N -k "regex:cross_attn_fwd_tc_(pipe|stream)" --launch-skip 3 --launch-count 1 -o $O/r02c__cross_attn_fwd__N4096_d40_mapsFalse $M --single fwd 128 4096 40 > /dev/null 2>&1; echo rc=$?
timeout 300 $N -k "regex:cross_attn_bwd_tc_(pipe|stream)" --launch-skip 3 --launch-count 1 -o $O/r02c__cross_attn_bwd__N4096_d40_mapsFalse $M --single bwd 128 4096 40 > /dev/null 2>&1; echo rc=$?
ls -la $O

# Same-box A/B of the fused UNet ops (boxes of the pool differ by several percent: compare within one call only; the
# first and the last line are the same configuration, which bounds the drift inside the call).
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-saturated --seeds-per-batch 0"
O=gpurun_out/ab; mkdir -p $O; rm -f $O/*.json
$B > $O/1_all_on.json 2> $O/err.txt
GA_FUSED_DISABLE=temb $B > $O/2_no_temb.json 2>> $O/err.txt
GA_FUSED_DISABLE=temb,layernorm $B > $O/2_no_layernorm.json 2>> $O/err.txt
GA_FUSED_DISABLE=temb,layernorm,geglu $B > $O/3_no_layernorm_geglu.json 2>> $O/err.txt
GA_FUSED_DISABLE=temb,layernorm,geglu,conv,resnet $B > $O/4_norms_only.json 2>> $O/err.txt
GA_FUSED_NORM=0 $B > $O/5_stock.json 2>> $O/err.txt
$B > $O/6_all_on_again.json 2>> $O/err.txt
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'],4), 'img/s', round(d['ms_per_step'],1), 'ms')
    except Exception as e: print(f, 'ERR', e)
P

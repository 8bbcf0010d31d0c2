B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-saturated --seeds-per-batch 0"
O=gpurun_out/ab2; mkdir -p $O
GA_GN_PDL=0 $B > $O/1_no_pdl.json 2> $O/err.txt
$B > $O/2_all_on.json 2>> $O/err.txt
GA_FUSED_DISABLE=geglu $B > $O/3_no_geglu.json 2>> $O/err.txt
GA_FUSED_DISABLE=geglu GA_GN_PDL=0 $B > $O/4_no_geglu_no_pdl.json 2>> $O/err.txt
$B > $O/5_all_on_again.json 2>> $O/err.txt
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab2/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'],4), 'img/s', round(d['ms_per_step'],1), 'ms')
    except Exception as e: print(f, 'ERR', e)
P

# Same-box A/B of the fused UNet ops (boxes of the pool differ by several percent: compare within one call only).
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-saturated --seeds-per-batch 0"
O=gpurun_out/ab; mkdir -p $O
$B > $O/all_on.json 2> $O/all_on.err; echo rc=$?
GA_GN_PDL=0 $B > $O/no_pdl.json 2>> $O/all_on.err; echo rc=$?
GA_FUSED_DISABLE=geglu $B > $O/no_geglu.json 2>> $O/all_on.err; echo rc=$?
GA_FUSED_DISABLE=conv,resnet,geglu $B > $O/norms_only.json 2>> $O/all_on.err; echo rc=$?
GA_FUSED_NORM=0 $B > $O/stock.json 2>> $O/all_on.err; echo rc=$?
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'],4), 'img/s', round(d['ms_per_step'],1), 'ms', list(d['seed_checksums'].values())[:1])
    except Exception as e: print(f, 'ERR', e)
P

set -x
B="python bench.py --steps 64 --warmup 8 --no-cpu-baseline"
pick() { python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['roofline']['frac'])"; }
$B | pick graph_nopdl
TCE_USE_PDL=1 $B | pick graph_pdl_late
TCE_USE_PDL=1 TCE_PDL_EARLY=1 $B | pick graph_pdl_early
TCE_USE_PDL=1 TCE_PDL_EARLY=2 $B | pick graph_pdl_early_no2cta
TCE_NO_GRAPH=1 $B | pick eager_nopdl
TCE_NO_GRAPH=1 TCE_USE_PDL=1 $B | pick eager_pdl_late
TCE_NO_GRAPH=1 TCE_USE_PDL=1 TCE_PDL_EARLY=1 $B | pick eager_pdl_early
TCE_NO_GRAPH=1 TCE_USE_PDL=1 TCE_PDL_EARLY=2 $B | pick eager_pdl_early_no2cta

run() { echo "== $1"; env $1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());r=d['regions'];print(round(d['device_ms_per_step'],2),'e2e',round(d['e2e']['value']),'flow',round(r['flow']['ms_per_step'],2),'mrf0',round(r['dec.mrf0']['ms_per_step'],2),'mrf1',round(r['dec.mrf1']['ms_per_step'],2),'mrf2',round(r['dec.mrf2']['ms_per_step'],2),'ups',round(r['dec.up0']['ms_per_step']+r['dec.up1']['ms_per_step']+r['dec.up2']['ms_per_step'],2))"; }
for a in "$@"; do run "$a"; done

for h in 0 1; do for st in 5 40; do
OJB_HOST_HEADERS=$h python bench.py --steps $st --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('host_headers=$h steps=$st value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), [(c['workload'][:5], round(c['Mpixels_per_s'])) for c in d['detail']['configs']])"
done; done

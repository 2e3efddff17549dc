(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -k query -x -q 2>&1 | tail -6)
for W in 4 2 1; do
python bench.py --steps 2 --warmup 2 --queries 37888 --no-cpu-baseline --no-recall --opt pruned_words=$W 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('W',$W,'qps',round(d['value']),'e2e',round(d['e2e']['value']),'launch_s',round(r['launch_seconds'],5),'QT',r['query_tile'],'hbm_frac',round(r['frac'],3),'smem',round(r['smem_gather']['frac'],3),r.get('pruning'))
"
done

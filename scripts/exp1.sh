for B in 65536 1048576; do
python bench.py --steps 1 --warmup 1 --queries 9472 --no-cpu-baseline --no-recall --opt boot_rows=$B 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('boot',$B,'value',round(d['value']),'launch_s',r['launch_seconds'],'share',r['kernel_share_of_step'],r['other_scan_kernel_share_of_step'],r.get('pruning'))
"
done

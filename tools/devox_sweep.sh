for t in 128 192 256 384; do for bkb in 24 32 40 48; do echo "threads=$t budget=$bkb"; BDM_DEVOX_THREADS=$t BDM_DEVOX_BUDGET_KB=$bkb python tools/op_bench.py --iters 10 --no-ref 2>&1 | grep -E "devoxelize \(plan" | grep "4096, 32" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ',d['shape'],round(d['ours_ms']*1000,1),'us',round(d['ours_frac']*100,1),'%')"; done; done

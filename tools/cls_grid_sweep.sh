for m in 2 3 4 6 8; do echo "CTAS_PER_SM=$m"; TRI_CLS_CTAS_PER_SM=$m timeout 200 python tools/bench_classify.py --skip-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('  %-8s %-18s %.4f s' % (r['dataset'], r['mode'], r['gpu_s']))
"; done

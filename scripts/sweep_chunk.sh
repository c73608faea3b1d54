for c in 64 256 1024 8192; do HB_PCL_CHUNK_IMGS=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk',$c,'value',round(d['value']),'ms',round(d['ms_per_step'],2),'bwd_ms',round(d['roofline']['families']['pcl_bwd (mid+img kernels)']['ms'],2))"; done

python -m pytest tests/test_voxel_gpu.py -m gpu -x -q 2>&1 | tail -1
for ce in 0 1 2; do python bench.py --steps 10 --warmup 3 --host-output 0 --clustered-every $ce 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ce', $ce, round(d['value']), round(d['ms_per_step'],3), 'splat', round(d['roofline']['kernel_share_of_step']['tri_strip_splat']*d['ms_per_step'],3))"; done
python tools/sweep.py --only trilinear 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if not l.strip(): continue
    r=json.loads(l)
    if r.get('kernel')=='voxel_trilinear' and r['mode']=='ordered' and r['events_per_frame']>=300000: print(r['events_per_frame'], r['clustered'], round(r['frames_per_s']))
"

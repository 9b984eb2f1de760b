python -m pytest tests/test_voxel_gpu.py -m gpu -q -x 2>&1 | tail -2
OESS_STRIP_PERSIST=24 python -m pytest tests/test_voxel_gpu.py -m gpu -q -x 2>&1 | tail -2
for ce in 2 0 1; do
python tools/bench_voxel.py --clustered-every $ce X=0 OESS_STRIP_PERSIST=16 OESS_STRIP_PERSIST=24 OESS_STRIP_PERSIST=32 OESS_STRIP_PERSIST=6,OESS_STRIP_WARPS=8 OESS_STRIP_PERSIST=12,OESS_STRIP_WARPS=4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['config'], d['ms_per_step'], d['bit_equal_to_first'], d['kernel_ms']['tri_strip_splat'])
"
done

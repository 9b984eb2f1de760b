python bench.py --steps 10 --warmup 3 --host-output 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k: round(v*d['ms_per_step'],3) for k,v in d['roofline']['kernel_share_of_step'].items()})"
python -m pytest tests/test_voxel_gpu.py -m gpu -x -q 2>&1 | tail -2

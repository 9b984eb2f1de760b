set -x
python -m pytest tests/test_voxel_gpu.py tests/test_dsec_slicer.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_voxel.py OESS_TRI_GROUPS=1 OESS_TRI_GROUPS=2 OESS_TRI_GROUPS=3 2>&1 | tee gpurun_out/r02f_groups.jsonl
python tools/bench_voxel.py --clustered-every 0 OESS_TRI_GROUPS=1 OESS_TRI_GROUPS=3 2>&1 | tee -a gpurun_out/r02f_groups.jsonl
python tools/bench_voxel.py --frames 20 OESS_TRI_GROUPS=1 OESS_TRI_GROUPS=3 2>&1 | tee -a gpurun_out/r02f_groups.jsonl
python -m pytest tests/test_models.py tests/test_pretrain_step.py -m gpu -x -q 2>&1 | tail -3

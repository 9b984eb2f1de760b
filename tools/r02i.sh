set -x
python -m pytest tests/test_losses_gpu.py -m gpu -x -q 2>&1 | tail -15
python -m pytest tests/test_pretrain_step.py -m gpu -x -q -k golden 2>&1 | grep -v Warning | tail -60
python tools/bench_infonce.py 2>&1 | tee gpurun_out/r02i_infonce.jsonl

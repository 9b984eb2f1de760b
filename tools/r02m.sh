set -x
python -m pytest tests/test_openess_step.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02m_bench_2gpu.json 2> gpurun_out/r02m_bench_2gpu.err
tail -5 gpurun_out/r02m_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02m_bench_2gpu_reference.json 2>/dev/null
nvidia-smi topo -m > gpurun_out/r02m_topo.txt 2>&1

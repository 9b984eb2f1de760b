# 2-GPU pass (gpurun --gpus 2): bench both arms under torchrun, configs 4 / 5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/r02_bench_2gpu_reference.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
tail -c 400 gpurun_out/r02_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_configs.py --config4 > gpurun_out/r02_configs_2gpu.jsonl 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/bench_configs.py --config5 >> gpurun_out/r02_configs_2gpu.jsonl 2>/dev/null
ls -la gpurun_out | grep 2gpu

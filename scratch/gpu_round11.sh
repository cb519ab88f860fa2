set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tests/dist_gpu_check.py 2>&1 | grep -v "^W\|^\[W" | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r11_bench_10m_2gpu.json 2> gpurun_out/r11_2gpu.err; tail -3 gpurun_out/r11_2gpu.err
python scratch/show2.py gpurun_out/r11_bench_10m_2gpu.json
timeout 300 python scratch/asm_bench.py 216,216,216 2>&1 | tail -6

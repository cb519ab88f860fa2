set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r15_bench_10m_8gpu.json 2> gpurun_out/r15_8gpu.err; tail -3 gpurun_out/r15_8gpu.err
python scratch/show2.py gpurun_out/r15_bench_10m_8gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r15_bench_10m_4gpu.json 2> gpurun_out/r15_4gpu.err; tail -3 gpurun_out/r15_4gpu.err
python scratch/show2.py gpurun_out/r15_bench_10m_4gpu.json

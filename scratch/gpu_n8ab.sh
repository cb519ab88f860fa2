N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_10m_g8.json 2> gpurun_out/bench_10m_g8.err; echo "rc=$?"
JB_OVERLAP=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_10m_g8_overlap.json 2> gpurun_out/bench_10m_g8_overlap.err; echo "rc=$?"

N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_10m_g$N.json 2> gpurun_out/bench_10m_g$N.err; echo "rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_10m_g$N.err | tail -20; wc -l gpurun_out/bench_10m_g$N.json

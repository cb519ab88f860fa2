mkdir -p gpurun_out
timeout 150 stdbuf -oL scratch/bb/barrier_bench 2>&1 | tee gpurun_out/r3_barrier_bench.txt

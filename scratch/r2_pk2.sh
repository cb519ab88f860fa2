mkdir -p gpurun_out
python scratch/r2_dbg_pk.py 2>&1 | tail -40
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15

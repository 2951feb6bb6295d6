mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|Warning\|^$" | tail -15
python scripts/time_k4.py | tail -8

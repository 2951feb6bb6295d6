mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "grid or post or dgei or pd or api" 2>&1 | grep -v "^  \|Warning\|^$" | tail -8
python scripts/time_k4.py | tail -8

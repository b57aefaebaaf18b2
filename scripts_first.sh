set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -40

python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200

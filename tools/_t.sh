python -m pytest tests/test_gmf_gpu.py -m gpu -x -q 2>&1 | tail -15
python bench.py --workload gmf 2>&1 | tail -1 | cut -c1-900

python bench.py --workload trajectory 2>&1 | tail -1 | cut -c1-1300

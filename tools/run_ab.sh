python bench.py --workload live 2>&1 | tail -3 | cut -c1-1500

for p in 300 10000; do python tools/profile_case.py --points $p --iters 20 --repeat 4 --noprof | tail -1; python tools/profile_case.py --points $p --iters 20 --repeat 4 | tail -1; done

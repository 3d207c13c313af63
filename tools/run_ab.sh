python tools/profile_case.py --iters 20 --grid 0 --repeat 2 | tail -1
python tools/profile_case.py --iters 20 --grid 0.085 --repeat 2 | tail -1
python tools/profile_case.py --iters 20 --grid 0.095 --repeat 2 | tail -1
ICPB_GRID_LIGHT_CM=30 python tools/profile_case.py --iters 20 --grid 0 --repeat 2 | tail -1
ICPB_GRID_LIGHT_CM=60 python tools/profile_case.py --iters 20 --grid 0 --repeat 2 | tail -1
python tools/profile_case.py --points 10000 --iters 20 --grid 0 --repeat 2 | tail -1
python tools/profile_case.py --points 60000 --iters 20 --grid 0 --repeat 2 | tail -1
python tools/profile_case.py --points 60000 --iters 20 --repeat 2 | tail -1

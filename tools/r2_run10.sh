cd $GRAFT_REPO_ROOT
python tools/profile_slab.py 0,500
python tools/profile_slab.py 0,289,348,440,500
timeout 300 ncu --set full --clock-control none --import-source on -k regex:map_rays -s 14 -c 1 -o gpurun_out/r2_rays_slab -f python tools/profile_slab.py 0,289,348,440,500 > gpurun_out/r2_ncu_slab.log 2>&1

cd $GRAFT_REPO_ROOT
python tools/profile_slab.py 0,500
python tools/profile_slab.py 0,289,348,440,500
python tools/profile_slab.py 0,200,289,320,348,400,440,470,500

cd /root/repo 2>/dev/null || cd $GRAFT_REPO_ROOT
for sl in 2 3 4 5; do NX=256 NY=256 SORT=1 SORTLOG=$sl timeout 200 python tools/time_phases.py 15625000 lean 2>&1 | grep "np=" | sed "s/^/sortlog=$sl /"; done
for sl in 2 3 4; do SORT=1 SORTLOG=$sl timeout 200 python tools/time_phases.py 12500000 lean 2>&1 | grep "np=" | sed "s/^/128x128 sortlog=$sl /"; done

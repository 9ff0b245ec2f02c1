MESHGEN_OPTIONS=pdl=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 1500 bash tests/ab_variants.sh opt:pdl=1 opt:pdl=1 2>&1 | tail -4
WORKLOAD=c2 timeout 1500 bash tests/ab_variants.sh opt:pdl=1 opt:pdl=1 2>&1 | tail -4
WORKLOAD=c1 timeout 1500 bash tests/ab_variants.sh opt:pdl=1 2>&1 | tail -4

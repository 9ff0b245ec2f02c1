timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 1500 bash tests/ab_variants.sh opt:observe_blocks=20 2>&1 | tail -3
WORKLOAD=c2 timeout 1500 bash tests/ab_variants.sh 2>&1 | tail -2

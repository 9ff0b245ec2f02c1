timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1500 bash tests/ab_variants.sh u24 2>&1 | tail -3
WORKLOAD=c4 timeout 1500 bash tests/ab_variants.sh u24 2>&1 | tail -3

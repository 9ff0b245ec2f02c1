timeout 1500 bash tests/ab_variants.sh et ur8 etur8 2>&1 | tail -5

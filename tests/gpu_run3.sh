timeout 1500 bash tests/ab_variants.sh af af 2>&1 | tail -4

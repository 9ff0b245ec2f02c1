timeout 900 python -m pytest tests -m gpu -x -q -s -k "more_domains" 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4

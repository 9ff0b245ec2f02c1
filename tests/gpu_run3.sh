timeout 900 python -m pytest tests -m gpu -x -q -k "move" 2>&1 | tail -12

timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CONFIGS=3 timeout 600 python tools/configs.py 2>/dev/null | tail -3

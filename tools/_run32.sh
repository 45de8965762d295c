echo "== default"; python tools/shard_check.py 2>&1 | tail -2
python -m pytest tests -m gpu -q -x -k "intile or linear or full_size" 2>&1 | tail -3

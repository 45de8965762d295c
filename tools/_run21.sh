python -m pytest tests -m gpu -q -x -k "not ddp_nccl and not two_ranks" 2>&1 | tail -5
for v in 1 0; do CPGB_EPILOGUE_STREAM=$v python bench.py --no-extras --steps 30 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('epilogue_stream=$v', d['ms_per_step'], d['loss'])"; done

python -m pytest tests -m gpu -q -k "fused_bn or padded_channels" 2>&1 | tail -3
for cfg in "8 16" "4 32" "2 64" "1 64"; do set -- $cfg; echo "== CS $1 TARGET $2"; CPGB_BN_CS=$1 CPGB_BN_TARGET_CHUNKS=$2 python tools/bn_ab.py 2 2>&1 | grep -v sum; done

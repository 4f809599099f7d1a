set -x
python -m kurosiwo_b200.build 2>&1 | tail -1
timeout 300 python scripts/dbg_epilogue.py 2>&1 | tail -8
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -k "permute or bn_forward" 2>&1 | tail -3

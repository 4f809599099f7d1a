set -x
timeout 120 python scripts/dbg_epilogue.py 2>&1 | tail -8

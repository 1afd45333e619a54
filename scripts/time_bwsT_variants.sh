#!/bin/bash
# dev (GPU box): full-sweep time of every variants/lib_bwsT_*.so
cd "$(dirname "$0")/.."
for v in variants/lib_bwsT_*.so; do echo "== $v"; MHPC_B200_LIB=$PWD/$v MHPC_BWS=T python scripts/bws_time.py ${1:-4096} 2>&1 | tail -1; done

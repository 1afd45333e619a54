#!/bin/bash
# dev: on the GPU box, time the kernels of every library in variants/ (scripts/lin_classes.py by default)
cd "$(dirname "$0")/.."
cp mhpc_minimal_env_b200/libmhpc_b200.so /tmp/lib_default.so
for v in variants/lib_*.so; do
  cp "$v" mhpc_minimal_env_b200/libmhpc_b200.so
  echo "== $v"
  python ${SCRIPT:-scripts/lin_classes.py} ${1:-4096}
done
cp /tmp/lib_default.so mhpc_minimal_env_b200/libmhpc_b200.so

#!/bin/bash
# dev: on the GPU box, time the per-class linearisation kernels of every library in variants/
cd "$(dirname "$0")/.."
cp mhpc_minimal_env_b200/libmhpc_b200.so /tmp/lib_default.so
for v in variants/lib_*.so; do
  cp "$v" mhpc_minimal_env_b200/libmhpc_b200.so
  echo "== $v"
  python scripts/lin_classes.py ${1:-4096}
done
cp /tmp/lib_default.so mhpc_minimal_env_b200/libmhpc_b200.so

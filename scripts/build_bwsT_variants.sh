#!/bin/bash
# dev: variants of the tensor-core backward sweep (k_bwsT.cu) into variants/lib_bwsT_<name>.so; timed with scripts/time_bwsT_variants.sh
#   scripts/build_bwsT_variants.sh name1 "flags1" name2 "flags2" ...
set -e
cd /root/repo
mkdir -p variants
CS=mhpc_minimal_env_b200/csrc; OB=mhpc_minimal_env_b200/build
while [ $# -gt 1 ]; do
  n=$1; f=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $f -I $CS -c $CS/k_bwsT.cu -o /tmp/k_bwsT_$n.o
  objs=$(ls $OB/*.o | grep -v k_bwsT.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_bwsT_$n.so $objs /tmp/k_bwsT_$n.o -Xcompiler -fPIC -cudart shared
  echo "built $n"
done

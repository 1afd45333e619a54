#!/bin/bash
# dev: build a few register-allocation / residency variants of the library into variants/ (timed on the GPU by scripts/time_variants.sh)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
build() { # name cap smem minblocks lookahead far
  MHPC_RA_CAP=$2 MHPC_RA_SMEM=$3 MHPC_LIN_MINBLOCKS=$4 MHPC_RA_LOOKAHEAD=${5:-24} MHPC_RA_FAR=${6:-128} python -m mhpc_minimal_env_b200.build > /dev/null
  cp mhpc_minimal_env_b200/libmhpc_b200.so variants/lib_$1.so
  echo "built $1"
}
build c64_s72_b3 64 72 3 12 48
build c56_s72_b3 56 72 3 10 40
build c44_s54_b4 44 54 4 8 24
build c104_s64_b2 104 64 2
python -m mhpc_minimal_env_b200.build > /dev/null   # back to the default

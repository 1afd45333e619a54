for lib in "" variants/lib_knotmajor.so; do
  echo "== lib ${lib:-default}"
  MHPC_B200_LIB=$lib python scripts/bws_time.py 4096
  MHPC_B200_LIB=$lib python scripts/rollout_time.py 4096 4
  MHPC_B200_LIB=$lib python bench.py --cpu-sample 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['stage_ms'])"
done
MHPC_B200_LIB=variants/lib_knotmajor.so python -m pytest tests -m gpu -x -q -k "ladder or receding or cost_only or mappings or odd" 2>&1 | tail -2

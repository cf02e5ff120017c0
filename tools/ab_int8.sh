#!/usr/bin/env bash
# int8-input counterpart of ab_variants.sh: the short bench command per library build (GNSSB200_LIB), kernel time of the 64-stream run
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
for lib in "$ROOT/gnss_sdr_ru_b200/libgnssb200.so" "$ROOT"/gnss_sdr_ru_b200/variants/*.so; do
  [ -f "$lib" ] || continue
  echo "== $(basename "$lib")"
  GNSSB200_LIB="$lib" python "$ROOT/bench.py" --steps 3 --warmup 3 --seconds 2 --no-cpu --no-also --no-acq --fmt int8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('int8 64 streams: kernel_ms', d['roofline']['kernel_ms'], 'value', d['value'])"
done

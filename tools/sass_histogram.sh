#!/usr/bin/env bash
# Opcode histogram (cuobjdump -sass) of the shipped library's hot kernels -> profiles/<tag>_sass_opcodes.txt.  Runs without a GPU.
TAG=${1:-r02}
SO=gnss_sdr_ru_b200/libgnssb200.so
OUT=profiles/${TAG}_sass_opcodes.txt
{
  echo "# cuobjdump -sass $SO (sm_100a only: $(cuobjdump -lelf $SO | tr '\n' ' '))"
  for pat in 'track_ws_kernelILi6ELi1ELi96ELi11' 'track_ws_kernelILi5ELi1ELi128ELi0' 'track_ws_kernelILi3ELi0ELi256ELi0' 'acq_rows_kernel' 'gsa_weak_kernel'; do
    echo; echo "## kernel matching $pat"
    cuobjdump -sass $SO | awk -v pat="$pat" '/Function : /{f = index($0, pat) > 0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); print op}' | sed -E 's/\..*//' | sort | uniq -c | sort -rn | awk '{printf "%7d %s\n", $1, $2}'
  done
  echo; echo "## markers: UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier, REDUX = warp reduction; no UTMALDG / UTCMMA / LDTM (no tensor-map loads, no tcgen05: nothing on this path is a dense contraction)"
  cuobjdump -sass $SO | grep -oE "\b(UBLKCP|SYNCS|REDUX|UTMALDG|UTCHMMA|UTCIMMA|UTCQMMA|LDTM|HMMA|IMMA)\b" | sort | uniq -c
} > $OUT
wc -l $OUT

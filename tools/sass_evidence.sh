#!/bin/bash
# SASS evidence for the kernels of the hot path (runs here: cuobjdump on the in-tree library, no GPU needed):
# per-kernel mnemonic histogram + the lines that prove tcgen05 / TMA / packed fp16 min-max / packed fp32 FMA.
# usage: tools/sass_evidence.sh > profiles/r2_sass_evidence.txt
LIB=stvo_pl_b200/lib/libplstvo_b200.so
echo "# cuobjdump -sass of $LIB ($(date -u +%Y-%m-%d)), nvcc $(nvcc --version | grep -o 'release [0-9.]*')"
for K in tc_hamming_kernel tc_expand_kernel tc_resolve_kernel gn_loop_stream_kernel gn_eval_stream_kernel stream_prepare_kernel stream_outlier_kernel track_solve_kernel; do
  for F in $(cuobjdump -elf $LIB 2>/dev/null | grep -oE "_Z[A-Za-z0-9_]*${K}[A-Za-z0-9_]*" | grep -v "_param_" | grep -v "^_ZZ" | sort -u | head -3); do
    echo; echo "## $K  ($F)"
    cuobjdump -sass -fun "$F" $LIB 2>/dev/null > /tmp/sass_$$.txt
    echo "instructions: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/sass_$$.txt)"
    echo "mnemonics (top 24): $(grep -oE '^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T] )?[A-Z][A-Z0-9_]*' /tmp/sass_$$.txt | awk '{print $NF}' | sort | uniq -c | sort -rn | head -24 | awk '{printf "%s x%s, ", $2, $1}')"
    echo "evidence lines:"
    grep -E "UTCQMMA|UTCHMMA|LDTM|UTCBAR|UBLKCP|UTMALDG|SYNCS|VHMNMX|HMNMX2|FFMA2|FMUL2|DFMA|MUFU|POPC|ATOMS|REDUX" /tmp/sass_$$.txt | sed -E 's/^\s+//' | awk '{m=$2; if (m ~ /^@/) m=$3; c[m]++; if (c[m] <= 2) print "  " $0}' | cut -c1-150
  done
done
rm -f /tmp/sass_$$.txt

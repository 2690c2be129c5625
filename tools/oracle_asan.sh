#!/bin/bash
# Runs the CPU test suite with an AddressSanitizer + UBSan build of the oracle (test infrastructure hygiene: the checker must
# not read or write out of bounds either).  Needs the system gcc's libasan.
set -e
cd "$(dirname "$0")/.."
OUT=${TMPDIR:-/tmp}/libplstvo_oracle_asan.so
/usr/bin/gcc -O1 -g -std=gnu11 -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -shared \
    -o "$OUT" oracle/plstvo_oracle.c -lm -lpthread
LD_PRELOAD=$(/usr/bin/gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 \
    PLSTVO_ORACLE_LIB="$OUT" python -m pytest tests -q -m "not gpu" -x "$@"

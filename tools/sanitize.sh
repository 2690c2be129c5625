#!/bin/bash
# compute-sanitizer passes (memcheck, racecheck, synccheck) over one small invocation of every kernel (GPU box)
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_run.py 2>&1 | tail -8
done

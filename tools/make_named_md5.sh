#!/bin/bash
# Oracle-driver runs of the big named systems (no golden exists for them): the md5 of each `-o` file
# goes to tests/golden/named_md5.txt, which tests/test_gpu_e2e.py compares the product binary's output with.
# The reduced Groebner basis is unique, so the pair cap only changes the step sequence: --max-spairs 0 (fewer,
# larger steps) is used where it is the faster run. Checker mode of the oracle (all host threads). CPU only.
set -u
cd "$(dirname "$0")/.."
OUT=${1:-/tmp/named_md5}
mkdir -p "$OUT"
T=${GAMBA_ORACLE_THREADS:-$(nproc)}
for spec in cyclic9-15:2000 cyclic9-31:2000 kat14-15:0 eco14-15:2000 eco15-31:0 kat15-15:0 kat15-31:0; do
  name=${spec%%:*}; sp=${spec##*:}
  [ -s "$OUT/$name.md5" ] && continue
  python -c "from gamba_b200 import systems; systems.write_system('$name', '$OUT/$name.txt')"
  t0=$(date +%s)
  GAMBA_ORACLE_THREADS=$T oracle/_build/gamba_oracle -i "$OUT/$name.txt" -o "$OUT/$name.out" -v 1 --max-spairs $sp > "$OUT/$name.log" 2>&1 \
    && md5sum < "$OUT/$name.out" | cut -d' ' -f1 > "$OUT/$name.md5"
  echo "$name $(cat $OUT/$name.md5) max_spairs=$sp $(( $(date +%s) - t0 )) s"
done

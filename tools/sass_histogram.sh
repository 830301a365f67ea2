#!/bin/sh
# SASS evidence of the built library (run here, no GPU needed): opcode counts per kernel family.
# usage: tools/sass_histogram.sh > profiles/r2_sass_histogram.txt
LIB=${1:-pyhf_b200/libhf_b200.so}
echo "# cuobjdump -sass $LIB  (sm_100a cubins only: $(cuobjdump -lelf $LIB | grep -c sm_100a) ELF, $(cuobjdump -lelf $LIB | grep -vc sm_100a) other)"
cuobjdump -sass "$LIB" > /tmp/hf_sass.txt
echo "## whole library"
for op in DMMA DFMA DADD DMUL UBLKCP SYNCS LDGSTS REDG ATOMS MUFU SHFL BAR; do
  printf "%-8s %s\n" $op "$(grep -cE "^\s+/\*[0-9a-f]+\*/\s+$op" /tmp/hf_sass.txt)"
done
echo "## per kernel: DMMA / DFMA / UBLKCP / SYNCS / BAR"
awk '/Function :/ {name=$3} /^\s+\/\*[0-9a-f]+\*\/\s+[A-Z]/ {op=$2; sub(/\..*/,"",op); c[name" "op]++; k[name]=1}
     END {for (n in k) printf "%s DMMA=%d DFMA=%d UBLKCP=%d SYNCS=%d BAR=%d\n", n, c[n" DMMA"], c[n" DFMA"], c[n" UBLKCP"], c[n" SYNCS"], c[n" BAR"]}' /tmp/hf_sass.txt | sort | c++filt | sed 's/(DevPlan.*//' | cut -c1-160

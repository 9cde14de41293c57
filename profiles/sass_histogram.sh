#!/bin/bash
# SASS opcode histogram per kernel of libflowtimes.so: the Blackwell-native evidence (UTCHMMA = tcgen05.mma,
# LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = TMA tensor load, UBLKCP = bulk copy, LDGSTS = cp.async,
# UCGABAR = cluster barrier).
#   bash profiles/sass_histogram.sh > profiles/r2_sass_histogram.txt        (no GPU needed: cuobjdump on the built .so)
lib=flow-timesnet_b200/lib/libflowtimes.so
echo "# cuobjdump -sass $lib (git $(git rev-parse --short HEAD)); counts of the tcgen05 / TMEM / TMA / cluster opcodes per kernel"
printf "%-58s %8s %6s %7s %8s %7s %7s %8s %8s\n" kernel UTCHMMA LDTM UTCBAR UTMALDG UBLKCP LDGSTS UCGABAR total
cuobjdump -sass "$lib" | awk '
  function flush() {
    if (name == "") return
    printf "%s %d %d %d %d %d %d %d %d\n", name, c["UTCHMMA"], c["LDTM"], c["UTCBAR"], c["UTMALDG"], c["UBLKCP"], c["LDGSTS"], c["UCGABAR_ARV"] + c["UCGABAR_WAIT"], total
  }
  /Function : / { flush(); name = $3; delete c; total = 0; next }
  /^[ \t]+\/\*[0-9a-f]+\*\// { op = $2; if (substr(op, 1, 1) == "@") op = $3; sub(/\..*/, "", op); sub(/;/, "", op); total++; c[op]++ }
  END { flush() }' | while read -r name a b c d e f g t; do
    if [ $((a + b + c + d + e + f + g)) -gt 0 ]; then
      short=$(echo "$name" | c++filt | sed -e 's/^void //' -e 's/ftn:://g' -e 's/(.*//' | cut -c1-58)
      printf "%-58s %8d %6d %7d %8d %7d %7d %8d %8d\n" "$short" "$a" "$b" "$c" "$d" "$e" "$f" "$g" "$t"
    fi
  done | sort -k2,2nr -k1,1

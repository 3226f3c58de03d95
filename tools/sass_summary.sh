#!/bin/bash
# SASS opcode evidence per object (no GPU needed): counts of the Blackwell-native mnemonics in every built object.
#   bash tools/sass_summary.sh > profiles/r02_sass_summary.txt
cd "$(dirname "$0")/../abstractgps.jl_b200/csrc/build" || exit 1
for o in *.o; do
  echo "== $o"
  cuobjdump -sass "$o" 2>/dev/null | grep -oE "\b(UTCIMMA|UTCHMMA|UTCQMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|SYNCS|DMMA|HMMA|IMMA|FFMA|DFMA|LDGSTS|ELECT|R2UR|UTCATOMSWS|UCGABAR_ARV|UCGABAR_WAIT)[A-Z0-9_.]*" | sed -E 's/\.(reuse|E|64|128|U32|S32)$//' | sort | uniq -c | sort -rn | head -25
done

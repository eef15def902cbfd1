#!/bin/bash
# Evidence: the TMA bulk-copy + mbarrier instructions in the hot kernels of the shipped library (the inflate kernel
# and its pull / dual instantiations, the parse kernel, the block scan).
# usage: profiles/sass_excerpt.sh > profiles/rNN_sass_tma_excerpt.txt
LIB="$(dirname "$0")/../parallelparsing_b200/lib/libppb200.so"
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$0; keep = (fn ~ /pp_inflate_kernel/ || fn ~ /pp_inflate_pull_kernel/ || fn ~ /pp_inflate_dual_kernel/ || fn ~ /pp_parse_kernel/ || fn ~ /pp_blockscan_kernel/); if (keep) { print ""; print fn; n[fn]=0 } next }
keep && /UBLKCP|SYNCS|UTMA|MBAR|ATOMS|REDUX/ { print "    " $0 }
keep && /\/\*[0-9a-f]{4}\*\// { cnt[fn]++ }
END { print ""; for (f in cnt) print "instructions in" f ": " cnt[f] }'

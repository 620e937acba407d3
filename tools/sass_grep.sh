#!/bin/bash
# SASS evidence that the tower kernels run on tcgen05 / TMEM / TMA (profiles/r02_sass_grep.txt): mnemonic counts per kernel
# from the built library (cuobjdump -sass), so the claim survives without the git-ignored objects.
cd "$(dirname "$0")/.."
LIB=rl-selfplay-mnk_b200/libmnk_b200.so
cuobjdump -sass $LIB | awk '
  /Function : / { fn=$3 }
  { for (i=1;i<=NF;i++) if ($i ~ /^(UTCHMMA|UTCQMMA|UTCBAR|UTCCP|LDTM|STTM|UBLKCP|UTMALDG|SYNCS|BAR\.SYNC|BAR\.ARV|FENCE\.VIEW\.ASYNC|UTCATOMSWS|REDUX|SHFL|ATOMG|ATOMS|LDGSTS)/) { split($i,a,"."); c[fn" "a[1]]++ } }
  END { for (k in c) print k, c[k] }' | sort | c++filt | awk '{ n=$NF; $NF=""; printf "%-110s %6d\n", $0, n }'

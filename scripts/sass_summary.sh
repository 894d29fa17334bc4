#!/bin/bash
# Counts of the Blackwell-only SASS mnemonics per kernel of libzernike_b200.so (evidence that the hot kernels are
# tcgen05 / TMEM / TMA code):  UTCHMMA / UTCQMMA = tcgen05.mma (f16 / tf32 kinds), .2CTA = cta_group::2,
# UTMALDG = TMA tensor loads (MULTICAST = cluster multicast), LDTM / STTM = tcgen05.ld / .st, UTCBAR = tcgen05.commit.
# Usage: scripts/sass_summary.sh > profiles/r02_sass_summary.txt
set -e
LIB="$(dirname "$0")/../motif-learn_b200/lib/libzernike_b200.so"
echo "# $(basename "$LIB"): $(stat -c %s "$LIB") bytes, $(cuobjdump -lelf "$LIB" | grep -c sm_100a) sm_100a cubin(s)"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { name=$3 }
  /UTC.MMA/ { mma[name]++; if ($0 ~ /2CTA/) mma2[name]++ }
  /UTMALDG/ { tma[name]++; if ($0 ~ /MULTICAST/) mc[name]++ }
  /UTMASTG|UBLKCP/ { bulk[name]++ }
  /LDTM/ { ldtm[name]++ }
  /STTM/ { sttm[name]++ }
  /UTCBAR/ { bar[name]++ }
  /SYNCS/ { syncs[name]++ }
  END {
    printf "%-100s %8s %6s %8s %6s %6s %6s %7s %7s\n", "kernel", "UTC*MMA", "2CTA", "UTMALDG", "MCAST", "LDTM", "STTM", "UTCBAR", "SYNCS"
    for (k in syncs) if (mma[k] || tma[k] || ldtm[k]) printf "%-100s %8d %6d %8d %6d %6d %6d %7d %7d\n", substr(k,1,100), mma[k], mma2[k], tma[k], mc[k], ldtm[k], sttm[k], bar[k], syncs[k]
  }' | sort

#!/bin/bash
# Counts of the SASS mnemonics that prove which hardware paths the kernels use (B200_PROFILING.md), per object file.
#   bash scripts/sass_markers.sh > profiles/r02_sass_markers.txt
cd "$(dirname "$0")/.."
echo "# cuobjdump -sass simple-vector-db_b200/lib/*.o | grep -c <mnemonic>   (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)"
echo "# UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA tiled load), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,"
echo "# UBLKCP = cp.async.bulk (1-D bulk copy), SYNCS = mbarrier, DMMA = FP64 tensor core, IDP = __dp4a (K13 integer keys), LDGSTS = cp.async, REDUX = warp reduce"
printf "%-22s" object; for m in UTCHMMA UTMALDG LDTM UTCBAR UBLKCP SYNCS DMMA IDP LDGSTS REDUX; do printf "%9s" $m; done; echo
for o in simple-vector-db_b200/lib/*.o; do
  cuobjdump -sass "$o" > /tmp/sass_markers.$$ 2>/dev/null
  printf "%-22s" "$(basename $o)"
  for m in UTCHMMA UTMALDG LDTM UTCBAR UBLKCP SYNCS DMMA IDP LDGSTS REDUX; do printf "%9s" "$(grep -c "$m" /tmp/sass_markers.$$)"; done; echo
done
rm -f /tmp/sass_markers.$$
echo
echo "# kernels (entry functions) per object"
for o in simple-vector-db_b200/lib/*.o; do
  echo "## $(basename $o)"; cuobjdump -sass "$o" 2>/dev/null | grep "Function :" | sed 's/.*Function : //' | c++filt | sed 's/(.*//' | sort | uniq -c | sort -rn | head -40
done

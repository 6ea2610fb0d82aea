#!/bin/bash
# Blackwell-specific SASS opcodes of the built library, whole library and per kernel (profiles/r2_sass_opcodes.txt):
#   UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit,
#   UTCATOMSWS = tcgen05.alloc / dealloc / relinquish, UBLKCP = cp.async.bulk (1-D bulk copy engine, .S.G global->shared,
#   .G.S shared->global), SYNCS = mbarrier ops, UTMACMDFLUSH = bulk-group commit / wait, REDG = red.global
LIB=${1:-r2l_b200/csrc/libr2l_b200.so}
PAT='UTC[A-Z]*MMA[.A-Z0-9_]*|LDTM[.A-Zx0-9_]*|STTM[.A-Zx0-9_]*|UBLKCP[.A-Z0-9_]*|UTMA[A-Z.0-9_]*|UTCBAR[.A-Z0-9_]*|UTCATOMSWS[.A-Z0-9_]*|SYNCS[.A-Z0-9_]*|REDG[.A-Z0-9_]*|MEMBAR[.A-Z0-9_]*|ACQBULK|CCTL[.A-Z0-9_]*'
echo "# $(basename $LIB): $(nvcc --version | grep release | sed 's/.*release //'), cuobjdump -sass, opcode counts (static instructions)"
echo "## whole library"
cuobjdump -sass "$LIB" | grep -oE "$PAT" | sort | uniq -c | sort -rn
echo
echo "## per kernel (kernels that contain at least one of the opcodes above)"
cuobjdump -sass "$LIB" | awk -v pat="$PAT" '
  /Function :/ { if (name != "" && n > 0) { printf "%s\n", name; for (k in c) printf "    %6d %s\n", c[k], k; } delete c; n = 0; name = $0; sub(/.*Function : /, "", name); next }
  { line = $0; while (match(line, pat)) { op = substr(line, RSTART, RLENGTH); c[op]++; n++; line = substr(line, RSTART + RLENGTH); } }
  END { if (name != "" && n > 0) { printf "%s\n", name; for (k in c) printf "    %6d %s\n", c[k], k; } }' | c++filt

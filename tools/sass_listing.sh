#!/bin/bash
# SASS of the dominant kernel (k_gemm_tc) from the built library + the opcode histogram that proves the tcgen05 / TMEM /
# bulk-copy path (UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit,
# SYNCS = mbarrier).  Writes profiles/<R>_k_gemm_tc.sass (+ .hist).
R=${1:-r2}
SO=aby3_b200/libaby3cu.so
cuobjdump -sass $SO | awk '/Function : .*k_gemm_tc/{p=1} /Function : /{if(p&&!/k_gemm_tc/)exit} p{print}' | sed 's#/\* 0x[0-9a-f]* \*/##' | sed 's/[[:space:]]*$//' | grep -v '^$' > profiles/${R}_k_gemm_tc.sass
grep -oE '^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+' profiles/${R}_k_gemm_tc.sass | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn > profiles/${R}_k_gemm_tc.hist
wc -l profiles/${R}_k_gemm_tc.sass; grep -E "UTCIMMA|LDTM|UBLKCP|UTCBAR|SYNCS|UBLKPF|UTCATOM" profiles/${R}_k_gemm_tc.hist

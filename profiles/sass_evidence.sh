#!/bin/bash
# Per-kernel SASS evidence for the claims in DESIGN.md (TMA bulk copies + mbarrier, 128-bit global accesses, warp match,
# no tensor-core ops in the memory-bound kernels).  Needs no GPU:  bash profiles/sass_evidence.sh > profiles/r2_sass_evidence.txt
SO=pnp_ovss_b200/libpnp_ovss_b200.so
echo "# cuobjdump -sass $SO : instruction counts per kernel (sm_100a); built from csrc $(python -c 'import bench; print(bench.source_fingerprint())')"
echo "# columns: UBLKCP (cp.async.bulk = TMA 1-D bulk copy) | SYNCS (mbarrier) | LDG.E.128 | STG.E.128 | LDS.128 | MATCH | RED/ATOM | MUFU | FFMA | HMMA (mma.sync) | UTCHMMA (tcgen05.mma) | LDTM/STTM (tcgen05.ld/st) | UTCBAR (tcgen05.commit) | LDGSTS (cp.async)"
cuobjdump -sass $SO 2>/dev/null | awk '
/Function : /{ f=$3; order[++n]=f }
/UBLKCP/{a[f]++} /SYNCS/{b[f]++} /LDG\.E(\.[A-Z0-9]+)*\.128/{c[f]++} /STG\.E(\.[A-Z0-9]+)*\.128/{d[f]++} /LDS(\.U)?\.128/{e[f]++}
/MATCH/{g[f]++} /(RED|ATOMG|ATOMS|ATOM)\./{h[f]++} /MUFU/{i[f]++} /FFMA/{j[f]++} / HMMA/{k[f]++} /UTC[A-Z]*MMA/{l[f]++} /(LDTM|STTM)/{m[f]++} /UTCBAR/{o[f]++} /LDGSTS/{q[f]++}
END{ for(x=1;x<=n;x++){ f=order[x]; printf "%6d %6d %6d %6d %6d %6d %6d %6d %6d %6d %6d %6d %6d %6d  %s\n", a[f],b[f],c[f],d[f],e[f],g[f],h[f],i[f],j[f],k[f],l[f],m[f],o[f],q[f],f } }' | c++filt | sed 's/(.*//' | sort -k15

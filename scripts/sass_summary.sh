#!/bin/bash
# SASS evidence per kernel of libwssdl_b200.so: counts of the instructions that prove TMA /
# mbarrier / bulk copies / 256-bit stores / cluster barriers / PDL / warp reductions are in the
# built code (cuobjdump -sass; no GPU needed).   usage: scripts/sass_summary.sh > profiles/rNN_sass_summary.txt
SO=${1:-wssdl_bus_b200/libwssdl_b200.so}
echo "# cuobjdump -sass $SO (sm_100a), instruction counts per kernel; $(nvcc --version | tail -1)"
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3; order[++n]=fn; next }
  fn=="" { next }
  {
    if ($0 ~ /UTMALDG/) c[fn,"UTMALDG"]++
    if ($0 ~ /UTMAPF/) c[fn,"UTMAPF"]++
    if ($0 ~ /UBLKCP/) c[fn,"UBLKCP"]++
    if ($0 ~ /SYNCS/) c[fn,"SYNCS"]++
    if ($0 ~ /LDGSTS/) c[fn,"LDGSTS"]++
    if ($0 ~ /STG\.E[^ ]*\.256/) c[fn,"STG.256"]++
    if ($0 ~ /STG\.E[^ ]*\.128/) c[fn,"STG.128"]++
    if ($0 ~ /LDS[^ ]*\.128/) c[fn,"LDS.128"]++
    if ($0 ~ /UCGABAR/) c[fn,"UCGABAR"]++
    if ($0 ~ /ACQBULK/) c[fn,"ACQBULK"]++
    if ($0 ~ /PREEXIT/) c[fn,"PREEXIT"]++
    if ($0 ~ /REDUX/) c[fn,"REDUX"]++
    if ($0 ~ /MATCH/) c[fn,"MATCH"]++
    if ($0 ~ /ATOMS/) c[fn,"ATOMS"]++
    if ($0 ~ /RED\.E/) c[fn,"RED"]++
    if ($0 ~ /DADD|DMUL|DFMA/) c[fn,"FP64"]++
    if ($0 ~ /\/\*[0-9a-f]{4}\*\//) c[fn,"total"]++
  }
  END {
    split("total UTMALDG UTMAPF UBLKCP SYNCS LDGSTS LDS.128 STG.256 STG.128 UCGABAR ACQBULK PREEXIT REDUX MATCH ATOMS RED FP64", k, " ")
    for (i=1;i<=n;i++) {
      fn=order[i]; line=""
      for (j=1;j<=17;j++) if (c[fn,k[j]]>0) line=line sprintf(" %s=%d", k[j], c[fn,k[j]])
      cmd="echo " fn " | c++filt"; cmd | getline dem; close(cmd)
      gsub(/\(anonymous namespace\)::/,"",dem)
      print substr(dem,1,110) "\n   " line
    }
  }'
echo "# legend: UTMALDG = cp.async.bulk.tensor (TMA tile load), UTMAPF = TMA L2 prefetch, UBLKCP = cp.async.bulk (bulk copy),"
echo "# SYNCS = mbarrier ops, LDGSTS = cp.async, UCGABAR = cluster barrier, ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents (PDL),"
echo "# REDUX = __reduce_*_sync, MATCH = __match_any_sync, ATOMS = shared atomics, RED = global reductions, FP64 = DADD/DMUL/DFMA"

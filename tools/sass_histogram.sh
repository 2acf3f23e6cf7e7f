#!/bin/bash
# SASS opcode histogram of the two tcgen05 kernels of aki_b200/libaki_mma.so (what proves the Blackwell-native path:
# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMAREDG = TMA load / reduce, UGETNEXTWORKID = cluster launch
# control, SYNCS = mbarrier, FFMA2/FADD2/FMNMX3 = packed fp32 math).  usage: tools/sass_histogram.sh > profiles/<name>.txt
cd "$(dirname "$0")/.."
for k in attn_fwd_sm100_kernelILb1 attn_bwd_sm100_kernel; do
  echo "== $k (aki_b200/libaki_mma.so, sm_100a)"
  cuobjdump -sass aki_b200/libaki_mma.so | awk -v k="$k" '/Function :/ {on = index($0, k) > 0} on && $1 ~ /^\/\*[0-9a-f]+\*\/$/ {op=$2; if (op ~ /^@/) op=$3; sub(/;$/, "", op); n[op]++} END {for (o in n) printf "%7d %s\n", n[o], o}' | sort -rn
done
# the decode-size linear layer: warp-level HMMA fed by 16-byte no-allocate loads, chained by programmatic dependent launch
# (ACQBULK = griddepcontrol.wait, PREEXIT = griddepcontrol.launch_dependents)
k=skinny_linear_kernelILi4ELi0
echo "== $k (aki_b200/libaki_mma.so, sm_100a)"
cuobjdump -sass aki_b200/libaki_mma.so | awk -v k="$k" '/Function :/ {on = index($0, k) > 0} on && $1 ~ /^\/\*[0-9a-f]+\*\/$/ {op=$2; if (op ~ /^@/) op=$3; sub(/;$/, "", op); n[op]++} END {for (o in n) printf "%7d %s\n", n[o], o}' | sort -rn | head -14

#!/bin/bash
# Opcode evidence per translation unit of libym_b200.so (B200_PROFILING.md "What proves a Blackwell-native kernel"):
#   tools/sass_summary.sh > profiles/r02_sass_summary.txt
cd "$(dirname "$0")/../yolo-master_b200/csrc" || exit 1
echo "# cuobjdump -sass <unit>.o | opcode counts (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a); UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,"
echo "# UTMALDG/UTMASTG = TMA tensor loads/stores, UBLKCP = cp.async.bulk, LDGSTS = cp.async, HMMA = mma.sync, ACQBULK/PREEXIT/... = griddepcontrol (PDL)"
printf "%-18s %8s %8s %6s %6s %8s %8s %7s %7s %6s %8s %8s\n" unit UTCHMMA UTC.2CTA LDTM STTM UTMALDG UTMASTG UBLKCP LDGSTS HMMA MUFU.EX2 SYNCS
for f in *.o; do
  s=$(cuobjdump -sass "$f" 2>/dev/null)
  c() { echo "$s" | grep -cE "$1"; }
  printf "%-18s %8d %8d %6d %6d %8d %8d %7d %7d %6d %8d %8d\n" "${f%.o}" "$(c 'UTCHMMA')" "$(c 'UTCHMMA\.2CTA')" "$(c 'LDTM')" "$(c 'STTM')" \
     "$(c 'UTMALDG')" "$(c 'UTMASTG')" "$(c 'UBLKCP')" "$(c 'LDGSTS')" "$(c ' HMMA')" "$(c 'MUFU\.EX2')" "$(c 'SYNCS')"
done
echo
echo "# kernels per unit (cuobjdump -elf symbol names, demangled)"
for f in tc_attention2.o tc_attention.o tc_conv.o tc_dispatch2.o tc_dispatch.o tc_gemm.o; do
  echo "## $f"; cuobjdump -sass "$f" 2>/dev/null | grep "Function :" | sed 's/.*Function : //' | c++filt | cut -c1-150
done

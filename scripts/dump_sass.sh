#!/bin/bash
# SASS listings of the contraction kernels of the shipped library -> profiles/ (no GPU needed: cuobjdump reads the .so).
# One file per kernel plus a mnemonic census that shows the tcgen05 / TMA / TMEM instructions at a glance.
set -e
cd "$(dirname "$0")/.."
LIB=financial_rag_system_b200/csrc/libfrs_b200.so
R=${1:-r2}
declare -A K=(
  [scan_kernel_bf16]='_ZN3frs11scan_kernelILb0ELb0EEEv14CUtensorMap_stS1_NS_10ScanParamsE'
  [scan_kernel_f32]='_ZN3frs11scan_kernelILb1ELb0EEEv14CUtensorMap_stS1_NS_10ScanParamsE'
  [gemm_kernel_qkv]='_ZN3frs11gemm_kernelILi192ELi0EEEv14CUtensorMap_stS1_S1_S1_NS_10GemmParamsE'
  [gemm_kernel_gelu]='_ZN3frs11gemm_kernelILi192ELi1EEEv14CUtensorMap_stS1_S1_S1_NS_10GemmParamsE'
  [gemm_kernel_resln]='_ZN3frs11gemm_kernelILi192ELi2EEEv14CUtensorMap_stS1_S1_S1_NS_10GemmParamsE'
  [attention_kernel]='_ZN3frs16attention_kernelE14CUtensorMap_stS0_S0_NS_10AttnParamsE'
  [merge_kernel_bf16]='_ZN3frs12merge_kernelILb0EEEvNS_11MergeParamsE'
)
OUT=profiles/${R}_sass_census.txt
echo "# SASS mnemonic census of libfrs_b200.so ($(date -u +%F), nvcc $(nvcc --version | grep -o 'release [0-9.]*')), per kernel:" > $OUT
echo "# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store, SYNCS = mbarrier" >> $OUT
for name in "${!K[@]}"; do
  f=profiles/${R}_sass_${name}.txt
  cuobjdump -sass -fun "${K[$name]}" $LIB 2>/dev/null | grep -E '^\s+/\*[0-9a-f]{4,6}\*/' | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s/[[:space:]]+$//' > $f
  echo "== $name ($(wc -l < $f) instructions) -> $f" >> $OUT
  awk '{print $2}' $f | sed 's/;//' | grep -E '^(UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|SYNCS|UTCBAR|UTCATOM|MUFU|HMMA|DFMA|FFMA2|FADD2|FMNMX3|F2FP|REDUX|ATOMS|ATOMG|LDG|STG|LDS|STS|LDL|STL|BAR|USETMAXREG)' | sort | uniq -c | sort -rn | awk '{printf "     %6d %s\n", $1, $2}' >> $OUT
done
echo wrote $OUT

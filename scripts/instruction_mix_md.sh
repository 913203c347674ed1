#!/bin/bash
# Dynamic SASS instruction mix per warp of the hot kernels from an `ncu --set full --import-source on` report.
#   scripts/instruction_mix_md.sh <report.ncu-rep> <warps per launch> > profiles/<tag>_instruction_mix.md
rep=$1; warps=$2
echo "# dynamic instruction mix of the hot kernels ($(basename $rep), \`ncu --set full --import-source on\`)"
echo
echo "Per warp (32 rows) = instructions executed / $warps warps; \`xN\` = blocks of SASS executed N times per warp; digest by \`scripts/sass_hist.py\`."
for k in session_knn_rerank_kernel session_nvt_smooth_kernel session_nvt_classify_kernel session_update_kernel session_class_max_pruned_kernel; do
  ncu -i $rep --page source --csv --kernel-name regex:$k --print-source sass 2>/dev/null > /tmp/src_$k.csv
  echo; echo "## $k"; echo; echo '```'
  python $(dirname $0)/sass_hist.py /tmp/src_$k.csv $warps 2>&1 | tail -8
  echo '```'
done

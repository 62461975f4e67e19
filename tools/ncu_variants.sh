#!/bin/bash
# developer A/B helper (GPU box): instruction-cache counters of the path kernel for library variants
#   tools/ncu_variants.sh out.csv cur wpc16 perstep ...
out=$1; shift
M=gpu__time_duration.sum,sm__inst_executed.sum,sm__icc_requests.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__pcsamp_warps_issue_stalled_barrier,smsp__pcsamp_warps_issue_stalled_no_instructions,smsp__pcsamp_sample_buffer_full
: > $out
for v in "$@"; do
  if [ "$v" = cur ]; then unset FSD_LIBFSDPLAN; else export FSD_LIBFSDPLAN=$PWD/ft_fsd_path_planning_b200/csrc/ab_$v.so; fi
  echo "== $v" >> $out
  ncu --metrics $M --kernel-name regex:path_kernel -c 1 --launch-skip 2 --clock-control none --csv python tools/profile_target.py 10240 3 stage 2>/dev/null | grep -E "path_kernel" | awk -F'","' '{print $(NF-2)" "$(NF)}' | tr -d '"' >> $out
done
cat $out

#!/bin/bash
# Round-end evidence job (run through gpurun): GPU test suite, stage timeline, ncu launch list, per-kernel memory metrics,
# ncu --set full captures of the kernels DESIGN.md discusses, compute-sanitizer logs.
set -u
O=gpurun_out
TAG=${TAG:-r4}
timeout 900 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
timeout 300 python tools/stage_timeline.py > $O/${TAG}_stage_timeline.txt 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv python tools/profile_step.py > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file $O/${TAG}_kernel_metrics.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__inst_executed_pipe_tc.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed \
  -k regex:'gather_gemm_tc|attention_tc|segmented_sum|nms_sweep|nms_mask|subm3_table|topk_select|layernorm|down2_fill|trim_' python tools/profile_step.py > /dev/null 2>&1
for spec in "gather_gemm_tc_kernel:3:3:gemm" "attention_tc_kernel:0:1:attention" "segmented_sum_kernel:0:1:segsum" "nms_sweep_kernel:0:1:nms" "subm3_table_kernel:0:1:subm3" "topk_select_kernel:0:1:topk"; do
  IFS=: read k s c name <<< "$spec"
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o $O/${TAG}_$name python tools/profile_step.py > /dev/null 2>&1
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_memcheck.log 2>&1; echo "rc=$?" >> $O/${TAG}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_racecheck.log 2>&1; echo "rc=$?" >> $O/${TAG}_racecheck.log
ls -la $O | tail -30
tail -3 $O/${TAG}_pytest_gpu.log

for v in base priv4 priv16; do
  SSRS_B200_LIB=$PWD/variants/libssrs_$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_var_$v.json 2> gpurun_out/r02_var_$v.err || tail -5 gpurun_out/r02_var_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_var_$v.json')); print('VAR $v', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms_alone'])"
done
M=gpu__time_duration.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_tag_requests.max.pct_of_peak_sustained_elapsed,lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed,lts__t_requests_srcunit_tex_op_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__inst_executed.sum
for v in nored priv16; do
  SSRS_B200_LIB=$PWD/variants/libssrs_$v.so ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_$v.csv -k regex:step_tracks --launch-skip 136 --launch-count 4 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/ncu_$v.log 2>&1
done

# one-off: launch list of the config-3 chain (noise K=3 -> mode 5) and of the heat-map sequence, final tree
set -u
mkdir -p gpurun_out
python scripts/seq_probe.py --density 100000 --frames 40 --reps 2 --mode 5 --noise 3 > gpurun_out/cfg3_probe.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_conv|k_stream|k_thresh|k_binar" -c 24 --csv --log-file gpurun_out/cfg3_launches.csv python scripts/seq_probe.py --density 100000 --frames 40 --reps 2 --mode 5 --noise 3 > gpurun_out/ncu_cfg3.log 2>&1
tail -n 2 gpurun_out/cfg3_probe.log gpurun_out/ncu_cfg3.log
bash scripts/gpu_check.sh

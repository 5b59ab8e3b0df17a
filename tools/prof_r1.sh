set -x
tools/bin/mufu_bench
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active"
# launch list of the second (warm) train step: the first step has 1050/.. launches; skip them
STEPS=2 timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches.csv python tools/one_step.py > gpurun_out/r1c_one_step.log 2>&1
STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_fwd2 -s 35 -c 1 -o gpurun_out/r1c_block_fwd2 -f python tools/one_step.py >> gpurun_out/r1c_one_step.log 2>&1
STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_bwd2 -s 35 -c 1 -o gpurun_out/r1c_block_bwd2 -f python tools/one_step.py >> gpurun_out/r1c_one_step.log 2>&1
STEPS=200 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gen_steps_bf16 -s 1 -c 1 -o gpurun_out/r1c_gen_bf16 -f python tools/gen_run.py >> gpurun_out/r1c_one_step.log 2>&1
tail -5 gpurun_out/r1c_one_step.log
ls -la gpurun_out/

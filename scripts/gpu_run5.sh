cd /root/repo
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -o gpurun_out/prof_conv_stream_v5 -f python tests/debug_prof5.py > gpurun_out/ncu_stream.log 2>&1
tail -3 gpurun_out/ncu_stream.log
ls -la gpurun_out/*.ncu-rep

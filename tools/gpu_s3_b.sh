mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_linear_kernel -s 50 -c 13 -f \
    -o gpurun_out/prof_decode_linear_s3 python tools/bench_decode.py --no-graphs --steps 16 > gpurun_out/ncu_decode_linear_s3.log 2>&1
tail -3 gpurun_out/ncu_decode_linear_s3.log
ls -la gpurun_out/prof_decode_linear_s3.ncu-rep

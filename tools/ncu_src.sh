mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 0 -c 2 -f -o gpurun_out/src_conv_enc1 python tests/dev_layer_times.py > gpurun_out/src_conv_enc1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 14 -c 1 -f -o gpurun_out/src_conv_dec2b python tests/dev_layer_times.py > gpurun_out/src_conv_dec2b.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:wgrad_igemm_kernel -s 18 -c 2 -f -o gpurun_out/src_wgrad_enc1 python tests/dev_layer_times.py > gpurun_out/src_wgrad_enc1.log 2>&1
ls -la gpurun_out

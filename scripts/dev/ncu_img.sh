#!/bin/bash
mkdir -p gpurun_out
cap() {
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -o gpurun_out/prof_img_$1 -f python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
}
cap fwd_3_64 'conv_gemm_kernel<[^0-9]*64[^0-9]+6[^0-9]+(0|false)' 1 1
cap fwd_64_3 'conv_gemm_kernel<[^0-9]*16[^0-9]+6[^0-9]+(1|true)' 1 1
cap dgrad_3_64 'conv_gemm_kernel<[^0-9]*16[^0-9]+3[^0-9]+(1|true)' 1 1
ls -la gpurun_out

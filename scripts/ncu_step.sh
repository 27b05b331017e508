#!/bin/bash
# ncu captures for profiles/ (run under gpurun, 1 GPU): launch list of the profiled steps + full sets of the dominant kernels.
# scripts/profile_step.py runs 3 warm-up steps and one measured eager step on a single stream.
# The reports of the template-heavy GEMM module are large: the raw metric pages are extracted on the box and reports above
# 15 MB are dropped, so that gpurun_out/ stays under the 64 MiB that travel back.
TAG=${1:-r2}
mkdir -p gpurun_out
if [ -z "$SKIP_LIST" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
fi
cap() {  # name, demangled-name regex, launches to skip, launches to capture
  ncu --set full --clock-control none --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -o gpurun_out/prof_${TAG}_$1 -f python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_$1_raw.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/prof_${TAG}_$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 15000000 ]; then rm -f gpurun_out/prof_${TAG}_$1.ncu-rep; fi
}
cap conv_fwd 'conv_gemm_pair_kernel<[^0-9]*256[^0-9]+3[^0-9]+(1|true)' 2 1     # 3rd launch of the step: D forward conv 256->256 @8x8, 512 images (two fp16 planes, 3 products)
cap conv_dgrad 'conv_gemm_pair_kernel<[^0-9]*256[^0-9]+3[^0-9]+(1|true)' 5 1   # 6th launch: D input gradient 512->512 @4x4, 768 rows-images (two bf16 planes, 3 products)
cap conv_n128 'conv_gemm_pair_kernel<[^0-9]*128[^0-9]+3[^0-9]+(1|true)' 2 1    # a 256 x 128 pair tile launch with concatenated weight planes
cap conv_dgrad_n64 'conv_gemm_kernel<[^0-9]*64[^0-9]+3[^0-9]+(1|true)' 0 1   # N = 64 (G transposed conv 128->64 forward / D 64->128 input gradient)
cap wgrad 'wgrad_gemm_kernel<[^0-9]*256[^0-9]+3' 3 1                 # a 128-row weight-gradient tile launch (TMA-staged gather), batch-1 SN launches skipped
cap wgrad_pair 'wgrad_gemm_pair_kernel<[^0-9]*256[^0-9]+3' 12 3      # a 256-row CTA-pair weight-gradient launch
cap refresh 'refresh_kernel' 0 2                                     # operand refresh of D and of G after the update
cap mmd 'mmd_' 0 1
python scripts/profile_step.py cifar 256 3 > gpurun_out/events_${TAG}.txt 2>&1
du -sh gpurun_out; ls -la gpurun_out/ | tail -16

# A/B of two builds of the CUDA library on one GPU: kernel ms of whole frames and of the 8 shards of a frame
# usage: bash tools/ab_probe.sh librtb200_x.so [workloads]
set -e
ALT=$1; WL=${2:-p5_sah_4k,p5_rgrid_4k,p5_kd_4k,p5_fgrid_4k}
for rep in 1 2; do
  for lib in librtb200.so $ALT; do
    RTB_CUDA_LIB_NAME=$lib python tools/frame_probe.py --workloads $WL --reps 10 --tag "$lib" | sed 's/first-frame.*| steady/steady/; s/ -> .*//'
  done
done
for lib in librtb200.so $ALT; do
  RTB_CUDA_LIB_NAME=$lib python tools/frame_probe.py --workloads p5_sah_4k,p5_rgrid_4k --world 8 --col-block 32 --reps 7 --tag "$lib" | sed 's/first-frame.*| steady/steady/'
done

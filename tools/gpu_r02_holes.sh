#!/bin/bash
# r02 visit 1: ncu --set full of the worst sweep rows (BG1 R=1/3 Z=208, 144, 224 float32; Z=288 packed-half; Z=384 for
# comparison) + batch-size dependence of the same rows (wave quantisation of the ~100 MB sweep batches).
mkdir -p gpurun_out/r02
cap() {  # name, args...
  local name=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:decode_nms -s 3 -c 1 -f -o gpurun_out/r02/$name \
    python tools/gpu_point.py "$@" --reps 1 > gpurun_out/r02/$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/r02/$name.ncu-rep gpurun_out/r02/$name.summary.txt > /dev/null
  rm -f gpurun_out/r02/$name.ncu-rep
}
cap z208_f32 --bg 1 --Z 208 --rate 1/3 --dtype f32
cap z144_f32 --bg 1 --Z 144 --rate 1/3 --dtype f32
cap z224_f32 --bg 1 --Z 224 --rate 1/3 --dtype f32
cap z288_f16 --bg 1 --Z 288 --rate 1/3 --dtype f16x2
cap z384_f32 --bg 1 --Z 384 --rate 1/3 --dtype f32
cap z384_f16 --bg 1 --Z 384 --rate 1/3 --dtype f16x2 --batch 4096
for Z in 144 208 224 288 320 384; do
  for mb in 100 400; do
    for dt in f32 f16x2; do
      python tools/gpu_point.py --bg 1 --Z $Z --rate 1/3 --dtype $dt --mb $mb --reps 10
    done
  done
done 2>&1 | grep '^{' | tee gpurun_out/r02/batch_dependence.jsonl
ls -la gpurun_out/r02

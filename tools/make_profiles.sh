#!/bin/bash
# ncu --set full captures of the dominant kernels at the bench sizes -> profiles/r02_*_raw.csv + meta
# (run under gpurun from the repo root; the CSV export happens on the box, only text comes back)
set -e
mkdir -p gpurun_out profiles
COMMIT=$(cat .git_commit_for_profiles 2>/dev/null || echo unknown)
run() { # tag what kernel-regex skip points extra-args
  local tag=$1 what=$2 regex=$3 skip=$4 pts=$5; shift 5
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 \
      -o gpurun_out/${tag} python tools/profile_targets.py $what "$@" > gpurun_out/${tag}.log 2>&1
  ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${tag}_src.csv 2>/dev/null
  python tools/ncu_lines.py gpurun_out/${tag}_src.csv 40 > gpurun_out/${tag}_lines.txt
  echo "{\"commit\": \"$COMMIT\", \"points\": $pts, \"command\": \"ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 python tools/profile_targets.py $what $*\"}" > gpurun_out/${tag}.meta.json
}
run r02_head c4 k_normals2 1 10000000 --n 10000000 --k 16 --reps 2
run r02_c4   c4 k_normals2 1 10000000 --n 10000000 --k 30 --reps 2
run r02_c2   c2 k_normals2 1 120000 --reps 2
run r02_c3   c3 k_icp_correspond 50 1000000 --reps 2 --iters 30
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_head.csv python tools/profile_targets.py c4 --n 10000000 --k 16 --reps 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4.csv python tools/profile_targets.py c4 --n 10000000 --k 30 --reps 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c2.csv python tools/profile_targets.py c2 --reps 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c3.csv python tools/profile_targets.py c3 --reps 2 --iters 30 > /dev/null 2>&1
ls -la gpurun_out/r02_*

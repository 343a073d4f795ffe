set -e
COMMIT=$(cat .git_commit_for_profiles)
tag=r02_c3
ncu --set full --clock-control none --import-source on -k regex:k_icp_correspond -s 50 -c 1 -o gpurun_out/${tag} -f python tools/profile_targets.py c3 --reps 2 --iters 30 > gpurun_out/${tag}.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${tag}_src.csv 2>/dev/null
python tools/ncu_lines.py gpurun_out/${tag}_src.csv 40 > gpurun_out/${tag}_lines.txt
echo "{\"commit\": \"$COMMIT\", \"points\": 1000000, \"command\": \"ncu --set full --clock-control none --import-source on -k regex:k_icp_correspond -s 50 -c 1 python tools/profile_targets.py c3 --reps 2 --iters 30\"}" > gpurun_out/${tag}.meta.json

# round-2 ncu evidence (one GPU).  Every command first runs plain, then under ncu.
cd $GRAFT_REPO_ROOT
set -x
# 1. the hint kernel as one rank of the 8-GPU partition-sharded run sees it: 2 sub-PIRs over its own 400 228 rows
python scripts/time_hintgen.py --n 400228 --batch 4 --iters 3 > gpurun_out/p_hg8_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hintgen_kernel -s 2 -c 1 -o gpurun_out/r02_hintgen_rank_of_8 python scripts/time_hintgen.py --n 400228 --batch 4 --iters 3 > gpurun_out/p_hg8_ncu.log 2>&1
# 2. the lock-step search step: launch list + full captures of the two client-side kernels
python scripts/time_search_device.py --lanes 32 --rounds 2 > gpurun_out/p_sd_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_search_launches.csv python scripts/time_search_device.py --lanes 32 --rounds 2 > gpurun_out/p_sd_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:client_prepare_kernel -s 10 -c 1 -o gpurun_out/r02_client_prepare python scripts/time_search_device.py --lanes 32 --rounds 2 > gpurun_out/p_sd_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_step_kernel -s 10 -c 1 -o gpurun_out/r02_search_step python scripts/time_search_device.py --lanes 32 --rounds 2 > gpurun_out/p_sd_ncu3.log 2>&1
# 3. the two HBM-bound stragglers of round 1
python scripts/prof_stragglers.py > gpurun_out/p_st_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"answer_kernel|ip_stream_kernel" -s 2 -c 1 -o gpurun_out/r02_answer_640B python scripts/prof_stragglers.py > gpurun_out/p_st_ncu1.log 2>&1
ncu --set full --clock-control none -k regex:ip_stream_kernel -s 2 -c 1 -o gpurun_out/r02_ip_stream_q1 python scripts/prof_stragglers.py > gpurun_out/p_st_ncu2.log 2>&1
# 4. launch list of the bench step
python bench.py --steps 2 --warmup 3 --no-search --no-other-configs --no-cpu-baseline > gpurun_out/p_b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-search --no-other-configs --no-cpu-baseline > gpurun_out/p_b_ncu.log 2>&1
ls -la gpurun_out/r02_*

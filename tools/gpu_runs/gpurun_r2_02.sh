#!/bin/bash
# round 2, second call: fused Fisher kernel (new library) -- GPU tier, cfg3 fused vs unfused, noise report, launch list and
# --set full of k_fisher_fused (reports converted to CSV on the box: the .ncu-rep files exceed the 64 MiB return limit)
O=gpurun_out/r2_02
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest.log
python bench.py --config 3 --steps 3 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
GWAT_B200_FISHER_UNFUSED=1 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_cfg3_unfused.json 2> $O/bench_cfg3_unfused.err
python tools/fisher_noise_report.py --sources 64 > $O/fisher_noise.json 2> $O/fisher_noise.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_fisher.csv \
    python bench.py --config 3 --steps 1 --warmup 3 --fisher-sources 4000 --no-cpu-baseline > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fisher_fused -s 2 -c 1 -o /tmp/prof_fused \
    python bench.py --config 3 --steps 1 --warmup 3 --fisher-sources 4000 --no-cpu-baseline > $O/ncu_fused.log 2>&1
ncu -i /tmp/prof_fused.ncu-rep --page raw --csv > $O/ncu_full_k_fisher_fused_raw.csv 2>/dev/null
ncu -i /tmp/prof_fused.ncu-rep --page details --csv > $O/ncu_full_k_fisher_fused_details.csv 2>/dev/null
ncu -i /tmp/prof_fused.ncu-rep --page source --csv 2>/dev/null | gzip > $O/ncu_source_k_fisher_fused.csv.gz
tail -25 $O/pytest.log; head -c 1500 $O/bench_cfg3.json; echo; tail -3 $O/bench_cfg3.err; head -c 400 $O/bench_cfg3_unfused.json; echo; cat $O/fisher_noise.json; tail -3 $O/fisher_noise.err; head -c 300 $O/bench_default.json; tail -3 $O/bench_default.err; du -sh gpurun_out

#!/bin/bash
# round 2, first call: GPU tier incl. the new full-size parity tests, the restructured bench (all configs in one line),
# the Fisher noise report, and a launch list + --set full of the two Fisher kernels (post-tiling).
mkdir -p gpurun_out/r2_01
O=gpurun_out/r2_01
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
python bench.py --config 3 --steps 3 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
python tools/fisher_noise_report.py --sources 64 > $O/fisher_noise.json 2> $O/fisher_noise.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_fisher.csv \
    python bench.py --config 3 --steps 1 --warmup 3 --fisher-sources 2000 --no-cpu-baseline > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fisher_deriv -s 2 -c 1 -o $O/prof_fisher_deriv \
    python bench.py --config 3 --steps 1 --warmup 3 --fisher-sources 2000 --no-cpu-baseline > $O/ncu_deriv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fisher_assemble -s 2 -c 1 -o $O/prof_fisher_assemble \
    python bench.py --config 3 --steps 1 --warmup 3 --fisher-sources 2000 --no-cpu-baseline > $O/ncu_asm.log 2>&1
tail -3 $O/pytest.log; tail -2 $O/smoke.log; head -c 600 $O/bench_default.json; echo; tail -3 $O/bench_default.err; head -c 1200 $O/bench_cfg3.json; echo; cat $O/fisher_noise.json; tail -3 $O/fisher_noise.err

#!/bin/bash
# round 2, third call: rolled setup loops, reference-compiled sampler tests, NCCL-capable sampler (1 GPU here), no-FMA Fisher variant
O=gpurun_out/r2_03
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
python bench.py --workload sampler --steps 200 --warmup 20 > $O/bench_sampler.json 2> $O/bench_sampler.err
GWAT_B200_LIB=$PWD/variants/libgwat_b200_nofma.so python tools/fisher_noise_report.py --sources 64 > $O/fisher_noise_nofma.json 2> $O/fisher_noise_nofma.err
python tools/fisher_noise_report.py --sources 64 > $O/fisher_noise.json 2> $O/fisher_noise.err
for c in 1 2; do
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 24 --csv --log-file $O/launches_cfg$c.csv \
    python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_launch_cfg$c.log 2>&1
done
tail -25 $O/pytest.log; head -c 400 $O/bench_default.json; echo; tail -3 $O/bench_default.err; head -c 1500 $O/bench_sampler.json; echo; tail -5 $O/bench_sampler.err; cat $O/fisher_noise_nofma.json; tail -3 $O/fisher_noise_nofma.err; grep -h "k_setup\|k_loglike\|k_finish" $O/launches_cfg1.csv | head -6; grep -h "k_setup\|k_loglike\|k_finish" $O/launches_cfg2.csv | head -6; du -sh gpurun_out

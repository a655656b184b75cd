#!/bin/bash
# round 2, sixth call: stage stamps of the setup kernel (profiling variant), drop-in link test after the tc write-back
O=gpurun_out/r2_06
mkdir -p $O
GWAT_B200_LIB=$PWD/variants/setupprof/libgwat_b200.so python tools/setup_stage_profile.py > $O/setup_stages.json 2> $O/setup_stages.err
python -m pytest tests/test_dropin_link.py tests/test_gwatpy_dropin.py -m gpu -q 2>&1 | tail -15 > $O/pytest.log
cat $O/setup_stages.json; tail -3 $O/setup_stages.err; tail -8 $O/pytest.log

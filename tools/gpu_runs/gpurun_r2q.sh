#!/bin/bash
# sky-averaged Fisher matrices (slim variant): parity against the reference build; Fisher regression tests
export GWAT_B200_LIB=$PWD/variants/slim/libgwat_b200.so
python -m pytest tests/test_fisher_sky.py tests/test_gpu_parity.py tests/test_sampler_gpu.py -m gpu -q -k "sky or fisher or Fisher" 2>&1 | tail -12

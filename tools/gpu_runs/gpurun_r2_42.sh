#!/bin/bash
# round 2, call 42 (1 GPU): the link-level drop-in with the Fisher matrices on a grid of their own (user_param->fisher_freq / fisher_PSD / fisher_length)
python -m pytest tests/test_dropin_link.py tests/test_cxx_adapter.py tests/test_gwatpy_dropin.py -m gpu -q 2>&1 | tail -6

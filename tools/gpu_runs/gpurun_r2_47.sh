#!/bin/bash
# round 2, call 47 (1 GPU): the sampling-vector path for families / modification layouts outside the BASELINE configs, against the compiled reference
python -m pytest tests/test_mcmc_variants.py -m gpu -q 2>&1 | tail -25

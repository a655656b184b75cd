#!/bin/bash
python -m pytest tests/test_sampler_gpu.py -m gpu -q 2>&1 | tail -60

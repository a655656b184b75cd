#!/bin/bash
# round 2, call 46 (1 GPU): networks of 1, 4 and 5 detectors against the compiled reference (likelihood cfg1/2/5 shapes, Fisher), 6 and 7 refused
python -m pytest tests/test_network_sizes.py -m gpu -q 2>&1 | tail -12

#!/bin/bash
# round 2, call 48 (1 GPU): pointed Fisher matrices of the NRT / gIMR / dCS / EdGB / precessing-ppE families against the compiled reference
python -m pytest tests/test_fisher_variants.py -m gpu -q 2>&1 | tail -40

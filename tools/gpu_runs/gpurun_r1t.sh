#!/bin/bash
python -m pytest tests/test_theories.py -m gpu -q 2>&1 | tail -25
python -m pytest tests -m gpu -q --deselect tests/test_theories.py 2>&1 | tail -3

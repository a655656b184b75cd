#!/bin/bash
# round 2, sixteenth call: the two tests that failed in r2_15, then the whole GPU tier
python -m pytest tests -m gpu -q 2>&1 | tail -8

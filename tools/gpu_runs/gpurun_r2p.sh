#!/bin/bash
# final library + amplitude/phase entry points and the plain-C waveform API: whole GPU tier, smoke, default bench
python -m pytest tests -m gpu -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py 2>/dev/null | tail -c 300

#!/bin/bash
# round 2, call 26 (1 GPU): the intrinsic sampling mode (repack of the 4/8-parameter sets, maximised likelihood from sampling vectors,
# sky-averaged MCMC_ Fishers, the link-level wrappers after gwat_b200_dropin_set_intrinsic) and the orientation Fishers
python -m pytest tests/test_intrinsic.py tests/test_dropin_link.py tests/test_orientation.py tests/test_fisher_sky.py tests/test_maximized.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -15

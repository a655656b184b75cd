#!/bin/bash
# round 2, thirteenth call: does programmatic dependent launch help once the CUDA events around k_loglike are gone?
export GWAT_B200_LIB=$PWD/variants/pdl/libgwat_b200.so
echo "== events, PDL"; python tools/e2e_quick.py 1 2
echo "== events, no PDL"; GWAT_B200_NO_PDL=1 python tools/e2e_quick.py 1 2
echo "== no events, PDL"; GWAT_B200_NO_KERNEL_EVENTS=1 python tools/e2e_quick.py 1 2
echo "== no events, no PDL"; GWAT_B200_NO_KERNEL_EVENTS=1 GWAT_B200_NO_PDL=1 python tools/e2e_quick.py 1 2

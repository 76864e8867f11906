#!/bin/bash
for w in 1 0 1 0; do echo "== wide=$w"; DICOW_DECODE_WIDE=$w timeout 300 python tools/profile_decode.py 2>&1 | grep -E "kernels over|decode_linear" ; done

#!/bin/bash
for f in 0 1; do echo "== DICOW_MEGA_FLAGS=$f"; DICOW_MEGA_FLAGS=$f python tools/probe_decode_mega.py 16 2>&1 | tail -22; done

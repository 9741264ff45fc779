#!/bin/bash
# warp-busy fraction of the potential phases, whole catalogue and one GPU's eighth of it
export HALMA_DEBUG_PHASES=1
timeout 200 python scripts/cfg3_parts.py --one 0/1 --steps 2 2>&1 | grep -E "potential phase|per_pass" | tail -8
timeout 200 python scripts/cfg3_parts.py --one 0/8 --steps 2 2>&1 | grep -E "potential phase|per_pass" | tail -8

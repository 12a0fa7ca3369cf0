#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python scripts/profile_step.py qiddm_ll 4096 > gpurun_out/c14_prof_ll.txt 2>&1; echo "prof rc=$?"
QIDDM_FUSED_TAIL=0 timeout 300 python scripts/profile_step.py qiddm_ll 4096 > gpurun_out/c14_prof_ll_off.txt 2>&1; echo "prof off rc=$?"

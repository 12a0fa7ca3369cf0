#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python scripts/profile_step.py qiddm_ll 4096 > gpurun_out/t_prof_ll.txt 2>&1
python scripts/profile_step.py qiddm_pl 1024 noise > gpurun_out/t_prof_pl.txt 2>&1
python scripts/profile_step.py unet 64 > gpurun_out/t_prof_unet.txt 2>&1
head -3 gpurun_out/t_prof_*.txt

#!/bin/bash
# register-tile width of the basis-column collapse (n = 10) on the final code: forward / adjoint with RB 5 (default), 4, 3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for v in "default" "QIDDM_RB_FWD=4" "QIDDM_RB_BWD=4" "QIDDM_RB_FWD=4 QIDDM_RB_BWD=4" "QIDDM_RB_FWD=3 QIDDM_RB_BWD=3"; do
  tag=$(echo "$v" | tr ' =' '__')
  if [ "$v" = "default" ]; then env python bench.py --no-cpu-baseline --no-extras --no-secondary --steps 5 --warmup 3 > gpurun_out/c12_$tag.json 2> gpurun_out/c12_$tag.err
  else env $v python bench.py --no-cpu-baseline --no-extras --no-secondary --steps 5 --warmup 3 > gpurun_out/c12_$tag.json 2> gpurun_out/c12_$tag.err; fi
  echo "$v rc=$?"
done

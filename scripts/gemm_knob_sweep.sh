#!/bin/bash
# usage: scripts/gemm_knob_sweep.sh "<ENV1=..;ENV2=..>" ... : one short bench run per environment, prints the per-kernel ms/step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for cfg in "$@"; do
  envs=$(echo "$cfg" | tr ';' ' ')
  out=$(env $envs python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --no-secondary 2>/dev/null | tail -1)
  python - "$cfg" <<PY
import json, sys
d = json.loads('''$out''')
k = d["roofline"]["kernel_ms_per_step"]
print(sys.argv[1] or "default", "| step %.3f ms | fwd %.3f dx %.3f dw %.3f | prep %.3f gy %.3f gb %.3f | gate %.3f/%.3f | e2e %.2fM" % (
    d["ms_per_step"], k.get("gemm_forward", 0), k.get("gemm_dx", 0), k.get("gemm_dw", 0), k.get("prep_x", 0), k.get("grad_y", 0),
    k.get("g_bound", 0), k.get("gate_forward", 0), k.get("gate_backward", 0), d["e2e"]["value"] / 1e6), flush=True)
PY
done

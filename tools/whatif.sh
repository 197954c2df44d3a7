#!/bin/bash
# Step-time contribution of kernel families: re-times the config_v2_2 training step with one family turned into
# no-ops at a time (VDN_SKIP, timing only - the results of such a step are wrong by construction).
for k in "" gn_bwd ln_bwd wgrad mha_bwd sla_bwd sla_fwd mha_fwd gn_fwd tail_fwd; do
  VDN_DEBUG=1 VDN_SKIP=$k python bench.py --no-cpu-baseline --no-sampling --steps 30 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('skip=%-9s %.3f ms/step' % ('$k' or '-', d['ms_per_step']))"
done

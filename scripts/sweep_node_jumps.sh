#!/bin/bash
# hand_node_jump_kernel: compositions per launch vs time of the launches (40k x 40k), via bench.py's per-kernel profile
for j in 1 2 3 4 6 8; do
  DTB_NODE_JUMPS=$j python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['roofline']['kernels']['hand_node_jump_kernel']
print('jumps',$j,'launches',k['launches_per_step'],'ms',round(k['ms_per_step'],4),'step',round(d['ms_per_step'],3),'verified',d['verified'])"
done

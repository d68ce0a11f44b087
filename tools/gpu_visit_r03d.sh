#!/bin/bash
# round 2, last visits: the whole GPU suite + smoke on the tree as committed
O=gpurun_out/${1:-r03d}; mkdir -p $O
timeout 240 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > $O/pytest_gpu_tail.log; echo "pytest rc=${PIPESTATUS[0]}" > $O/rc.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
cat $O/pytest_gpu_tail.log $O/smoke.log $O/rc.txt

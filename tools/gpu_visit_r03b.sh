#!/bin/bash
# round 2, late visit: the whole GPU suite after the K-b2 wide path + SR flavour templates, then A/B runs of the knobs
O=gpurun_out/r03b; mkdir -p $O
timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > $O/pytest_gpu_tail.log; echo "pytest rc=${PIPESTATUS[0]}" > $O/rc.txt
for v in "OTB_NB2_WIDE=0" "OTB_NB2_UNROLL=1" "OTB_NB2_UNROLL=2"; do echo "$v" >> $O/kb2_ab.log; env $v timeout 40 python tools/kprof.py kb2 >> $O/kb2_ab.log 2>&1; done
for v in "OTB_SR_UNROLL=1" "OTB_SR_UNROLL=2"; do echo "$v" >> $O/sr_ab.log; env $v timeout 40 python tools/kprof.py sr >> $O/sr_ab.log 2>&1; done
cat $O/pytest_gpu_tail.log $O/rc.txt $O/kb2_ab.log $O/sr_ab.log

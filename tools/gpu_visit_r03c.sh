#!/bin/bash
# round 2, late visit: A/B of the block-count cap of K-b1 / K-b2 (one slab per block vs looping blocks) and of K-c's chunk size
O=gpurun_out/r03c; mkdir -p $O
ab() { k=$1; shift; for v in "$@"; do echo "[$v]" >> $O/ab.log; env $v timeout 40 python tools/kprof.py $k 2>&1 | tail -1 >> $O/ab.log; done; }
ab kb2 "X=0" "OTB_NB_WAVES=1024" "OTB_NB_WAVES=1024 OTB_NB2_UNROLL=1"
ab kb1 "X=0" "OTB_NB_WAVES=1024" "OTB_NB_WAVES=1024 OTB_NB1_UNROLL=1"
ab kc_bf16 "X=0" "OTB_KC_CHUNK=2048" "OTB_KC_CHUNK=4096" "OTB_KC_CHUNK=8192"
cat $O/ab.log

#!/bin/bash
# tests/tools/probe_matrix.sh "<variants>" "<workloads>" [extra args]: kernel_probe for each pair; main = default library
out=gpurun_out/probe_$(date +%H%M%S).jsonl
for wl in $2; do for v in $1; do
  if [ "$v" = main ]; then unset RSF_LIBRARY; else export RSF_LIBRARY=$PWD/build/alt/librsfusion_$v.so; fi
  case $wl in tum) a="60 30";; hires) a="40 20";; c5) a="30 16";; c5hires) a="24 12";; esac
  timeout 300 python tests/tools/kernel_probe.py $wl $a $3 2>gpurun_out/probe_err.txt | tail -1 | tee -a $out
done; done

#!/bin/bash
# scratch: A/B/C of prebuilt libraries in scratch/ab: bash scratch/ab3.sh "v1 v2 ..." "<cmd>" ...
VS=$1; shift
for round in 1 2; do
for v in $VS; do
  cp scratch/ab/$v.so ivlnce_b200/csrc/libivlnmap.so; touch ivlnce_b200/csrc/libivlnmap.so
  for c in "$@"; do echo -n "$v | $c | "; bash -c "$c" 2>&1 | tail -1; done
done
done

#!/bin/bash
# scratch: one round of A/B over prebuilt libraries: bash scratch/ab1.sh "v1 v2" "<cmd>"
VS=$1; shift
for v in $VS; do
  cp scratch/ab/$v.so ivlnce_b200/csrc/libivlnmap.so; touch ivlnce_b200/csrc/libivlnmap.so
  for c in "$@"; do echo "== $v | $c"; bash -c "$c" 2>&1 | tail -3; done
done

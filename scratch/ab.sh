#!/bin/bash
# scratch: A/B of two prebuilt libraries (scratch/ab/head.so, scratch/ab/new.so) on the same box: bash scratch/ab.sh "<cmd>" ...
for round in 1 2; do
for v in head new; do
  cp scratch/ab/$v.so ivlnce_b200/csrc/libivlnmap.so; touch ivlnce_b200/csrc/libivlnmap.so
  for c in "$@"; do echo -n "$v | $c | "; bash -c "$c" 2>&1 | tail -1; done
done
done

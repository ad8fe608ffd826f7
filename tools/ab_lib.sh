#!/bin/bash
# In-run A/B of two builds of libsvb_vocoder.so (box-to-box variance is ~10 %, only same-run comparisons count):
#   tools/ab_lib.sh neuralsvb_b200/libsvb_old.so "<command>"   runs <command> with the old, the new, the old, the new library
set -e
OLD=$1; shift
cp neuralsvb_b200/libsvb_vocoder.so /tmp/svb_new.so
for round in 1 2; do
  for which in old new; do
    if [ $which = old ]; then cp "$OLD" neuralsvb_b200/libsvb_vocoder.so; else cp /tmp/svb_new.so neuralsvb_b200/libsvb_vocoder.so; fi
    echo "== $which"
    bash -c "$*" || true
  done
done
cp /tmp/svb_new.so neuralsvb_b200/libsvb_vocoder.so

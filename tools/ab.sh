#!/bin/bash
# A/B kernel timing on the GPU box: tools/ab.sh "<cfgs>" <lib-tag> [<lib-tag> ...]   ("base" = product library)
cfgs="$1"; shift
for rep in 1; do
for tag in "$@"; do
  lib=normalizingflownetwork_b200/libnfn_b200.so; [ "$tag" != base ] && lib=normalizingflownetwork_b200/libnfn_b200_$tag.so
  for c in $cfgs; do
    NFN_B200_LIB=$PWD/$lib python bench.py $NFN_AB_FLAGS --config $c --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); r=j['roofline']; print('$tag'.ljust(12), '$c', 'kern_ms %.4f frac %.3f value %.3e'%(r['kernel_ms'], r['frac'], j['value']))"
  done
done
done

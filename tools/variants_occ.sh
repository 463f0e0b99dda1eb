#!/bin/bash
# usage: tools/variants_occ.sh cfgA cfgB ... -- like variants_cfg.sh, plus the residency line of k_decode (BNFLAC_TRACE)
for lib in _build/variants/*.so; do
  for c in "$@"; do
    BNFLAC_LIB=$PWD/$lib BNFLAC_TRACE=1 python tools/run_cfg.py $c 5 > /tmp/occ.out 2> /tmp/occ.err
    echo "$lib $(tail -1 /tmp/occ.out) | $(grep -h 'k_decode<' /tmp/occ.err | sort -u | head -1)"
  done
done

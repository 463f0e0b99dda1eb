#!/bin/bash
# usage: tools/variants_cfg.sh cfgA cfgB ...  -- run_cfg.py for each library variant under _build/variants
for lib in _build/variants/*.so; do
  for c in "$@"; do
    echo -n "$lib "; BNFLAC_LIB=$PWD/$lib python tools/run_cfg.py $c 5 2>&1 | tail -1
  done
done

#!/bin/bash
# usage: tools/sass_fn.sh obj pattern  -- SASS of the first function whose mangled name matches pattern (instructions only, numbered)
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/{f=($0 ~ pat)} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//' | awk '{printf "%d %s\n", NR-1, $0}'

#!/bin/bash
# usage: tools/sass_spills.sh <object> <function-substring>: local-memory (spill) instructions and hand-over markers of one kernel
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/{f=index($0,pat)>0} f' | grep -v "^\s*$" | sed 's/\/\* 0x[0-9a-f]* \*\///' | cut -c1-110 > /tmp/k.sass
grep -n "SETMAXREG\|LDL\|STL\|SYNCS.ARRIVE\|UTCBAR\|EXIT" /tmp/k.sass | awk '{$1=$1; print}' | cut -c1-100

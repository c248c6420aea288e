#!/bin/bash
# usage: tools/sass_hist.sh <binary> <function-substring>   -> opcode histogram + total for one kernel
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/ {on = index($0, pat) > 0} on {print}' > /tmp/_k.sass
grep -cE "^\s+/\*[0-9a-f]{4,6}\*/" /tmp/_k.sass | sed 's/^/total instructions: /'
grep -oE "^\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" /tmp/_k.sass | awk '{print $NF}' | sort | uniq -c | sort -rn | head -${3:-25}

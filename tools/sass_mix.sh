#!/bin/bash
# usage: tools/sass_mix.sh <so> <function-substring>  -> opcode histogram of the function's SASS
cuobjdump -sass "$1" | awk -v pat="$2" '/Function : /{f=index($0,pat)>0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{op=$2; if (op ~ /^@/) op=$3; sub(/\..*/,"",op); print op}' | sort | uniq -c | sort -rn

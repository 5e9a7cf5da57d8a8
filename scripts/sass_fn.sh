#!/bin/bash
# scripts/sass_fn.sh <object-or-so> <kernel-name-substring>: SASS of one kernel (instruction lines only)
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function :/ {on = index($0, pat) > 0} on && /^ +\/\*[0-9a-f]{4}\*\//' 

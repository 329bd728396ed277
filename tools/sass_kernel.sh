#!/bin/bash
# tools/sass_kernel.sh <lib.so> <c++filt'd name regex>  -> compact SASS listing (address opcode operands) of the first match
cuobjdump -sass "$1" 2>/dev/null | c++filt | awk -v pat="$2" '
/Function :/ { on = ($0 ~ pat) ? (found ? 0 : 1) : 0; if (on) found = 1; next }
on && /^[ \t]+\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\// { sub(/^[ \t]+\/\*/, ""); sub(/\*\/[ \t]+/, " "); sub(/[ \t]*\/\*.*$/, ""); print }'
